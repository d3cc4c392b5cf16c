"""parse_args() of the vision package (reference: inbatch_sasrec_e2e_vision/parameters.py): every reference flag plus the
launcher-compat fixes (--news, --local-rank / LOCAL_RANK); implementation: idvs/morec_b200/host/params.py."""
from data_utils.utils import *  # noqa: F401,F403  (the reference's parameters.py leaks these names too)

from idvs.morec_b200.host.params import build_parser  # noqa: E402,F401
from idvs.morec_b200.host import params as _params  # noqa: E402


def parse_args(argv=None):
    return _params.parse_args(argv, kind="vision")


if __name__ == "__main__":
    print(parse_args())
