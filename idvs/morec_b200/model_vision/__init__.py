"""Drop-in replacement of inbatch_sasrec_e2e_vision/model/__init__.py."""
from .model import Model  # noqa: F401
