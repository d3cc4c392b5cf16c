"""Drop-in replacement of the reference's `model` package (inbatch_sasrec_e2e_text/model/__init__.py:1)."""
from .model import Model  # noqa: F401
