"""Glue between the two passes of the reference pipelines."""
from . import trajectory  # noqa: F401
from . import filt  # noqa: F401
