"""Drop-in for the reference's `wan` package, hot path only (mount as models/Wan22/wan)."""
from . import distributed, modules  # noqa: F401
