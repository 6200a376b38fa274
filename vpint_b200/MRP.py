"""Reference module name (VPint/MRP.py) -> vpint_b200.base."""
from .base import MRP, SMRP, STMRP  # noqa: F401
from .errors import VPintError  # noqa: F401
