"""Reference module name (VPint/SD_MRP.py) -> vpint_b200.sd."""
from .sd import SD_SMRP, SD_STMRP  # noqa: F401
