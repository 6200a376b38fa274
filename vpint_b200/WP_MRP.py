"""Reference module name (VPint/WP_MRP.py) -> vpint_b200.wp."""
from .errors import VPintError  # noqa: F401
from .sd import SD_SMRP, SD_STMRP  # noqa: F401
from .wp import WP_SMRP, WP_STMRP  # noqa: F401
