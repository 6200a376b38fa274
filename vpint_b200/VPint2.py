"""Reference module name (VPint/VPint2.py) -> vpint_b200.vpint2."""
from .vpint2 import VPint2_interpolator, run_patches  # noqa: F401
from .wp import WP_SMRP  # noqa: F401
