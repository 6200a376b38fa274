"""vpint_b200 -- B200-native drop-in for VPint's value-propagation hot path.

Public surface mirrors the reference package for this path:

    from vpint_b200 import SD_SMRP, WP_SMRP, SD_STMRP, WP_STMRP, VPint2_interpolator

(also importable under the reference's module names: ``vpint_b200.SD_MRP``,
``vpint_b200.WP_MRP``, ``vpint_b200.VPint2``, ``vpint_b200.MRP``).  The loop runs
in hand-written sm_100a CUDA kernels behind the C ABI of include/vpint_b200.h;
there is no CPU fallback.
"""

from .errors import VPintError
from .base import MRP, SMRP, STMRP
from .sd import SD_SMRP, SD_STMRP
from .wp import WP_SMRP, WP_STMRP
from .vpint2 import VPint2_interpolator, run_patches

__version__ = "0.1.0"
__all__ = ["VPintError", "MRP", "SMRP", "STMRP", "SD_SMRP", "SD_STMRP", "WP_SMRP", "WP_STMRP",
           "VPint2_interpolator", "run_patches"]
