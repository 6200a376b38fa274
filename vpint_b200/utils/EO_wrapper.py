"""Earth-observation wrappers around the per-band solve (reference: VPint/utils/EO_wrapper.py).

Only the two functions that sit directly on either side of the propagation path are mirrored; the GeoTIFF
reader and the plotting helpers of the reference module need rasterio / matplotlib and are out of scope."""

import numpy as np

from ..vpint2 import solve_bands
from ..wp import WP_SMRP, wp_run_options


def multiband_VPint(target_img, feature_img, iterations=-1, method='exact', max_gamma=np.inf, min_gamma=0,
                    prioritise_identity=True, priority_intensity=1):
    """WP-MRP on every band independently (VPint/utils/EO_wrapper.py:15-21).

    Same signature and result as the reference: a copy of ``target_img`` (its dtype) whose band ``b`` is
    ``WP_SMRP(target_img[:, :, b], feature_img[:, :, b], max_gamma=..., min_gamma=...).run(iterations=..., method=...)``.
    Like the reference, ``prioritise_identity`` and ``priority_intensity`` are accepted and NOT forwarded to
    ``run``.  With exact weights all bands are one batched device solve; the pairwise weight methods, whose
    clamp depends on ``min_gamma`` / ``max_gamma``, run band by band."""
    pred_img = target_img.copy()
    if method == 'exact' and target_img.ndim == 3 and target_img.shape[2] > 0:
        opts = wp_run_options(iterations=iterations, method=method)
        res = solve_bands(np.asarray(target_img), np.asarray(feature_img), opts, out_dtype=np.float64, out_layout="hwb")
        for b in range(target_img.shape[2]):
            pred_img[:, :, b] = res[:, :, b]
        return pred_img
    for b in range(target_img.shape[2]):
        MRP = WP_SMRP(target_img[:, :, b], feature_img[:, :, b], max_gamma=max_gamma, min_gamma=min_gamma)
        pred_img[:, :, b] = MRP.run(iterations=iterations, method=method)
    return pred_img


def apply_cloud_mask(target_img, mask, threshold=None):
    """Blank every band of the pixels the mask flags (VPint/utils/EO_wrapper.py:23-37), vectorised: NaN in
    ``mask[:, :, 0]`` when ``threshold`` is None, ``mask[:, :, 0] > threshold`` otherwise."""
    cloud_img = target_img.copy()
    m = np.asarray(mask)[:, :, 0]
    flagged = np.isnan(m) if threshold is None else (m > threshold)
    cloud_img[flagged, :] = np.nan
    return cloud_img
