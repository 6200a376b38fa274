"""Feature-weighted interpolators: drop-in ``WP_SMRP`` / ``WP_STMRP``.

Mirrors the public surface of VPint/WP_MRP.py for the propagation path: the
constructors, every ``run`` keyword and its default, the returned arrays and the
side effects on the object.  Weight precomputation, the Jacobi loop and the
stopping test run on the B200 through libvpint_b200.so; options that are not on
the device path yet raise :class:`VPintError` instead of being ignored.
"""

from __future__ import annotations

import numpy as np

from .base import SMRP, STMRP
from .errors import VPintError
from .solve import propagate

SPATIAL_CAP = 5000     # VPint/WP_MRP.py:112-115
TEMPORAL_CAP = 10000   # VPint/WP_MRP.py:1300-1303

_DEVICE_METHODS = ("exact", "cosine_similarity", "exact_inverse")


def wp_run_options(iterations=-1, method='exact', auto_terminate=True, auto_terminate_threshold=1e-4,
                   track_delta=False,
                   confidence=False, confidence_model=None,
                   save_gif=False, gif_path="convergence.gif", gif_ms_per_frame=100, store_gif_source=False,
                   gif_size_up=20, gif_clip_iter=100,
                   auto_adapt=False, auto_adaptation_epochs=100, auto_adaptation_proportion=0.5,
                   auto_adaptation_strategy='random', auto_adaptation_max_iter=-1,
                   auto_adaptation_subsample_strategy='max_contrast',
                   auto_adaptation_verbose=False,
                   prioritise_identity=False, priority_intensity=1, known_value_bias=0,
                   resistance=False, epsilon=0.01, mu=1.0,
                   clip_val=np.inf):
    """Normalise the keyword arguments of ``WP_SMRP.run`` (same names and defaults as
    VPint/WP_MRP.py:82-91; an unknown keyword raises TypeError just as forwarding it to
    the reference's ``run`` would) and reject what the device path cannot do."""
    if iterations > -1:          # VPint/WP_MRP.py:112-115
        auto_terminate = False
    else:
        iterations = SPATIAL_CAP
    if save_gif or store_gif_source:
        raise VPintError("save_gif / store_gif_source record every sweep on the host and are not supported "
                         "by the device path")
    if method not in _DEVICE_METHODS:
        if method == 'predict':
            raise VPintError("method='predict' needs a host-side sklearn model and is not supported by the "
                             "device path; use 'exact', 'cosine_similarity' or 'exact_inverse'")
        raise VPintError("Invalid weight prediction method: " + str(method))
    return dict(iterations=iterations, method=method, auto_terminate=auto_terminate,
                threshold=auto_terminate_threshold, track_delta=track_delta, clip_val=clip_val,
                prioritise_identity=bool(prioritise_identity), priority_intensity=priority_intensity,
                known_value_bias=known_value_bias, resistance=bool(resistance), epsilon=epsilon, mu=mu,
                auto_adapt=bool(auto_adapt), auto_adaptation_epochs=auto_adaptation_epochs,
                auto_adaptation_proportion=auto_adaptation_proportion, auto_adaptation_strategy=auto_adaptation_strategy,
                auto_adaptation_max_iter=auto_adaptation_max_iter,
                auto_adaptation_subsample_strategy=auto_adaptation_subsample_strategy,
                auto_adaptation_verbose=auto_adaptation_verbose)


class WP_SMRP(SMRP):
    """2-D value propagation with per-direction weights derived from a feature
    grid (reference: VPint/WP_MRP.py:13-393)."""

    def __init__(self, grid, feature_grid, model=None, init_strategy='mean', max_gamma=np.inf, min_gamma=0,
                 mask=None, clip_val=np.inf):
        # shape rules of VPint/WP_MRP.py:55-79
        if grid.shape[0] != feature_grid.shape[0] or grid.shape[1] != feature_grid.shape[1]:
            raise VPintError("Target and feature grids have different shapes: " + str(grid.shape) + " and "
                             + str(feature_grid.shape))
        if len(grid.shape) > 2:
            if grid.shape[2] == 1:
                grid = grid[:, :, 0]
            else:
                raise VPintError("Input grid needs to be two-dimensional, got " + str(grid.shape))
        super().__init__(grid, init_strategy=init_strategy, mask=mask)
        if len(feature_grid.shape) == 3:
            self.feature_grid = feature_grid.copy().astype(float)
        elif len(feature_grid.shape) == 2:
            self.feature_grid = feature_grid.copy().astype(float).reshape(
                (feature_grid.shape[0], feature_grid.shape[1], 1))
        else:
            raise VPintError("Improper feature dimensions; expected 2 or 3, got " + str(len(feature_grid.shape)))
        self.model = model
        self.max_gamma = max_gamma
        self.min_gamma = min_gamma
        self.clip_val = clip_val
        self._run_state = False
        self._run_method = "predict"

    def run(self, iterations=-1, method='exact', auto_terminate=True, auto_terminate_threshold=1e-4,
            track_delta=False,
            confidence=False, confidence_model=None,
            save_gif=False, gif_path="convergence.gif", gif_ms_per_frame=100, store_gif_source=False,
            gif_size_up=20, gif_clip_iter=100,
            auto_adapt=False, auto_adaptation_epochs=100, auto_adaptation_proportion=0.5,
            auto_adaptation_strategy='random', auto_adaptation_max_iter=-1,
            auto_adaptation_subsample_strategy='max_contrast',
            auto_adaptation_verbose=False,
            prioritise_identity=False, priority_intensity=1, known_value_bias=0,
            resistance=False, epsilon=0.01, mu=1.0,
            clip_val=np.inf):
        """Same contract as VPint/WP_MRP.py:82-393.

        Device path: ``method`` in {'exact', 'cosine_similarity', 'exact_inverse'},
        ``iterations``, ``auto_terminate``(+threshold), ``track_delta``, ``clip_val``.
        ``confidence`` / ``confidence_model`` are accepted and unused, as in the
        reference.  ``prioritise_identity`` (+ ``priority_intensity``, ``known_value_bias``) and
        ``resistance`` (+ ``epsilon``, ``mu``) run on the device as well.  Not available here
        (VPintError): ``method='predict'`` (needs the caller's sklearn model on the host) and GIF
        recording.  ``auto_adapt`` runs the reference's host-side search with every trial on the device.
        """
        opts = wp_run_options(iterations=iterations, method=method, auto_terminate=auto_terminate,
                              auto_terminate_threshold=auto_terminate_threshold, track_delta=track_delta,
                              save_gif=save_gif, store_gif_source=store_gif_source,
                              prioritise_identity=prioritise_identity, priority_intensity=priority_intensity,
                              known_value_bias=known_value_bias, resistance=resistance, epsilon=epsilon, mu=mu,
                              clip_val=clip_val)
        iterations, auto_terminate = opts["iterations"], opts["auto_terminate"]
        self.clip_val = clip_val
        if method == 'exact':
            # the reference replaces zeros in self.feature_grid in place (VPint/WP_MRP.py:143-145)
            flat = self.feature_grid.reshape(self.feature_grid.size)
            flat[flat == 0] = 0.01
        if auto_adapt:
            # host-side search (VPint/WP_MRP.py:212-239): every trial is a device solve; any failure is
            # swallowed and retried three times, then the given parameters are used -- as in the reference
            params = []
            if prioritise_identity:
                params.append('beta')
            if resistance:
                params.append('epsilon')
                params.append('mu')
            fail_counter = 0
            while True:
                if fail_counter >= 3:
                    print("Auto-adaptation failed 3 times; proceeding with default parameters")
                    break
                try:
                    params_opt = self.auto_adapt(params, auto_adaptation_epochs, auto_adaptation_proportion,
                                                 search_strategy=auto_adaptation_strategy,
                                                 max_sub_iter=auto_adaptation_max_iter,
                                                 subsample_strategy=auto_adaptation_subsample_strategy)
                    if auto_adaptation_verbose:
                        print("Best found params: " + str(params_opt))
                    for k, v in params_opt.items():
                        if k == 'beta':
                            priority_intensity = params_opt[k]
                        elif k == 'epsilon':
                            epsilon = params_opt[k]
                        elif k == 'mu':
                            mu = params_opt[k]
                    break
                except Exception:
                    fail_counter += 1
        priority = None
        if prioritise_identity:
            bias = None
            if known_value_bias > 0:
                # four identical SD_SMRP sub-solves of the known-cell indicator in the reference
                # (VPint/WP_MRP.py:263-272): solved once here
                from .sd import SD_SMRP
                ind = self.original_grid.copy()
                ind[~np.isnan(ind)] = 1
                bias = SD_SMRP(ind, init_strategy='zero', gamma=(1 - known_value_bias)).run()
            priority = dict(beta=priority_intensity, bias=bias, want_planes=True)
        delta_vec = np.zeros(iterations) if track_delta else None
        if iterations > 0 or priority is not None:
            # iterations == 0 with prioritise_identity: the reference still builds self.priority_grid before its
            # (empty) loop (VPint/WP_MRP.py:260), so the weights and the planes are formed and no sweep runs
            info = {}
            new_grid, deltas, sweeps = propagate(
                self.original_grid, self.pred_grid, planes=True, features=self.feature_grid, method=method,
                clamp=(method != 'exact'), min_gamma=self.min_gamma, max_gamma=self.max_gamma,
                max_iter=iterations, auto_terminate=auto_terminate, threshold=auto_terminate_threshold,
                clip_val=clip_val, track_delta=track_delta, priority=priority,
                resistance=(epsilon, mu) if resistance else None, info=info)
            if iterations > 0:
                self.pred_grid = new_grid
            if priority is not None:
                self.priority_grid = info.get("priority_planes")     # debugging access, as in the reference (:260)
            if track_delta and iterations > 0:
                delta_vec[:sweeps] = deltas
                if auto_terminate:
                    delta_vec = delta_vec[0:sweeps]
        self.run_state = True
        self.run_method = method
        if track_delta:
            return self.pred_grid, delta_vec
        return self.pred_grid

    def auto_adapt(self, params, search_epochs, subsample_proportion, search_strategy='random',
                   subsample_strategy='max_contrast', ranges={}, max_sub_iter=-1, hill_climbing_threshold=5):
        """Parameter search on a sub-sampled copy of the grid (VPint/WP_MRP.py:429-759); see
        :mod:`vpint_b200.adapt`."""
        from .adapt import auto_adapt
        return auto_adapt(self, params, search_epochs, subsample_proportion, search_strategy=search_strategy,
                          subsample_strategy=subsample_strategy, ranges=ranges, max_sub_iter=max_sub_iter,
                          hill_climbing_threshold=hill_climbing_threshold)

    def contrast_map(self, grid):
        """VPint/WP_MRP.py:893-940."""
        from .adapt import contrast_map
        return contrast_map(grid)

    def predict_weight(self, f1, f2, method):
        """Host utility with the reference's semantics (VPint/WP_MRP.py:396-427); the
        device computes the same quantity for whole grids."""
        if method == "predict":
            f = np.concatenate((f1, f2)).reshape(1, -1)
            gamma = 1.0
            try:
                gamma = self.model.predict(f)[0]
            except Exception:
                gamma = 1.0
        elif method == "cosine_similarity":
            gamma = np.dot(f1, f2) / max(np.sum(f1) * np.sum(f2), 0.01)
        elif method == "exact":
            d = f1.copy()
            d[d == 0] = 0.01
            gamma = np.mean(f2 / d)
        elif method == "exact_inverse":
            d = f2.copy()
            d[d == 0] = 0.01
            gamma = np.mean(f1 / d)
        else:
            raise VPintError("Invalid weight prediction method: " + str(method))
        return max(self.min_gamma, min(gamma, self.max_gamma))


class WP_STMRP(STMRP):
    """(H, W, T) value propagation with feature-derived weights; features are an
    (H, W, F) grid, static in time (reference: VPint/WP_MRP.py:1254-1468)."""

    def __init__(self, grid, feature_grid, model_spatial=None, model_temporal=None, auto_timesteps=False,
                 max_gamma=np.inf, min_gamma=0):
        super(WP_STMRP, self).__init__(grid, auto_timesteps)
        self.feature_grid = feature_grid.copy().astype(float)
        self.model_spatial = model_spatial
        self.model_temporal = model_temporal
        self.max_gamma = max_gamma
        self.min_gamma = min_gamma

    def run(self, iterations=-1, method='predict', auto_terminate=True, auto_terminate_threshold=1e-4,
            track_delta=False):
        """Same contract as VPint/WP_MRP.py:1289-1440 for ``method`` in
        {'exact', 'cosine_similarity'}.  The default ``method='predict'`` needs the
        caller's sklearn models on the host and raises VPintError."""
        if iterations > -1:
            auto_terminate = False
        else:
            iterations = TEMPORAL_CAP
        if method == 'predict':
            raise VPintError("method='predict' needs host-side sklearn models and is not supported by the device "
                             "path; use 'exact' or 'cosine_similarity'")
        if method not in ('exact', 'cosine_similarity'):
            raise VPintError("Invalid method: " + str(method))
        if self.feature_grid.ndim != 3:
            raise VPintError("Improper feature dimensions; expected (H, W, F), got " + str(self.feature_grid.shape))
        delta_vec = np.zeros(iterations) if track_delta else None
        if iterations > 0:
            new_grid, deltas, sweeps = propagate(
                self.original_grid, self.pred_grid, planes=True, features=self.feature_grid, method=method,
                max_iter=iterations, auto_terminate=auto_terminate, threshold=auto_terminate_threshold,
                track_delta=track_delta)
            self.pred_grid = new_grid
            if track_delta:
                delta_vec[:sweeps] = deltas
                if auto_terminate:
                    delta_vec = delta_vec[0:sweeps]
        if track_delta:
            return self.pred_grid, delta_vec
        return self.pred_grid
