"""Host-side parameter search of ``WP_SMRP.run(auto_adapt=True)``.

Mirrors ``WP_SMRP.auto_adapt`` / ``contrast_map`` (VPint/WP_MRP.py:429-759, 893-940): hide part of the known
cells, run trial interpolations with candidate (beta, epsilon, mu) and keep the candidates with the lowest
mean absolute error.  The search itself is host logic; every trial ``run()`` goes through the device path.
The global NumPy RNG is consumed in the reference's order, and the reference's control-flow quirks are kept
on purpose (a drop-in must pick the same parameters):

* the 'sequential' strategy always ends in a RuntimeError (its result dict is edited while it is being
  iterated, :751-756), which ``run()`` swallows and retries three times before using the defaults;
* under 'hill_climbing' the "best" dict is the SAME object as the dict that keeps being modified (:728);
* when no trial beats ``inf`` (all-NaN errors) the result is never bound and the lookup fails (:751).
"""

from __future__ import annotations

import numpy as np

from .errors import VPintError


def contrast_map(grid):
    """Mean absolute difference between a cell and its 4 neighbours, scaled to [0, 1]
    (VPint/WP_MRP.py:893-940).  The reference fills the neighbour planes with slices that are one element
    short (``up[1:-1] = grid[0:-2]`` ...), so the last row / column of every plane stays 0; kept."""
    H, W = grid.shape[0], grid.shape[1]
    count = np.full(grid.shape, 4.0)
    count[:, 0] -= 1.0
    count[:, W - 1] -= 1.0
    count[0, :] -= 1.0
    count[H - 1, :] -= 1.0
    nb = np.zeros((H, W, 4))
    nb[1:-1, :, 0] = grid[0:-2, :]
    nb[:, 0:-2, 1] = grid[:, 1:-1]
    nb[0:-2, :, 2] = grid[1:-1, :]
    nb[:, 1:-1, 3] = grid[:, 0:-2]
    diff = np.absolute(nb - np.repeat(grid[:, :, np.newaxis], 4, axis=2))
    avg = np.nansum(diff, axis=-1) / count
    lo, hi = np.nanmin(avg), np.nanmax(avg)
    return np.clip((avg - lo) / (hi - lo), 0, 1)


def _subsample(interp, strategy, proportion):
    sub = interp.original_grid.copy()
    flat = sub.reshape(sub.size)
    if strategy == 'random':
        flat[np.random.rand(sub.size) < proportion] = np.nan
        return flat.reshape(sub.shape)
    if strategy == 'max_diff':
        score = np.absolute(interp.original_grid - np.mean(interp.feature_grid, axis=2))
    elif strategy == 'max_contrast':
        score = contrast_map(sub)
    else:
        raise VPintError("Invalid subsample strategy: " + str(strategy))
    n_hide = int(proportion * len(flat[~np.isnan(flat)]))
    ranked = np.nan_to_num(score.reshape(sub.size), nan=-1.0)
    flat[np.argpartition(-ranked, n_hide)[:n_hide]] = np.nan      # the n_hide highest scores
    return flat.reshape(sub.shape)


def auto_adapt(interp, params, search_epochs, subsample_proportion, search_strategy='random',
               subsample_strategy='max_contrast', ranges={}, max_sub_iter=-1, hill_climbing_threshold=5):
    """Returns {'beta': ..., 'epsilon': ..., 'mu': ...} restricted to ``params``."""
    from .wp import WP_SMRP
    original = interp.original_grid
    sub_grid = _subsample(interp, subsample_strategy, subsample_proportion)
    bounds = {'beta': (0, 4), 'epsilon': (0.1, 0.01), 'mu': (np.nanmean(original), 3 * np.nanmean(original))}
    for k, v in ranges:
        bounds[k] = v
    best_loss = np.inf
    best_val = {}
    for k in params:
        if k not in ['beta', 'epsilon', 'mu']:
            raise VPintError("Invalid parameter to optimise: " + str(k))
        best_val[k] = -1
    default_mu = lambda: np.nanmean(original) + 2 * np.nanstd(original)

    def trial(vals, **forced):
        """One interpolation of the sub-sampled grid with the candidate values; mean absolute error over
        the cells the original grid knows."""
        scratch = WP_SMRP(sub_grid, interp.feature_grid.copy())
        if forced:
            pred = scratch.run(auto_adapt=False, **forced)
        elif 'beta' in vals and 'epsilon' in vals:
            pred = scratch.run(prioritise_identity=True, priority_intensity=vals['beta'], iterations=max_sub_iter,
                               auto_adapt=False, resistance=True, epsilon=vals['epsilon'],
                               mu=vals['mu'] if 'mu' in vals else default_mu())
        elif 'beta' in vals:
            pred = scratch.run(prioritise_identity=True, priority_intensity=vals['beta'], iterations=max_sub_iter,
                               auto_adapt=False)
        elif 'epsilon' in vals:
            pred = scratch.run(resistance=True, epsilon=vals['epsilon'],
                               mu=vals['mu'] if 'mu' in vals else default_mu(), auto_adapt=False)
        return np.nanmean(np.absolute(pred.reshape(pred.size) - original.reshape(original.size)))

    def defaults(eps):
        out = {}
        for k in params:
            if k == 'beta':
                out[k] = 1
            if k == 'epsilon':
                out[k] = eps
            if k == 'mu':
                out[k] = default_mu()
        return out

    if search_strategy == 'random':
        # The candidates of this strategy do not depend on one another and a trial draws nothing from the RNG, so all
        # of them are drawn first -- same calls, same order as the reference's epoch loop -- and run as ONE batched
        # solve (problem = candidate); the pick below walks the errors in epoch order like the reference.
        drawn = []
        for ep in range(search_epochs):
            if ep == 0:
                vals = defaults(0)
            else:
                vals = {}
                for k in params:
                    if k == 'beta':
                        vals[k] = np.random.randint(low=bounds[k][0], high=bounds[k][1])
                    if k == 'epsilon':
                        vals[k] = min(max(0, np.random.uniform(low=0, high=0.3)), 1)
                    if k == 'mu':
                        vals[k] = np.random.uniform(low=bounds[k][0], high=bounds[k][1])
            drawn.append(vals)
        if params and ('beta' in params or 'epsilon' in params):
            from .batch import wp_trials
            cand = [dict(v, mu=v['mu'] if 'mu' in v else default_mu()) if 'epsilon' in v else dict(v) for v in drawn]
            maes = wp_trials(sub_grid, interp.feature_grid, original, cand, max_sub_iter=max_sub_iter)
        else:
            maes = [trial(v) for v in drawn]
        for vals, mae in zip(drawn, maes):
            if mae < best_loss:
                best_vals, best_loss = vals, mae

    elif search_strategy == 'sequential':
        if 'beta' not in params and 'epsilon' not in params:
            raise VPintError("Cannot use sequential auto-adaptation without both beta and epsilon/mu")
        best_vals = {}
        for val in range(0, 8):                                 # grid search for beta
            mae = trial(None, prioritise_identity=True, priority_intensity=val, iterations=max_sub_iter)
            if mae < best_loss:
                best_val, best_loss = val, mae
        best_beta = best_val
        params.pop(params.index('beta'))
        for ep in range(search_epochs):                         # random search for epsilon / mu, same best_loss
            if ep == 0:
                vals = defaults(0)
            else:
                vals = {}
                for k in params:
                    if k == 'epsilon':
                        vals[k] = min(max(0, np.random.uniform(low=0, high=0.5)), 1)
                    if k == 'mu':
                        vals[k] = np.random.uniform(low=bounds[k][0], high=bounds[k][1])
            mae = trial(None, resistance=True, epsilon=vals['epsilon'], mu=vals['mu'] if 'mu' in vals else default_mu(),
                        prioritise_identity=True, priority_intensity=best_beta)
            if mae < best_loss:
                best_vals, best_loss = vals, mae
        best_vals['beta'] = best_beta

    elif search_strategy == 'hill_climbing':
        explored_default = False
        stale = 0
        for ep in range(search_epochs):
            if not explored_default:
                vals = defaults(0)
            else:
                for k in params:                                # one random step per parameter, in place
                    if k == 'beta':
                        r = np.random.rand()
                        if r <= 0.33:
                            vals[k] = max(bounds[k][0], vals[k] - 1)
                        elif r > 0.33 and r <= 0.66:
                            vals[k] = min(bounds[k][1], vals[k] + 1)
                    if k == 'epsilon':
                        r = np.random.rand()
                        if r <= 0.33:
                            vals[k] = max(0, vals[k] - bounds[k][1])
                        elif r > 0.33 and r <= 0.66:
                            vals[k] = min(1, vals[k] + bounds[k][1])
                    if k == 'mu':
                        factor = np.random.uniform(low=0, high=np.nanstd(original))
                        if r <= 0.33:                           # the draw of the previous parameter decides
                            vals[k] = max(bounds[k][0], vals[k] - factor)
                        elif r > 0.33 and r <= 0.66:
                            vals[k] = min(bounds[k][1], vals[k] + factor)
            mae = trial(vals)
            if mae < best_loss:
                best_vals, best_loss = vals, mae                # same object as vals: later steps change it too
                stale = 0
            else:
                stale += 1
                if stale >= hill_climbing_threshold:
                    break
            if not explored_default:
                vals = defaults(0.1)
                explored_default = True
    else:
        raise VPintError("Invalid search strategy: " + str(search_strategy))

    for k, v in best_vals.items():
        if k in params and v == -1:
            print("WARNING: no " + k + " better than dummy, please check your code (defaulting to 1)")
            best_vals[k] = 1
        if not (k in params):
            best_vals.pop(k)
    return best_vals
