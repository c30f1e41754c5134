"""CPU restatement of the reference's validation metrics.

TEST INFRASTRUCTURE - not part of the product path.  Only ``tests/``, ``oracle/make_golden_metrics.py`` and
``bench.py``'s CPU legs may import this module.  Pinned by ``tests/golden/metrics.npz``, which
``oracle/make_golden_metrics.py`` produced by running the reference's own functions
(``tests/test_metrics_oracle.py``).

What it restates, for waypoint trajectories ``(n, T, 3)`` in the reference's ``[x, y, t]`` column order:

* ``waypoint_velocities``     ``Distribution.calculate_human_velocities`` (``Distribution.py:248-296``): per
  trajectory the speed between consecutive points, the last point repeating the last speed; a step whose time
  difference is not above 1e-6 repeats the value appended before it (possibly the previous trajectory's), 0 at the
  very beginning.  float32 inputs stay float32 (NumPy scalar arithmetic), the result array is float64.
* ``js_divergence``           the Jensen-Shannon block of ``Distribution.plot_velocity_distribution``
  (``Distribution.py:309-331``): 50 common edges between the joint min and max, counts, base-2 divergence.
* ``trajectories_per_cell``   ``Spatial_Distribution._count_trajectories_per_grid`` (``Spatial_Distribution.py:387-431``)
  with the scenario grids of ``_get_grid_edges`` (``:362-384``): a cell counts the trajectories that visit it.
* ``rmse_frequency``          ``Spatial_Distribution.calculate_rmse_frequency_new`` (``:434-493``).
"""
from __future__ import annotations

import numpy as np


def waypoint_velocities(trajs_xyt) -> np.ndarray:
    velocities = []
    for traj in trajs_xyt:
        times, xs, ys = traj[:, 2], traj[:, 0], traj[:, 1]
        n = len(traj)
        steps = list(range(n - 1)) + ([n - 2] if n > 1 else [])      # the last point repeats the last step (:282-294)
        for i in steps:
            dt = times[i + 1] - times[i]
            if dt > 1e-6:
                dx, dy = xs[i + 1] - xs[i], ys[i + 1] - ys[i]
                velocities.append(np.sqrt(dx ** 2 + dy ** 2) / dt)
            else:
                velocities.append(velocities[-1] if velocities else 0.0)
    return np.array(velocities)


def js_edges(generated_velocities, human_velocities) -> np.ndarray:
    v_min = min(np.min(generated_velocities), np.min(human_velocities))
    v_max = max(np.max(generated_velocities), np.max(human_velocities))
    return np.linspace(v_min, v_max, 50)


def js_from_counts(hist_gen, hist_human) -> float:
    """Distribution.py:316-331 from the two count vectors; scipy.stats.entropy(p, q, base=2) = sum(p log2(p / q)) after
    normalising p and q."""
    p = hist_gen / (hist_gen.sum() + 1e-10)
    q = hist_human / (hist_human.sum() + 1e-10)
    m = 0.5 * (p + q)
    eps = 1e-10

    def kl(a, b):
        a = np.asarray(a, dtype=np.float64)
        b = np.asarray(b, dtype=np.float64)
        a = a / a.sum()
        b = b / b.sum()
        return float(np.sum(np.where(a > 0, a * np.log(a / b), 0.0)) / np.log(2.0))

    return 0.5 * (kl(p + eps, m + eps) + kl(q + eps, m + eps))


def js_divergence(generated_velocities, human_velocities) -> float:
    edges = js_edges(generated_velocities, human_velocities)
    hg, _ = np.histogram(generated_velocities, bins=edges)
    hh, _ = np.histogram(human_velocities, bins=edges)
    return js_from_counts(hg, hh)


def grid_edges(model_name: str, grid_size: float = 1.0):
    if "sce1" in model_name:
        return np.arange(-198, -188 + 1, grid_size), np.arange(40, 80 + 1, grid_size)
    if "sce2" in model_name:
        return np.arange(-200, -120, grid_size), np.arange(-8, 6, grid_size)
    if "sce3" in model_name:
        return np.arange(148, 158, grid_size), np.arange(-80, 22, grid_size)
    return np.arange(0, 20, grid_size), np.arange(-20, 100, grid_size)


def trajectories_per_cell(trajs_xy, model_name: str, grid_size: float = 1.0) -> np.ndarray:
    x_edges, y_edges = grid_edges(model_name, grid_size)
    H = np.zeros((len(y_edges) - 1, len(x_edges) - 1), dtype=np.int64)
    for traj in trajs_xy:
        xi = np.clip(np.digitize(traj[:, 0], x_edges) - 1, 0, len(x_edges) - 2)
        yi = np.clip(np.digitize(traj[:, 1], y_edges) - 1, 0, len(y_edges) - 2)
        for y_idx, x_idx in set(zip(yi.tolist(), xi.tolist())):
            H[y_idx, x_idx] += 1
    return H


def rmse_frequency(H_sim, H_obs) -> float:
    f_sim, f_obs = np.asarray(H_sim).flatten(), np.asarray(H_obs).flatten()
    mask = (f_sim > 0) | (f_obs > 0)
    if not mask.any():
        return 0.0
    return float(np.sqrt(np.mean((f_sim[mask] - f_obs[mask]) ** 2)))
