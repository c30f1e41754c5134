"""Validation metrics over bulk-generated waypoint trajectories.

The reference judges its generator by two distribution comparisons between generated and human trajectories:
the Jensen-Shannon divergence of the waypoint speeds (``Distribution.py:248-331``) and the RMSE between the maps of
"how many trajectories visit this grid cell" (``Spatial_Distribution.py:362-493``); its own numbers are in
``results/ModelValidation/JS_divergence.txt``.  The per-trajectory passes run on the GPU (``dmvae_waypoint_speeds``,
``dmvae_histogram``, ``dmvae_trajectories_per_cell``: three HBM-bound scans, sized for the 10^6 trajectories per scenario
that ``model.generate`` produces); what remains - 49 counts, one count map - is a handful of float64 operations here.

Trajectories are ``(n, T, 3)`` float32 tensors or arrays; ``layout="txy"`` is what ``model.generate`` /
``Tools.generate_trajectories`` return, ``"xyt"`` the order of the reference's tracker and human data.  There is no CPU
path: inputs are moved to the current CUDA device.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _lib
from ._lib import check, ptr, stream_ptr

_LAYOUT = {"txy": 0, "xyt": 1}


def _device_traj(traj):
    t = torch.as_tensor(traj)
    if t.dim() != 3 or t.shape[2] != 3 or t.shape[0] < 1:
        raise ValueError(f"expected (n, T, 3) trajectories, got {tuple(t.shape)}")
    if not torch.cuda.is_available():
        raise _lib.DmvaeError("no CUDA device is visible and dmvae has no CPU path")
    return t.detach().to(device=torch.device("cuda", torch.cuda.current_device()), dtype=torch.float32).contiguous()


def waypoint_speeds(traj, layout: str = "txy"):
    """``Distribution.calculate_human_velocities`` (``Distribution.py:248-296``) for all trajectories at once.
    Returns (speeds ``(n * T,)`` float32 device tensor, (min, max) as Python floats)."""
    t = _device_traj(traj)
    n, T = int(t.shape[0]), int(t.shape[1])
    speeds = torch.empty(n * T, dtype=torch.float32, device=t.device)
    minmax = torch.empty(2, dtype=torch.float32, device=t.device)
    with torch.cuda.device(t.device):
        check(_lib.lib().dmvae_waypoint_speeds(ptr(t), n, T, _LAYOUT[layout], ptr(speeds), ptr(minmax), stream_ptr()),
              "dmvae_waypoint_speeds")
    lo, hi = minmax.cpu().tolist()
    return speeds, (lo, hi)


def histogram(values: torch.Tensor, edges) -> np.ndarray:
    """``np.histogram(values, bins=edges)[0]`` for a float32 device tensor; ``edges``: increasing float64."""
    v = values.detach().to(dtype=torch.float32).contiguous()
    if not v.is_cuda:
        v = v.cuda()
    e = np.ascontiguousarray(np.asarray(edges, dtype=np.float64))
    nb = len(e) - 1
    counts = torch.empty(nb, dtype=torch.int64, device=v.device)
    with torch.cuda.device(v.device):
        check(_lib.lib().dmvae_histogram(ptr(v), v.numel(), e.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), nb, ptr(counts),
                                         stream_ptr()), "dmvae_histogram")
    return counts.cpu().numpy()


def js_from_counts(hist_gen, hist_human) -> float:
    """``Distribution.py:316-331``: base-2 Jensen-Shannon divergence of two count vectors over the same edges."""
    p = np.asarray(hist_gen, dtype=np.float64)
    q = np.asarray(hist_human, dtype=np.float64)
    p = p / (p.sum() + 1e-10)
    q = q / (q.sum() + 1e-10)
    m = 0.5 * (p + q)
    eps = 1e-10

    def kl(a, b):   # scipy.stats.entropy(a, b, base=2): both normalised, then sum(a log2(a / b))
        a, b = a / a.sum(), b / b.sum()
        return float(np.sum(np.where(a > 0, a * np.log(a / b), 0.0)) / np.log(2.0))

    return 0.5 * (kl(p + eps, m + eps) + kl(q + eps, m + eps))


def velocity_js_divergence(generated, human, layout: str = "txy", human_layout: str = "xyt") -> float:
    """Jensen-Shannon divergence between the waypoint speeds of generated and human trajectories
    (``Distribution.plot_velocity_distribution``, ``Distribution.py:309-331``): 50 common edges between the joint minimum
    and maximum, counts, base-2 divergence in [0, 1]."""
    vg, (g_lo, g_hi) = waypoint_speeds(generated, layout)
    vh, (h_lo, h_hi) = waypoint_speeds(human, human_layout)
    # min / max of float32 values, then np.linspace in float64 like the reference (its arrays hold float32 values)
    edges = np.linspace(min(np.float64(g_lo), np.float64(h_lo)), max(np.float64(g_hi), np.float64(h_hi)), 50)
    return js_from_counts(histogram(vg, edges), histogram(vh, edges))


def grid_edges(model_name: str, grid_size: float = 1.0):
    """Scenario grids of ``Spatial_Distribution._get_grid_edges`` (``Spatial_Distribution.py:362-384``) as
    (x0, number of x edges, y0, number of y edges): the edges are ``np.arange(start, stop, grid_size)``."""
    if "sce1" in model_name:
        xr, yr = (-198, -188 + 1), (40, 80 + 1)
    elif "sce2" in model_name:
        xr, yr = (-200, -120), (-8, 6)
    elif "sce3" in model_name:
        xr, yr = (148, 158), (-80, 22)
    else:
        xr, yr = (0, 20), (-20, 100)
    nx, ny = len(np.arange(xr[0], xr[1], grid_size)), len(np.arange(yr[0], yr[1], grid_size))
    return float(xr[0]), nx, float(yr[0]), ny


def trajectories_per_cell(traj, model_name: str, grid_size: float = 1.0, layout: str = "txy") -> np.ndarray:
    """``Spatial_Distribution._count_trajectories_per_grid`` (``Spatial_Distribution.py:387-431``): ``H[i, j]`` = number of
    trajectories with at least one point in y cell i, x cell j (int64, ``(ny - 1, nx - 1)``)."""
    t = _device_traj(traj)
    x0, nx, y0, ny = grid_edges(model_name, grid_size)
    counts = torch.empty((ny - 1) * (nx - 1), dtype=torch.int64, device=t.device)
    with torch.cuda.device(t.device):
        check(_lib.lib().dmvae_trajectories_per_cell(ptr(t), int(t.shape[0]), int(t.shape[1]), _LAYOUT[layout], x0, float(grid_size), nx,
                                                     y0, float(grid_size), ny, ptr(counts), stream_ptr()), "dmvae_trajectories_per_cell")
    return counts.cpu().numpy().reshape(ny - 1, nx - 1)


def rmse_frequency(H_sim, H_obs) -> float:
    """``Spatial_Distribution.calculate_rmse_frequency_new`` (``Spatial_Distribution.py:434-493``): RMSE over the cells
    that either map visits."""
    f_sim, f_obs = np.asarray(H_sim).flatten(), np.asarray(H_obs).flatten()
    mask = (f_sim > 0) | (f_obs > 0)
    if not mask.any():
        return 0.0
    return float(np.sqrt(np.mean((f_sim[mask].astype(np.float64) - f_obs[mask].astype(np.float64)) ** 2)))
