"""Golden vectors for the validation metrics, produced by RUNNING THE REFERENCE'S OWN FUNCTIONS.

TEST INFRASTRUCTURE.  Run once in the build container (where ``/root/reference`` is mounted):

    python -m oracle.make_golden_metrics

tests/golden/metrics.npz: seeded float32 waypoint trajectories ``[x, y, t]`` for two scenario grids (including
repeated time stamps, points outside the grid and points exactly on cell borders) with the outputs of
``Distribution.calculate_human_velocities``, the Jensen-Shannon value computed inside
``Distribution.plot_velocity_distribution`` (taken from that function's frame when the stubbed matplotlib stops it),
``Spatial_Distribution._count_trajectories_per_grid`` and ``calculate_rmse_frequency_new``.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)

from oracle.make_golden_entry import frame_locals_of  # noqa: E402
from oracle.ref_loader import REFERENCE_ROOT, _stub_matplotlib, load_reference  # noqa: E402


def load_metric_modules():
    """Distribution.py and Spatial_Distribution.py of the reference (they import matplotlib pieces and Tools)."""
    ref = load_reference()
    _stub_matplotlib()
    for extra in ("matplotlib.lines", "matplotlib.collections", "matplotlib.ticker", "mpl_toolkits", "mpl_toolkits.mplot3d"):
        if extra not in sys.modules:
            m = types.ModuleType(extra)
            m.Axes3D = object
            sys.modules[extra] = m
    sys.modules["matplotlib.colors"].LinearSegmentedColormap = object
    saved = {n: sys.modules.get(n) for n in ("Tools", "Training_VAE")}
    sys.modules["Tools"], sys.modules["Training_VAE"] = ref.Tools, ref.Training_VAE
    sys.dont_write_bytecode = True
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        import Distribution
        import Spatial_Distribution
    finally:
        sys.path.remove(REFERENCE_ROOT)
        for n, m in saved.items():
            if m is None:
                sys.modules.pop(n, None)
            else:
                sys.modules[n] = m
        for n in ("Distribution", "Spatial_Distribution", "MPC", "MPC.MPC_Tracking"):
            sys.modules.pop(n, None)
    return Distribution, Spatial_Distribution


def synth(rng, n, T, x_box, y_box, axis):
    """float32 [x, y, t] waypoints that start in a box and move along `axis`, some with repeated time stamps."""
    t = np.cumsum(rng.uniform(0.3, 2.0, size=(n, T)), axis=1)
    t -= t[:, :1]
    speed = rng.uniform(2.0, 15.0, size=(n, 1))
    along = speed * t + rng.normal(0, 0.3, size=(n, T))
    lat = np.cumsum(rng.normal(0, 0.2, size=(n, T)), axis=1)
    x0 = rng.uniform(*x_box, size=(n, 1))
    y0 = rng.uniform(*y_box, size=(n, 1))
    x = x0 + (along if axis == 0 else lat)
    y = y0 + (along if axis == 1 else lat)
    traj = np.stack([x, y, t], -1).astype(np.float32)
    traj[1, 3, 2] = traj[1, 2, 2]            # a zero time step inside a trajectory
    traj[4, 1, 2] = traj[4, 0, 2]            # ... at its very first step: takes the previous trajectory's last speed
    traj[4, 2, 2] = traj[4, 0, 2]
    traj[7, -1, 2] = traj[7, -2, 2]          # ... at its last step (and so at the repeated last point)
    traj[0, 0, 2] = traj[0, 1, 2] = 0.0      # ... at the very beginning of the whole array: 0.0
    return traj


def main() -> None:
    Distribution, SD = load_metric_modules()
    rng = np.random.default_rng(2025)
    out = {}
    for tag, model_name, xb, yb, axis in (("sce1", "vae_offset_sce1_cond_ld8_epoch3000.pth", (-196.0, -190.0), (42.0, 50.0), 1),
                                          ("sce4", "vae_offset_sce4_cond_ld8_epoch3000.pth", (2.0, 18.0), (-25.0, 60.0), 1)):
        gen = synth(rng, 300, 10, xb, yb, axis)
        hum = synth(rng, 40, 10, xb, yb, axis)
        gen[10, 5, 0] = np.float32(np.floor(gen[10, 5, 0]))        # exactly on a cell border
        gen[11, 5, 1] = np.float32(np.floor(gen[11, 5, 1]))
        gen[12, :, 0] += 100.0                                       # outside the grid: clipped into the border cells
        out[f"{tag}/model_name"] = np.array(model_name)
        out[f"{tag}/gen"], out[f"{tag}/hum"] = gen, hum
        vg = Distribution.calculate_human_velocities(list(gen))
        vh = Distribution.calculate_human_velocities(list(hum))
        out[f"{tag}/v_gen"], out[f"{tag}/v_hum"] = vg, vh
        try:
            Distribution.plot_velocity_distribution(vg, vh, save_path=None)
        except Exception as e:  # noqa: BLE001 - stubbed matplotlib ends the function after the divergence is computed
            loc = frame_locals_of(e, "plot_velocity_distribution")
        else:
            raise AssertionError("expected the plotting tail to stop")
        out[f"{tag}/js"] = np.array(loc["js_divergence"])
        out[f"{tag}/bins_js"] = np.asarray(loc["bins_js"])
        hg, _ = np.histogram(vg, bins=loc["bins_js"])
        hh, _ = np.histogram(vh, bins=loc["bins_js"])
        out[f"{tag}/hist_gen"], out[f"{tag}/hist_hum"] = hg, hh
        Hs, xe, ye = SD._count_trajectories_per_grid(list(gen), model_name, 1.0)
        Ho, _, _ = SD._count_trajectories_per_grid(list(hum), model_name, 1.0)
        out[f"{tag}/H_gen"], out[f"{tag}/H_hum"], out[f"{tag}/x_edges"], out[f"{tag}/y_edges"] = Hs, Ho, xe, ye
        out[f"{tag}/rmse"] = np.array(SD.calculate_rmse_frequency_new(list(gen), list(hum), model_name, 1.0))
    np.savez(os.path.join(GOLD, "metrics.npz"), **out)
    print("wrote tests/golden/metrics.npz")


if __name__ == "__main__":
    main()
