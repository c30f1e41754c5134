"""Validation-metric kernels (dmvae_waypoint_speeds, dmvae_histogram, dmvae_trajectories_per_cell)
against the reference's own outputs (tests/golden/metrics.npz, produced by running Distribution.py / Spatial_Distribution.py)
and against the CPU oracle on seeded inputs.

Tolerances: cell counts and histogram counts are integers and must be exact; a waypoint speed may differ from the
reference's by one unit in the last place (NumPy squares float32 scalars through powf, which is not correctly rounded,
where the kernel multiplies): 3e-7 relative; the Jensen-Shannon divergence and the RMSE follow from counts: 1e-9 when the
counts agree."""
import os

import numpy as np
import pytest
import torch

from oracle import metrics_oracle as MO

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "metrics.npz"))


@pytest.mark.parametrize("tag", ["sce1", "sce4"])
def test_metrics_reproduce_the_reference_outputs(gold, tag):
    from dmvae import validation as V
    gen, hum, name = gold[f"{tag}/gen"], gold[f"{tag}/hum"], str(gold[f"{tag}/model_name"])
    # speeds: the golden trajectories are in the reference's [x, y, t] order and contain repeated time stamps at the
    # start of the array, inside, at the first step and at the last step of a trajectory
    vg, (lo, hi) = V.waypoint_speeds(gen, layout="xyt")
    want = gold[f"{tag}/v_gen"]
    got = vg.cpu().numpy().astype(np.float64)
    assert got.shape == want.shape == (gen.shape[0] * gen.shape[1],)
    assert np.abs(got - want).max() <= 3e-7 * np.abs(want).max()
    assert np.all(got[want == 0.0] == 0.0)
    assert abs(lo - want.min()) <= 3e-7 * want.max() and abs(hi - want.max()) <= 3e-7 * want.max()
    # the same trajectories in this library's own [t, x, y] order
    vg2, _ = V.waypoint_speeds(np.ascontiguousarray(gen[:, :, [2, 0, 1]]), layout="txy")
    assert torch.equal(vg2, vg)
    # histogram over the reference's 50 edges: np.histogram of the SAME values, exactly
    edges = gold[f"{tag}/bins_js"]
    hg = V.histogram(vg, edges)
    np.testing.assert_array_equal(hg, np.histogram(got, bins=edges)[0])
    assert np.abs(hg - gold[f"{tag}/hist_gen"]).sum() <= 2          # a speed one ulp off may change bins at an edge
    # the divergence end to end, and from the reference's own counts
    js = V.velocity_js_divergence(gen, hum, layout="xyt", human_layout="xyt")
    assert abs(js - float(gold[f"{tag}/js"])) <= 2e-4
    assert abs(V.js_from_counts(gold[f"{tag}/hist_gen"], gold[f"{tag}/hist_hum"]) - float(gold[f"{tag}/js"])) <= 1e-12
    # trajectories per cell (points on cell borders, points outside the grid) and the RMSE: exact
    Hg = V.trajectories_per_cell(gen, name, 1.0, layout="xyt")
    Hh = V.trajectories_per_cell(hum, name, 1.0, layout="xyt")
    np.testing.assert_array_equal(Hg, gold[f"{tag}/H_gen"])
    np.testing.assert_array_equal(Hh, gold[f"{tag}/H_hum"])
    assert abs(V.rmse_frequency(Hg, Hh) - float(gold[f"{tag}/rmse"])) <= 1e-12


@pytest.mark.parametrize("n,T,grid", [(1, 2, 1.0), (257, 10, 1.0), (5000, 10, 0.5), (300, 43, 2.0), (64, 400, 1.0)])
def test_metrics_vs_oracle_on_seeded_trajectories(n, T, grid):
    from dmvae import validation as V
    rng = np.random.default_rng(n * 1000 + T)
    t = np.cumsum(rng.uniform(0.0, 1.5, size=(n, T)), axis=1)
    t[rng.random((n, T)) < 0.05] = 0.0                                  # plenty of non-increasing time stamps
    xy = np.cumsum(rng.normal(0, 1.0, size=(n, T, 2)), axis=1) + rng.uniform([-5, -30], [25, 110], size=(n, 1, 2))
    traj_txy = np.concatenate([t[..., None], xy], -1).astype(np.float32)          # this library's order
    traj_xyt = np.ascontiguousarray(traj_txy[:, :, [1, 2, 0]])
    want = MO.waypoint_velocities(list(traj_xyt))
    got, (lo, hi) = V.waypoint_speeds(traj_txy)
    got = got.cpu().numpy().astype(np.float64)
    scale = max(np.abs(want).max(), 1e-30)
    assert np.abs(got - want).max() <= 3e-7 * scale
    assert lo == got.min() and hi == got.max()
    edges = np.linspace(got.min(), got.max(), 50) if got.max() > got.min() else np.array([0.0, 1.0])
    np.testing.assert_array_equal(V.histogram(torch.from_numpy(got.astype(np.float32)), edges), np.histogram(got, bins=edges)[0])
    for name in ("vae_offset_sce4_x", "vae_offset_sce2_x"):
        np.testing.assert_array_equal(V.trajectories_per_cell(traj_txy, name, grid), MO.trajectories_per_cell(list(traj_xyt), name, grid))


def test_metrics_on_a_million_generated_trajectories(golden_dir):
    """Size-independent properties at BASELINE configs[2] scale (2^20 decoded trajectories of a shipped checkpoint):
    every speed is counted exactly once in the histogram, a cell never counts more trajectories than exist, every
    trajectory is counted in at least one cell, and the counts of two halves add up to the counts of the whole."""
    from dmvae import ConditionalTrajectoryVAE
    from dmvae import validation as V
    ck = np.load(os.path.join(golden_dir, "ckpt_sce4_cond.npz"))
    model = ConditionalTrajectoryVAE(10, 3, 8)
    model.load_state_dict({k: torch.from_numpy(ck[k]) for k in ck.files})
    n = 1 << 20
    traj = model.to("cuda").eval().generate(torch.tensor([[11.0, 0.0]]), n=n, seed=3)
    v, (lo, hi) = V.waypoint_speeds(traj)
    assert v.numel() == n * 10 and bool(torch.isfinite(v).all()) and lo == float(v.min()) and hi == float(v.max())
    edges = np.linspace(lo, hi, 50)
    h = V.histogram(v, edges)
    assert h.sum() == n * 10
    H = V.trajectories_per_cell(traj, "vae_offset_sce4_cond", 1.0)
    assert H.max() <= n and H.sum() >= n
    Ha = V.trajectories_per_cell(traj[: n // 2], "vae_offset_sce4_cond", 1.0)
    Hb = V.trajectories_per_cell(traj[n // 2:], "vae_offset_sce4_cond", 1.0)
    np.testing.assert_array_equal(Ha + Hb, H)
    ha = V.histogram(v[: n * 5], edges)
    hb = V.histogram(v[n * 5:], edges)
    np.testing.assert_array_equal(ha + hb, h)
    js_self = V.js_from_counts(h, h)
    assert abs(js_self) < 1e-12                                   # a distribution against itself
