"""Generate -> track hand-off: host logic on CPU against the reference's own lines, the
batched decode on the GPU against the one-trajectory-per-call helper the reference loop uses."""
import os

import numpy as np
import pytest
import torch

from dmvae import handoff

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reorder_matches_the_reference_lines():
    rng = np.random.default_rng(0)
    traj = rng.normal(size=(7, 10, 3)).astype(np.float32)
    got = handoff.to_tracker_waypoints(traj)
    for i in range(7):                      # Distribution.py:77-78, verbatim
        waypoints = traj[i].copy()
        waypoints = waypoints[:, [1, 2, 0]]
        waypoints[0, 2] = 0.0
        np.testing.assert_array_equal(got[i], waypoints)
    assert got.dtype == np.float32 and traj[0, 0, 0] != 0.0      # input untouched
    with pytest.raises(ValueError):
        handoff.to_tracker_waypoints(np.zeros((10, 2)))


def test_time_guard_is_the_interpolators_rule():
    way = np.zeros((3, 4, 3))
    way[0, :, 2] = [0.0, 0.5, 1.0, 1.5]       # fine
    way[1, :, 2] = [0.0, 0.5, 0.5, 1.5]       # equal times: np.diff(t) > 0 fails (MPC_Tracking.py:118)
    way[2, :, 2] = [0.0, 0.7, 0.6, 1.5]       # decreasing
    assert handoff.times_increase_strictly(way).tolist() == [True, False, False]
    assert not handoff.times_increase_strictly(np.zeros((1, 3)))   # a single point (MPC_Tracking.py:113)


def test_names_and_time_steps():
    m = "training/models/vae_offset_sce2_cond_ld8_epoch3000.pth"
    assert handoff.tracked_name(m, "DefensiveData/DynamicBlindTown05/left/exp_12_left_3.csv") == \
        "tracked_trajectory_sce2_exp12_3.npy"
    # the reference expression, literally (Distribution.py:124-125, :144-145, :157)
    model_name_parts = os.path.basename(m).split('_')
    csv_name_parts = "exp_7_brake_10.csv".split('_')
    ref = f"tracked_trajectory_{model_name_parts[2]}_exp{csv_name_parts[1]}_{csv_name_parts[-1].split('.')[0]}.npy"
    assert handoff.tracked_name(m, "x/exp_7_brake_10.csv") == ref
    assert [handoff.scenario_time_step(f"vae_offset_{s}_cond") for s in ("sce1", "sce2", "sce3", "sce4", "other")] == \
        [0.02, 0.025, 0.015, 0.02, 0.02]


def _write_csv(path, ego_x, ego_y, yaw=-90.0, vx=0.5, vy=-7.0, usable=True):
    import pandas as pd
    n = 6
    df = pd.DataFrame({
        "frame": np.arange(n), "ego_x": np.full(n, ego_x), "ego_y": np.full(n, ego_y), "ego_yaw": np.full(n, yaw),
        "ego_vx": np.full(n, vx), "ego_vy": np.full(n, vy),
        # sce4 rule (Tools.py:101-108): ego within 40 m of sv1 and sv1_yaw >= -89.9
        "sv1_x": np.full(n, ego_x + (5.0 if usable else 500.0)), "sv1_y": np.full(n, ego_y), "sv1_yaw": np.full(n, -45.0),
        "sv1_vx": np.ones(n), "sv1_vy": np.ones(n), "sv2_vx": np.ones(n), "sv2_vy": np.ones(n),
    })
    df.to_csv(path, index=False)


@pytest.mark.gpu
def test_batched_jobs_equal_the_per_csv_helper(tmp_path, golden_dir):
    import Tools
    ck = np.load(os.path.join(golden_dir, "ckpt_sce4_cond.npz"))
    model_path = str(tmp_path / "vae_offset_sce4_cond_ld8_epoch3000.pth")
    torch.save({k: torch.from_numpy(ck[k]) for k in ck.files}, model_path)
    csvs = []
    for i, (x, y, ok) in enumerate([(14.2, 80.1, True), (15.9, -20.5, True), (13.3, 5.0, False), (16.1, 60.0, True)]):
        p = str(tmp_path / f"exp_{i + 1}_unpred_{i}.csv")
        _write_csv(p, x, y, usable=ok)
        csvs.append(p)
    z = torch.randn(3, 8, generator=torch.Generator().manual_seed(4))
    jobs = handoff.generate_tracker_jobs(model_path, csvs, seq_len=10, dim=3, latent_dim=8, z=z)
    assert [os.path.basename(j.csv_path) for j in jobs] == ["exp_1_unpred_0.csv", "exp_2_unpred_1.csv", "exp_4_unpred_3.csv"]
    assert [j.save_name for j in jobs] == ["tracked_trajectory_sce4_exp1_0.npy", "tracked_trajectory_sce4_exp2_1.npy",
                                           "tracked_trajectory_sce4_exp4_3.npy"]
    model = Tools._cached_model(model_path, 10, 3, 8)
    for j, zi in zip(jobs, z):
        sx, sy = j.initial_state[0], j.initial_state[1]
        one = model.generate(torch.tensor([[sx, sy]], dtype=torch.float64).float(), z=zi[None], add_start=True).cpu().numpy()[0]
        ref = one[:, [1, 2, 0]]
        ref[0, 2] = 0.0
        np.testing.assert_array_equal(j.waypoints, ref)          # batching changes nothing, bit for bit
        assert j.waypoints.dtype == np.float32 and j.waypoints[0, 2] == 0.0
        assert j.time_step == 0.02 and j.total_time == float(ref[-1, -1])
        assert j.initial_state.shape == (5,) and abs(j.initial_state[2] + np.pi / 2) < 1e-12
        assert j.trackable == bool(np.all(np.diff(ref[:, 2]) > 0))
