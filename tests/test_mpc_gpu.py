"""Batched MPC tracker on the GPU (dmvae_mpc_prepare / dmvae_mpc_track / dmvae_mpc_windows through dmvae.tracker;
SURVEY.md 8f row 2) against the reference's own runs (tests/golden/mpc_track.npz, produced by MPC/MPC_Tracking.py's
PathTracker) and against the CPU oracle.

Tolerances, all absolute on [x, y, theta, v] / [a, delta]:
* what PathInterpolator derives and the reference windows: the kernel evaluates the not-a-knot cubics as piecewise
  polynomials where SciPy evaluates B-splines - same function, different rounding: 1e-9 (values up to ~15);
* closed loop against the CONVERGED CPU solver of the same problem (oracle solve_exact): 1e-7 on states, 1e-6 on controls;
* closed loop against the REFERENCE's runs: the reference stops SLSQP at ftol = 1e-6, i.e. ~1e-3 short of the minimiser in
  the controls; measured gap of the converged CPU solver to the reference <= 1e-4 on states and 1.5e-3 on controls over
  these segments: stated 5e-4 / 5e-3 ("statistical parity")."""
import os

import numpy as np
import pytest
import torch

from oracle import mpc_oracle as O

pytestmark = pytest.mark.gpu

NAMES = ["sce1_brake", "sce2_west", "sce4_south", "turn", "sce3_stop", "f64_beyond"]


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "mpc_track.npz"))


def _tracker(gold, name, **kw):
    from dmvae.tracker import BatchTracker
    return BatchTracker(gold[f"{name}_way"][None], gold[f"{name}_init"][None], float(gold[f"{name}_dt"]), 30, 20, **kw)


@pytest.mark.parametrize("name", NAMES)
def test_profile_and_windows_match_the_reference(gold, name):
    bt = _tracker(gold, name)
    assert int(bt.status[0]) == 0
    assert np.abs(bt.profile[0].cpu().numpy() - gold[f"{name}_profile"]).max() < 1e-9
    win = bt.windows(gold[f"{name}_win_times"])[0].cpu().numpy()
    assert np.abs(win - gold[f"{name}_windows"]).max() < 1e-9
    assert int(bt.n_steps[0]) == int(gold[f"{name}_steps_total"])
    init = gold[f"{name}_init"]
    th = init[2] + 2 * np.pi if init[2] < -2.8 else init[2]
    np.testing.assert_allclose(bt.state[0].cpu().numpy(), [init[0], init[1], th, np.hypot(init[3], init[4])], rtol=0, atol=1e-15)


@pytest.mark.parametrize("name", NAMES)
def test_closed_loop_against_reference_runs_and_converged_oracle(gold, name):
    from dmvae.tracker import track_batch
    way, init, dt = gold[f"{name}_way"], gold[f"{name}_init"], float(gold[f"{name}_dt"])
    segs = gold[f"{name}_segments"]
    # segment 0: the tracker's own start
    K = int(segs[0][1])
    res = track_batch(way[None], init[None], dt, max_steps=K)
    _, st, ct = res.trajectory(0)
    assert st.shape == (K + 1, 4) and ct.shape == (K, 2)
    assert np.abs(st - gold[f"{name}_seg0_states"]).max() < 5e-4
    assert np.abs(ct - gold[f"{name}_seg0_controls"]).max() < 5e-3
    _, so, co = O.track(way, init, dt, max_steps=K, solver="exact")
    assert np.abs(st - so).max() < 1e-7, np.abs(st - so).max()
    assert np.abs(ct - co).max() < 1e-6, np.abs(ct - co).max()
    assert 1 <= res.iterations[0] <= 50 * K
    # later segments: resumed from a recorded state and previous control, as the golden generator did with the reference
    for s in range(1, len(segs)):
        j0, K = int(segs[s][0]), int(segs[s][1])
        bt = _tracker(gold, name)
        bt.n_steps_dev.fill_(j0 + K)
        bt.state[0].copy_(torch.from_numpy(gold[f"{name}_seg{s}_states"][0]))
        bt.set_previous_control(gold[f"{name}_seg{s}_last"][None])
        bt.step = j0
        states = torch.zeros(1, j0 + K + 1, 4, dtype=torch.float64, device="cuda")
        controls = torch.zeros(1, j0 + K, 2, dtype=torch.float64, device="cuda")
        bt.advance(K, states, controls)
        st = states[0, j0 + 1:].cpu().numpy()
        ct = controls[0, j0:].cpu().numpy()
        assert np.abs(st - gold[f"{name}_seg{s}_states"][1:]).max() < 5e-4
        assert np.abs(ct - gold[f"{name}_seg{s}_controls"]).max() < 5e-3


@pytest.mark.parametrize("name", ["sce1_brake", "sce2_west"])
def test_complete_runs_stay_with_the_reference(golden_dir, name):
    """Two COMPLETE reference runs (run_simulation(waypoints[-1, -1]) as Distribution.py:104-105 calls it: 401 and 222
    controller calls, tests/golden/mpc_track_full.npz).  The feedback loop keeps the early-stopping noise from adding up:
    measured (converged CPU solver vs the reference) <= 9e-5 m, 1e-5 rad, 6e-5 m/s, 1e-3 m/s^2 over the whole run."""
    from dmvae.tracker import track_batch
    full = np.load(os.path.join(golden_dir, "mpc_track_full.npz"))
    way, init, dt = full[f"{name}_way"], full[f"{name}_init"], float(full[f"{name}_dt"])
    res = track_batch(way[None], init[None], dt)
    times, st, ct = res.trajectory(0)
    assert st.shape == full[f"{name}_states"].shape and ct.shape == full[f"{name}_controls"].shape
    np.testing.assert_array_equal(times, full[f"{name}_times"])
    assert np.abs(st - full[f"{name}_states"]).max() < 5e-4
    assert np.abs(ct - full[f"{name}_controls"]).max() < 5e-3


@pytest.mark.parametrize("name", ["three", "two", "three_f64"])
def test_three_and_two_waypoints(golden_dir, name):
    """The reference's quadratic / linear interpolants for fewer than four waypoints (tests/golden/mpc_track_small.npz)."""
    from dmvae.tracker import BatchTracker, track_batch
    g = np.load(os.path.join(golden_dir, "mpc_track_small.npz"))
    way, init, dt = g[f"{name}_way"], g[f"{name}_init"], float(g[f"{name}_dt"])
    bt = BatchTracker(way[None], init[None], dt, 30, 20)
    assert int(bt.status[0]) == 0
    assert np.abs(bt.profile[0].cpu().numpy() - g[f"{name}_profile"]).max() < 1e-9
    assert np.abs(bt.windows(g[f"{name}_win_times"])[0].cpu().numpy() - g[f"{name}_windows"]).max() < 1e-9
    _, st, ct = track_batch(way[None], init[None], dt, max_steps=4).trajectory(0)
    assert np.abs(st - g[f"{name}_states"]).max() < 5e-4 and np.abs(ct - g[f"{name}_controls"]).max() < 5e-3
    _, so, co = O.track(way, init, dt, max_steps=4, solver="exact")
    assert np.abs(st - so).max() < 1e-7 and np.abs(ct - co).max() < 1e-6


def test_reference_class_surface(gold):
    """The drop-in module (repo root MPC/MPC_Tracking.py) imported the way Distribution.py:9 imports the reference's, its
    PathTracker used the way Distribution.process_single_trajectory (:91-105) uses it."""
    from MPC.MPC_Tracking import PathTracker
    name = "sce2_west"
    way, init, dt = gold[f"{name}_way"], gold[f"{name}_init"].copy(), float(gold[f"{name}_dt"])
    tr = PathTracker(waypoints=way, initial_state=init, wheelbase=2.8, prediction_horizon=30, control_horizon=20, dt=dt)
    assert init[2] > 0          # the heading of the caller's array was wrapped in place, like the reference does
    K = int(gold[f"{name}_segments"][0][1])
    times, states, controls = tr.run_simulation(total_time=K * dt + 1e-9)
    assert times.shape == (K + 1,) and states.shape == (K + 1, 4) and controls.shape == (K, 2)
    np.testing.assert_allclose(times, [0.0] + [i * dt + dt for i in range(K)], rtol=0, atol=0)
    assert np.abs(states - gold[f"{name}_seg0_states"]).max() < 5e-4
    # step by step gives the same numbers as one launch (the previous solution only warm-starts the solver)
    tr2 = PathTracker(way, gold[f"{name}_init"].copy(), 2.8, 30, 20, dt)
    for i in range(K):
        state, control = tr2.step(i * dt)
    assert np.abs(np.array(tr2.trajectory) - states).max() < 1e-9
    assert np.abs(np.array(tr2.controls) - controls).max() < 1e-8
    with pytest.raises(ValueError):
        bad = way.copy()
        bad[3, 2] = bad[2, 2]
        PathTracker(bad, gold[f"{name}_init"].copy(), 2.8, 30, 20, dt)


def test_batch_rows_are_independent_and_chunks_compose(gold):
    """A row's result does not depend on its neighbours, the batch size or how the steps are split into launches."""
    from dmvae.tracker import BatchTracker, track_batch
    rng = np.random.default_rng(5)
    base = gold["sce1_brake_way"]
    n, K = 300, 12
    way = np.repeat(base[None], n, 0).copy()
    way[:, 1:, 0] += rng.normal(0, 0.2, (n, 9)).astype(np.float32)
    way[:, 1:, 1] += rng.normal(0, 0.5, (n, 9)).astype(np.float32)
    way[7, 4, 2] = way[7, 3, 2]                        # one untrackable row: repeated time stamp
    init = np.repeat(gold["sce1_brake_init"][None], n, 0).copy()
    init[:, 4] += rng.normal(0, 1.0, n)
    res = track_batch(way, init, 0.02, max_steps=K)
    assert not res.trackable[7] and res.trackable.sum() == n - 1
    assert torch.isnan(res.states[7]).all()
    for j in (0, 8, 150, 299):
        one = track_batch(way[j:j + 1], init[j:j + 1], 0.02, max_steps=K)
        assert torch.equal(one.states[0], res.states[j]) and torch.equal(one.controls[0], res.controls[j])
    bt = BatchTracker(way, init, 0.02, 30, 20)
    bt.n_steps_dev.clamp_(max=K)
    states = torch.full((n, K + 1, 4), float("nan"), dtype=torch.float64, device="cuda")
    controls = torch.zeros((n, K, 2), dtype=torch.float64, device="cuda")
    for c in (5, 4, 3):
        bt.advance(c, states, controls)
    ok = torch.from_numpy(res.trackable).cuda()
    assert (states[ok] - res.states[ok]).abs().max() < 1e-9      # chunks restart the interval cursor, nothing else
    assert (controls[ok] - res.controls[ok]).abs().max() < 1e-8
    # a converged CPU solve of one of the perturbed rows
    _, so, co = O.track(way[150], init[150], 0.02, max_steps=6, solver="exact")
    assert np.abs(res.states[150, :7].cpu().numpy() - so).max() < 1e-7


@pytest.mark.parametrize("world", [2, 3, 8])
def test_tracking_shards_like_generation(gold, world):
    """Rank r tracks rows [r n / G, (r + 1) n / G): the rows of every sharding are bit-identical to the single launch."""
    from dmvae.parallel import track_shard
    from dmvae.tracker import track_batch
    rng = np.random.default_rng(world)
    n, K = 37, 8
    way = np.repeat(gold["sce4_south_way"][None], n, 0).copy()
    way[:, 1:, :2] += rng.normal(0, 0.3, (n, 9, 2)).astype(np.float32)
    init = np.repeat(gold["sce4_south_init"][None], n, 0)
    whole = track_batch(way, init, 0.02, max_steps=K)
    seen = 0
    for r in range(world):
        lo, hi, part = track_shard(way, init, 0.02, r, world, max_steps=K)
        if part is None:
            assert lo == hi
            continue
        assert torch.equal(part.states, whole.states[lo:hi]) and torch.equal(part.controls, whole.controls[lo:hi])
        seen += hi - lo
    assert seen == n


def test_rows_of_different_lengths_come_back_in_the_callers_order(gold):
    """track_batch tracks the rows in the order of their step counts (so that the lanes of a warp finish together) and
    undoes that on the way out: every row equals the same trajectory tracked on its own, bit for bit, tails included."""
    from dmvae.tracker import track_batch
    base = gold["sce2_west_way"]
    scales = [0.30, 1.0, 0.12, 0.55, 0.12, 0.80]
    way = np.stack([base * np.array([1.0, 1.0, k], dtype=np.float32) for k in scales])
    way[:, :, :2] = base[None, :, :2] * np.array(scales, dtype=np.float32)[:, None, None]     # same speeds, shorter paths
    init = np.repeat(gold["sce2_west_init"][None], len(scales), 0)
    res = track_batch(way, init, 0.025)
    assert len(set(res.n_steps.tolist())) >= 4 and res.n_steps[1] == res.n_steps.max()
    for j in range(len(scales)):
        one = track_batch(way[j:j + 1], init[j:j + 1], 0.025)
        s = int(res.n_steps[j])
        assert s == int(one.n_steps[0])
        assert torch.equal(one.states[0], res.states[j, :s + 1]) and torch.equal(one.controls[0], res.controls[j, :s])
        assert torch.equal(res.states[j, s:], res.states[j, s:s + 1].expand(res.states.shape[1] - s, 4))   # tail = last state
        assert res.iterations[j] == one.iterations[0]


def test_saturated_controls_follow_the_effective_bounds():
    """Speed far above the reference: the first rows brake hard, rows 10..19 are held at the -0.5 that the reference's
    bounds list gives them (MPC_Tracking.py:390-398); the applied control equals the converged CPU solve."""
    from dmvae.tracker import track_batch
    t = np.arange(10) * 0.8
    way = np.stack([np.zeros(10), 2.0 * t, t], 1).astype(np.float32)           # 2 m/s along +y
    init = np.array([[0.0, 0.0, np.pi / 2, 0.0, 12.0]])                         # entering at 12 m/s
    res = track_batch(way[None], init, 0.02, max_steps=5)
    _, st, ct = res.trajectory(0)
    _, so, co = O.track(way, init[0], 0.02, max_steps=5, solver="exact")
    assert ct[0, 0] < -3.0
    assert np.abs(st - so).max() < 1e-7 and np.abs(ct - co).max() < 1e-6


def test_tracker_jobs_are_tracked_and_saved_like_the_reference_driver(gold, tmp_path):
    """run_tracker_jobs = the tracking half of Distribution.batch_process_trajectories (:114-166): every trackable job
    tracked (one launch per time step value), its (S + 1, 4) states saved under the reference's file name, untrackable
    jobs skipped; the three return values of the reference."""
    from dmvae.handoff import TrackerJob
    from dmvae.tracker import run_tracker_jobs
    jobs = []
    for k, name in enumerate(["sce1_brake", "sce2_west", "sce4_south"]):
        way, dt = gold[f"{name}_way"], float(gold[f"{name}_dt"])
        jobs.append(TrackerJob(csv_path=f"log_{k}.csv", save_name=f"tracked_trajectory_sce{k}_exp1_{k}.npy", waypoints=way,
                               initial_state=gold[f"{name}_init"].copy(), time_step=dt, total_time=float(way[-1, -1]), trackable=True))
    bad = gold["sce1_brake_way"].copy()
    bad[5, 2] = bad[4, 2]
    jobs.append(TrackerJob("log_bad.csv", "tracked_trajectory_bad.npy", bad, gold["sce1_brake_init"].copy(), 0.02, 1.0, False))
    trajs, times, saved = run_tracker_jobs(jobs, str(tmp_path))
    assert len(trajs) == len(times) == len(saved) == 3
    assert sorted(os.path.basename(f) for f in saved) == sorted(j.save_name for j in jobs[:3])
    full = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mpc_track_full.npz"))
    by_name = {os.path.basename(f): np.load(f) for f in saved}
    for k, name in enumerate(["sce1_brake", "sce2_west"]):          # the two complete reference runs
        got = by_name[f"tracked_trajectory_sce{k}_exp1_{k}.npy"]
        assert got.shape == full[f"{name}_states"].shape and got.dtype == np.float64
        assert np.abs(got - full[f"{name}_states"]).max() < 5e-4
    assert by_name["tracked_trajectory_sce2_exp1_2.npy"].shape == (int(gold["sce4_south_steps_total"]) + 1, 4)
    assert not os.path.exists(tmp_path / "tracked_trajectory_bad.npy")


def test_argument_errors_are_loud():
    from dmvae import DmvaeError
    from dmvae.tracker import BatchTracker
    way = np.zeros((1, 1, 3), dtype=np.float32)
    with pytest.raises(DmvaeError):
        BatchTracker(way, np.zeros((1, 5)), 0.02, 30, 20)         # one waypoint: the reference raises too (MPC_Tracking.py:114-115)
    with pytest.raises(ValueError):
        BatchTracker(np.zeros((1, 10, 3), dtype=np.float32), np.zeros((1, 5)), 0.02, 5, 10)
