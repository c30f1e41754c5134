"""Host-side logic of dmvae.tracker (no GPU): the step count on the reference's dtypes, configuration checks, the
input checks PathTracker makes before it touches the device, the job driver's handling of untrackable jobs."""
import numpy as np
import pytest
import torch

from dmvae import tracker
from dmvae.handoff import TrackerJob


def test_step_count_follows_the_dtype_of_the_total_time():
    """int(total_time / dt) (MPC_Tracking.py:505): float32 last waypoint times (Distribution.py:104 under NumPy 2)
    divide in float32, float64 ones in float64 - the two can differ by one step."""
    dt = 0.02
    t32 = np.array([8.02, 8.0199995, 3.0], dtype=np.float32)
    got = tracker.steps_of(t32, dt)
    want = [int(v / dt) for v in t32]                    # np.float32 scalar / Python float, as the reference computes it
    assert got.tolist() == want
    t64 = t32.astype(np.float64)
    assert tracker.steps_of(t64, dt).tolist() == [int(v / dt) for v in t64]
    assert tracker.steps_of(np.float32(1.0), 0.25).tolist() == [4]
    # a case where the two precisions disagree exists in the neighbourhood of a whole number of steps
    grid = np.nextafter((np.arange(50, 1000) * dt).astype(np.float32), np.float32(0))     # just below a whole number of steps
    a, b = tracker.steps_of(grid, dt), tracker.steps_of(grid.astype(np.float64), dt)
    assert (a != b).any() and np.abs(a - b).max() == 1


def test_configuration_matches_the_reference_defaults():
    c = tracker.mpc_config(10, True)
    assert (c.horizon, c.blocks) == (10, 5)                                   # MPCController defaults, MPC_Tracking.py:283-284
    assert (c.wheelbase, c.max_steer, c.max_accel) == (2.8, 0.5, 7.0)         # VehicleModel, :26
    assert (c.q_theta, c.q_v, c.r_accel, c.r_steer) == (20.0, 5.0, 1.0, 50.0)  # :304-306
    assert c.way_f32 == 1 and c.n_way == 10
    with pytest.raises(ValueError):
        tracker.mpc_config(10, True, prediction_horizon=5, control_horizon=6)  # :300-301


def test_path_tracker_checks_its_waypoints_like_the_reference():
    init = np.array([0.0, 0.0, -3.0, 1.0, 0.0])
    way = np.stack([np.arange(5.0), np.zeros(5), np.array([0.0, 1.0, 1.0, 2.0, 3.0])], 1)
    with pytest.raises(ValueError, match="increase strictly"):
        tracker.PathTracker(way, init)
    assert init[2] > 0            # the heading was wrapped in the caller's array first, as the reference does (:435-436)
    with pytest.raises(ValueError, match="two waypoints"):
        tracker.PathTracker(way[:1], np.zeros(5))
    with pytest.raises(ValueError):
        tracker.PathTracker(np.zeros((5, 2)), np.zeros(5))


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the behaviour WITHOUT a CUDA device")
def test_no_cpu_path():
    way = np.stack([np.arange(5.0), np.zeros(5), np.arange(5.0)], 1).astype(np.float32)
    with pytest.raises(Exception) as err:
        tracker.track_batch(way[None], np.zeros((1, 5)), 0.02)
    assert "no CPU path" in str(err.value)


def test_job_driver_skips_untrackable_jobs_without_touching_the_device(tmp_path, capsys):
    bad = np.zeros((10, 3), dtype=np.float32)
    jobs = [TrackerJob(f"log_{k}.csv", f"tracked_trajectory_x_{k}.npy", bad, np.zeros(5), 0.02, 0.0, False) for k in range(3)]
    trajs, times, saved = tracker.run_tracker_jobs(jobs, str(tmp_path))
    assert trajs == [] and times == [] and saved == [] and not list(tmp_path.iterdir())
    assert capsys.readouterr().out.count("Error processing") == 3
