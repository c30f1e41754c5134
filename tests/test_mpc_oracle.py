"""CPU restatement of the reference's MPC tracker (oracle/mpc_oracle.py) against the reference's own outputs
(tests/golden/mpc_track.npz, produced by running MPC/MPC_Tracking.py's PathTracker: oracle/make_golden_mpc.py).

Tolerances: what PathInterpolator derives (headings, end velocity) and the reference windows are the same SciPy calls on
the same values: exact.  Closed-loop states go through SLSQP with finite-difference gradients, which amplifies
last-place differences of the objective (Python floats here, small NumPy arrays there): measured 2e-9 .. 2e-5 on states,
<= 4e-4 on controls over the golden segments; stated 1e-4 / 2e-3.  The converged solver (solve_exact, not in the
reference) lands within the reference's early-stopping noise: measured <= 1e-4 on states, <= 1.5e-3 on controls; stated
5e-4 / 5e-3."""
import os

import numpy as np
import pytest

from oracle import mpc_oracle as O


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "mpc_track.npz"))


NAMES = ["sce1_brake", "sce2_west", "sce4_south", "turn", "sce3_stop", "f64_beyond"]


def _profile(gold, name):
    init = gold[f"{name}_init"].copy()
    if init[2] < -2.8:
        init[2] += 2 * np.pi
    return O.SpeedProfile(gold[f"{name}_way"], init)


@pytest.mark.parametrize("name", NAMES)
def test_profile_and_windows_equal_the_reference(gold, name):
    assert list(gold["names"]) == NAMES
    p = _profile(gold, name)
    got = np.array([p.start_theta, p.end_vx, p.end_vy, p.end_theta, p.t_end])
    np.testing.assert_array_equal(got, gold[f"{name}_profile"])
    dt = float(gold[f"{name}_dt"])
    for k, ct in enumerate(gold[f"{name}_win_times"]):
        np.testing.assert_array_equal(p.window(float(ct), dt, 30), gold[f"{name}_windows"][k])
    assert O.tracker_steps(gold[f"{name}_way"][-1, -1], dt) == int(gold[f"{name}_steps_total"])
    assert p.turned == (name == "turn")


@pytest.mark.parametrize("name", NAMES)
def test_closed_loop_follows_the_reference(gold, name):
    way, init, dt = gold[f"{name}_way"], gold[f"{name}_init"], float(gold[f"{name}_dt"])
    K = min(4, int(gold[f"{name}_segments"][0][1]))
    _, states, controls = O.track(way, init, dt, max_steps=K)
    assert np.abs(states - gold[f"{name}_seg0_states"][:K + 1]).max() < 1e-4
    assert np.abs(controls - gold[f"{name}_seg0_controls"][:K]).max() < 2e-3
    _, states, controls = O.track(way, init, dt, max_steps=K, solver="exact")
    assert np.abs(states - gold[f"{name}_seg0_states"][:K + 1]).max() < 5e-4
    assert np.abs(controls - gold[f"{name}_seg0_controls"][:K]).max() < 5e-3


def test_a_complete_run_of_the_converged_solver_stays_with_the_reference(golden_dir):
    """tests/golden/mpc_track_full.npz: the reference's complete sce2 run (222 controller calls).  The feedback loop keeps
    SLSQP's early-stopping noise from adding up: measured <= 6e-5 on states, 4e-4 on controls over the whole run."""
    full = np.load(os.path.join(golden_dir, "mpc_track_full.npz"))
    way, init, dt = full["sce2_west_way"], full["sce2_west_init"], float(full["sce2_west_dt"])
    times, states, controls = O.track(way, init, dt, solver="exact")
    np.testing.assert_array_equal(times, full["sce2_west_times"])
    assert np.abs(states - full["sce2_west_states"]).max() < 5e-4
    assert np.abs(controls - full["sce2_west_controls"]).max() < 5e-3


@pytest.mark.parametrize("name", ["three", "two", "three_f64"])
def test_three_and_two_waypoints_use_the_quadratic_and_linear_interpolants(golden_dir, name):
    """tests/golden/mpc_track_small.npz (the reference run on 3 / 2 waypoints: interp1d kind 'quadratic' / 'linear',
    MPC_Tracking.py:126-137, :173-178)."""
    g = np.load(os.path.join(golden_dir, "mpc_track_small.npz"))
    way, init, dt = g[f"{name}_way"], g[f"{name}_init"], float(g[f"{name}_dt"])
    p = O.SpeedProfile(way, init)
    np.testing.assert_array_equal(np.array([p.start_theta, p.end_vx, p.end_vy, p.end_theta, p.t_end]), g[f"{name}_profile"])
    for k, ct in enumerate(g[f"{name}_win_times"]):
        np.testing.assert_array_equal(p.window(float(ct), dt, 30), g[f"{name}_windows"][k])
    _, states, controls = O.track(way, init, dt, max_steps=4)
    assert np.abs(states - g[f"{name}_states"]).max() < 1e-4 and np.abs(controls - g[f"{name}_controls"]).max() < 2e-3


def test_analytic_gradient_matches_differences():
    rng = np.random.default_rng(3)
    ref = np.stack([1.5 + 0.01 * rng.standard_normal(31), 8 + rng.standard_normal(31)], 1)
    u = np.stack([rng.uniform(-0.4, 0.4, 20), rng.uniform(-0.2, 0.2, 20)], 1).reshape(-1)
    for last in (None, np.array([0.3, -0.01])):
        f, g = O.mpc_cost_grad(u, 1.45, 7.0, ref, last, 0.02, 30, 20)
        rows = [(float(r[0]), float(r[1])) for r in ref]
        assert abs(f - O.mpc_cost(u, 1.45, 7.0, rows, last, 0.02, 30, 20)) < 1e-10 * abs(f)
        for i in range(40):
            e = np.zeros(40)
            e[i] = 1e-6
            fd = (O.mpc_cost(u + e, 1.45, 7.0, rows, last, 0.02, 30, 20) - O.mpc_cost(u - e, 1.45, 7.0, rows, last, 0.02, 30, 20)) / 2e-6
            assert abs(fd - g[i]) < 1e-6 * max(1.0, abs(g[i]))


def test_the_late_rows_brake_at_most_half_a_metre_per_second_squared():
    """The reference's bounds list puts the steering bound on the second half of the flattened controls
    (MPC_Tracking.py:390-398): with a reference speed far below the current one, rows 10..19 stop at -0.5."""
    ref = np.stack([np.full(31, 1.5), np.full(31, 2.0)], 1)
    seq, _ = O.solve_slsqp(1.5, 10.0, ref, None, 0.02)
    assert seq[:10, 0].min() < -3.0 and np.all(seq[10:, 0] >= -0.5 - 1e-9) and seq[10:, 0].min() < -0.49
    exact, _ = O.solve_exact(1.5, 10.0, ref, None, 0.02)
    assert np.abs(exact - seq).max() < 2e-2
    assert np.all(exact[10:, 0] >= -0.5 - 1e-12)
