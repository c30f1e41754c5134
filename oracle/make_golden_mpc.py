"""Golden vectors for the MPC tracker, produced by RUNNING THE REFERENCE'S OWN ``PathTracker``.

TEST INFRASTRUCTURE.  Run once in the build container (where ``/root/reference`` is mounted); takes a few minutes
(the reference needs about a second per controller call):

    python -m oracle.make_golden_mpc            # tests/golden/mpc_track.npz
    python -m oracle.make_golden_mpc --full     # tests/golden/mpc_track_full.npz (two complete runs, ~10 minutes)
    python -m oracle.make_golden_mpc --small    # tests/golden/mpc_track_small.npz (three / two waypoints)

tests/golden/mpc_track.npz, per case ``c``: the waypoints ``[x, y, t]`` (float32 as the VAE hands them over, one
case float64), the initial state ``[x, y, theta, vx, vy]``, the time step, what ``PathInterpolator`` derived
(start / end heading, end velocity), reference windows ``[theta_ref, v_ref]`` assembled exactly as
``PathTracker.step`` does (``MPC/MPC_Tracking.py:465-478``) at several times including beyond the last waypoint, and
closed-loop segments: ``PathTracker.step`` called ``K`` times from step index ``j0`` with a given state and previous
control (``j0 = 0``: the tracker's own start, no previous control), giving states ``(K + 1, 4)`` and the first
controls ``(K, 2)``.
"""
from __future__ import annotations

import contextlib
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)

from oracle.ref_loader import REFERENCE_ROOT, _stub_matplotlib  # noqa: E402


def load_tracker_module():
    _stub_matplotlib()
    sys.dont_write_bytecode = True
    sys.path.insert(0, os.path.join(REFERENCE_ROOT, "MPC"))
    try:
        import MPC_Tracking
    finally:
        sys.path.pop(0)
        sys.modules.pop("MPC_Tracking", None)
    return MPC_Tracking


def waypoints_along(rng, n, start, heading, v0, v1, dt_lo, dt_hi, lateral, dtype, turn=0.0):
    """[x, y, t] waypoints that start at `start`, move along `heading` (rad) with the speed falling from v0 to v1,
    a lateral random walk of `lateral` metres per waypoint and an extra heading change of `turn` rad over the set."""
    t = np.concatenate([[0.0], np.cumsum(rng.uniform(dt_lo, dt_hi, n - 1))])
    v = np.linspace(v0, v1, n)
    ds = v[:-1] * np.diff(t)
    th = heading + np.linspace(0.0, turn, n - 1)
    lat = rng.normal(0.0, lateral, n - 1)
    dx = ds * np.cos(th) - lat * np.sin(th)
    dy = ds * np.sin(th) + lat * np.cos(th)
    x = start[0] + np.concatenate([[0.0], np.cumsum(dx)])
    y = start[1] + np.concatenate([[0.0], np.cumsum(dy)])
    return np.stack([x, y, t], 1).astype(dtype)


def cases():
    rng = np.random.default_rng(20261019)
    out = []
    # sce1: +y, braking from 12 to 3 m/s (the late control rows hit their +-0.5 limit, see oracle/mpc_oracle.py)
    w = waypoints_along(rng, 10, (-194.25, 18.99), np.pi / 2, 12.0, 3.0, 0.5, 1.2, 0.15, np.float32)
    out.append(dict(name="sce1_brake", way=w, init=[w[0, 0], w[0, 1], np.pi / 2, 0.3, 12.0], dt=0.02, steps0=10,
                    late=[(150, 8)]))
    # sce2: -x (heading near +-pi: the one-sided wrap at -2.8), 0.025 s steps
    w = waypoints_along(rng, 10, (-110.0, -1.5), np.pi, 9.0, 6.0, 0.4, 0.9, 0.10, np.float32)
    out.append(dict(name="sce2_west", way=w, init=[w[0, 0], w[0, 1], -3.1, -9.0, -0.2], dt=0.025, steps0=10,
                    late=[(100, 6)]))
    # sce4: -y with a swerve, speeding up
    w = waypoints_along(rng, 10, (14.5, 100.0), -np.pi / 2, 6.0, 11.0, 0.6, 1.0, 0.40, np.float32)
    out.append(dict(name="sce4_south", way=w, init=[w[0, 0], w[0, 1], -np.pi / 2, 0.1, -6.0], dt=0.02, steps0=8,
                    late=[(200, 6)]))
    # a turn of 120 degrees: the end velocity comes from the middle of the last interval, late references switch to it
    w = waypoints_along(rng, 12, (0.0, 0.0), 0.3, 8.0, 7.0, 0.5, 0.8, 0.05, np.float32, turn=2.1)
    out.append(dict(name="turn", way=w, init=[0.0, 0.0, 0.3, 7.6, 2.4], dt=0.02, steps0=6, late=[(250, 6)]))
    # nearly stopping: reference speeds below 0.1 m/s hold the previous heading
    w = waypoints_along(rng, 10, (155.0, 39.5), -np.pi / 2, 4.0, 0.0, 0.8, 1.4, 0.02, np.float32)
    out.append(dict(name="sce3_stop", way=w, init=[w[0, 0], w[0, 1], -np.pi / 2, 0.0, -4.0], dt=0.015, steps0=6,
                    late=[(560, 6)]))
    # float64 waypoints (a caller that does not come from the VAE), beyond the last waypoint
    w = waypoints_along(rng, 10, (5.0, -3.0), 0.8, 10.0, 9.0, 0.3, 0.5, 0.10, np.float64)
    out.append(dict(name="f64_beyond", way=w, init=[5.0, -3.0, 0.8, 7.0, 7.1], dt=0.02, steps0=6,
                    late=[(int(w[-1, 2] / 0.02) - 3, 8)]))
    return out


def main():
    M = load_tracker_module()
    gold = {}
    names = []
    for c in cases():
        name, way, dt = c["name"], c["way"], c["dt"]
        init = np.array(c["init"], dtype=np.float64)
        names.append(name)
        quiet = io.StringIO()
        with contextlib.redirect_stdout(quiet):
            tracker = M.PathTracker(way.copy(), init.copy(), 2.8, 30, 20, dt)
        pi = tracker.path_interp
        gold[f"{name}_way"] = way
        gold[f"{name}_init"] = init
        gold[f"{name}_dt"] = np.float64(dt)
        gold[f"{name}_profile"] = np.array([pi.start_theta, pi.end_vx, pi.end_vy, pi.end_theta, pi.t_end])
        gold[f"{name}_steps_total"] = np.int64(int(way[-1, -1] / dt))      # run_simulation(total_time=waypoints[-1, -1])
        # reference windows as PathTracker.step assembles them
        t_end = float(way[-1, 2])
        win_times = np.array([0.0, 0.37 * t_end, 0.81 * t_end, t_end - 10 * dt, t_end + 5 * dt])
        wins = np.zeros((len(win_times), 31, 2))
        with contextlib.redirect_stdout(quiet):
            for k, ct in enumerate(win_times):
                held = 0.0
                for i in range(31):
                    t_ref = float(ct) + i * dt
                    _, _, vx, vy = pi.get_reference(t_ref)
                    v_ref = np.sqrt(vx ** 2 + vy ** 2)
                    if v_ref >= 0.1:
                        held = pi.get_reference_heading(t_ref)
                    wins[k, i] = [held, v_ref]
        gold[f"{name}_win_times"] = win_times
        gold[f"{name}_windows"] = wins
        # closed-loop segments
        segs = [(0, c["steps0"])] + list(c["late"])
        gold[f"{name}_segments"] = np.array(segs, dtype=np.int64)
        for s, (j0, K) in enumerate(segs):
            with contextlib.redirect_stdout(quiet):
                tr = M.PathTracker(way.copy(), init.copy(), 2.8, 30, 20, dt)
                if j0 > 0:
                    _, _, vx, vy = tr.path_interp.get_reference(j0 * dt)
                    th = float(np.arctan2(vy, vx))
                    th = th if th >= -2.8 else th + 2 * np.pi
                    tr.current_state = np.array([1.0 + s, -2.0, th + 0.02, float(np.hypot(vx, vy)) * 0.97 + 0.1])
                    tr.mpc.last_control = np.array([-0.4, 0.004])
                    gold[f"{name}_seg{s}_last"] = tr.mpc.last_control.copy()
                states = [tr.current_state.copy()]
                controls = []
                for j in range(j0, j0 + K):
                    st, u = tr.step(j * dt)
                    states.append(st.copy())
                    controls.append(u.copy())
            gold[f"{name}_seg{s}_states"] = np.array(states)
            gold[f"{name}_seg{s}_controls"] = np.array(controls)
            print(name, "segment", s, "from step", j0, "->", np.array(states)[-1], flush=True)
    gold["names"] = np.array(names)
    os.makedirs(GOLD, exist_ok=True)
    np.savez_compressed(os.path.join(GOLD, "mpc_track.npz"), **gold)
    print("wrote", os.path.join(GOLD, "mpc_track.npz"))


def main_full():
    """tests/golden/mpc_track_full.npz: two COMPLETE reference runs (run_simulation(waypoints[-1, -1]) as
    Distribution.py:104-105 calls it, 401 and 222 controller calls), for the end-to-end gap between the reference's
    early-stopped SLSQP and a converged solver over a whole trajectory.  About ten minutes."""
    M = load_tracker_module()
    gold = {}
    for c in cases()[:2]:
        name, way, dt = c["name"], c["way"], c["dt"]
        init = np.array(c["init"], dtype=np.float64)
        with contextlib.redirect_stdout(io.StringIO()):
            tr = M.PathTracker(way.copy(), init.copy(), 2.8, 30, 20, dt)
            times, states, controls = tr.run_simulation(total_time=way[-1, -1])
        gold[f"{name}_way"], gold[f"{name}_init"], gold[f"{name}_dt"] = way, init, np.float64(dt)
        gold[f"{name}_times"], gold[f"{name}_states"], gold[f"{name}_controls"] = times, states, controls
        print(name, states.shape, states[-1], flush=True)
    np.savez_compressed(os.path.join(GOLD, "mpc_track_full.npz"), **gold)
    print("wrote", os.path.join(GOLD, "mpc_track_full.npz"))


def main_small():
    """tests/golden/mpc_track_small.npz: three and two waypoints - the reference's quadratic / linear interpolants
    (MPC_Tracking.py:126-137, :173-178): what PathInterpolator derives, windows, four controller calls."""
    M = load_tracker_module()
    rng = np.random.default_rng(7)
    gold = {}
    for name, n, dtype in (("three", 3, np.float32), ("two", 2, np.float32), ("three_f64", 3, np.float64)):
        way = waypoints_along(rng, n, (3.0, -2.0), 0.6, 9.0, 6.0, 0.8, 1.3, 0.2, dtype)
        init = np.array([3.0, -2.0, 0.6, 7.4, 5.1])
        dt = 0.02
        quiet = io.StringIO()
        with contextlib.redirect_stdout(quiet):
            tr = M.PathTracker(way.copy(), init.copy(), 2.8, 30, 20, dt)
            pi = tr.path_interp
            t_end = float(way[-1, 2])
            win_times = np.array([0.0, 0.4 * t_end, t_end - 10 * dt, t_end + 3 * dt])
            wins = np.zeros((len(win_times), 31, 2))
            for k, ct in enumerate(win_times):
                held = 0.0
                for i in range(31):
                    t_ref = float(ct) + i * dt
                    _, _, vx, vy = pi.get_reference(t_ref)
                    v_ref = np.sqrt(vx ** 2 + vy ** 2)
                    if v_ref >= 0.1:
                        held = pi.get_reference_heading(t_ref)
                    wins[k, i] = [held, v_ref]
            states, controls = [tr.current_state.copy()], []
            for j in range(4):
                st, u = tr.step(j * dt)
                states.append(st.copy())
                controls.append(u.copy())
        gold[f"{name}_way"], gold[f"{name}_init"], gold[f"{name}_dt"] = way, init, np.float64(dt)
        gold[f"{name}_profile"] = np.array([pi.start_theta, pi.end_vx, pi.end_vy, pi.end_theta, pi.t_end])
        gold[f"{name}_win_times"], gold[f"{name}_windows"] = win_times, wins
        gold[f"{name}_states"], gold[f"{name}_controls"] = np.array(states), np.array(controls)
        print(name, np.array(states)[-1], flush=True)
    np.savez_compressed(os.path.join(GOLD, "mpc_track_small.npz"), **gold)
    print("wrote", os.path.join(GOLD, "mpc_track_small.npz"))


if __name__ == "__main__":
    if "--small" in sys.argv:
        main_small()
    elif "--full" in sys.argv:
        main_full()
    else:
        main()
