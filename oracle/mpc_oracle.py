"""CPU restatement of the reference's MPC path tracker.

TEST INFRASTRUCTURE - not part of the product path.  Only ``tests/``, ``__graft_entry__.smoke()`` and the CPU legs
of ``bench.py`` may import this module.

What it restates: ``/root/reference/MPC/MPC_Tracking.py`` as driven by ``Distribution.py:91-105``
(``PathTracker(waypoints, initial_state, 2.8, 30, 20, dt).run_simulation(waypoints[-1, -1])``):

* ``velocity_knots`` / ``SpeedProfile``      <- ``PathInterpolator._create_interpolators`` (``:103-221``): only the
  velocity interpolants, the end velocity and the two headings reach the controller (the position interpolants are
  evaluated by ``get_reference`` but ``PathTracker.step`` drops them, ``:469-470``);
* ``SpeedProfile.velocity`` / ``.window``    <- ``get_reference`` (``:224-252``), ``get_reference_heading``
  (``:254-277``) and the reference window of ``PathTracker.step`` (``:464-478``);
* ``mpc_cost``                               <- the objective closed over in ``MPCController.solve_mpc`` (``:329-373``)
  with the bicycle model of ``VehicleModel`` (``:39-86``) reduced to the two states the cost reads (theta, v);
* ``solve_slsqp``                            <- the ``scipy.optimize.minimize(method='SLSQP')`` call and its failure
  handling (``:322-327``, ``:389-415``);
* ``track``                                  <- ``PathTracker.__init__`` / ``step`` / ``run_simulation`` (``:421-523``).

Third-party arithmetic: the cubic interpolants are ``scipy.interpolate.interp1d(kind='cubic')`` (= not-a-knot
``make_interp_spline(k=3)``, extrapolating with the end pieces) and the optimiser is SciPy's SLSQP with two-point
finite-difference gradients; the reference pins neither (container: SciPy 1.18.1, NumPy 2.3).  Mixed float32 / float64
expressions follow NumPy 2 promotion (a float32 array combined with a Python float stays float32): the waypoints that
``Distribution.py`` hands over are float32, so the mid-interval knot times, the mid time of the last interval and the
end of the heading scan are float32 values, exactly as in the reference run here.

Pinned by ``tests/golden/mpc_track.npz`` = states and controls produced by RUNNING THE REFERENCE'S OWN ``PathTracker``
(``oracle/make_golden_mpc.py``).  The objective here is evaluated with Python floats instead of small NumPy arrays
(same operations; a BLAS dot of two elements may fuse one multiply-add), and SLSQP differentiates it numerically, so the
restatement follows the reference to within the optimiser's own noise rather than bit for bit: measured
<= 1e-6 on every state over the golden runs (tolerance in ``tests/test_mpc_oracle.py``).

``solve_exact`` is NOT in the reference: the same objective minimised to machine precision (analytic gradient, SLSQP
with ftol 1e-15).  It is the yardstick for the CUDA tracker, which also converges every problem instead of stopping at
ftol = 1e-6: CUDA vs ``solver='exact'`` is held to 1e-7, CUDA vs the reference's own (early-stopped) runs to the
measured gap between the two CPU solvers.
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import numpy as np
from scipy.interpolate import interp1d
from scipy.optimize import minimize

WHEELBASE = 2.8          # MPC_Tracking.py:26 / Distribution.py:97
MAX_STEER = 0.5          # :26
MAX_ACCEL = 7.0          # :26
Q_THETA, Q_V = 20.0, 5.0         # :304, :306 (terminal weights equal the running ones)
R_ACCEL, R_STEER = 1.0, 50.0     # :305
HEADING_WRAP = -2.8              # :202, :211, :221, :273, :435
V_THRESHOLD = 0.1                # :471
SCAN_STEP = 0.001                # :204


def wrap_heading(theta: float) -> float:
    """The reference's one-sided wrap: headings below -2.8 rad move up by 2 pi (:202)."""
    return theta if theta >= HEADING_WRAP else theta + 2 * np.pi


def _kind(n: int) -> str:
    return "cubic" if n >= 4 else ("quadratic" if n >= 3 else "linear")     # :126-137, :173-178


def velocity_knots(waypoints: np.ndarray, initial_state: np.ndarray) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Knot times and values of the velocity interpolants (:146-168).  ``waypoints`` ``(n, 3)`` ``[x, y, t]`` in its
    own dtype (float32 from the VAE); ``initial_state`` ``[x, y, theta, vx, vy]``."""
    t = waypoints[:, 2]
    if len(t) < 2:
        raise ValueError("at least two waypoints")                     # :114-115
    if not np.all(np.diff(t) > 0):
        raise ValueError("waypoint times must increase strictly")      # :118-119
    kind = _kind(len(t))
    x_smooth = interp1d(t, waypoints[:, 0], kind=kind, bounds_error=False, fill_value="extrapolate")(t)
    y_smooth = interp1d(t, waypoints[:, 1], kind=kind, bounds_error=False, fill_value="extrapolate")(t)
    dt = np.diff(t)
    dt = np.where(dt == 0, 1e-6, dt)
    vx = np.concatenate((np.array([initial_state[-2]]), np.diff(x_smooth) / dt))
    vy = np.concatenate((np.array([initial_state[-1]]), np.diff(y_smooth) / dt))
    t_vel = np.concatenate((np.array([0.0]), t[:-1] + dt / 2))
    return t_vel, vx, vy


class SpeedProfile:
    """What the controller sees of a waypoint set: reference velocity / heading as functions of time."""

    def __init__(self, waypoints: np.ndarray, initial_state: np.ndarray):
        t = waypoints[:, 2]
        t_vel, vx, vy = velocity_knots(waypoints, initial_state)
        kind = _kind(len(t_vel))
        self.vx_interp = interp1d(t_vel, vx, kind=kind, bounds_error=False, fill_value="extrapolate")
        self.vy_interp = interp1d(t_vel, vy, kind=kind, bounds_error=False, fill_value="extrapolate")
        self.t_end = float(t[-1])
        start_vx = float(self.vx_interp(float(t[0])))
        start_vy = float(self.vy_interp(float(t[0])))
        self.start_theta = wrap_heading(float(np.arctan2(start_vy, start_vx)))
        # end velocity (:204-218): the mid-interval value of the last segment if the heading ever leaves a 45 degree
        # cone around the start heading on a 1 ms grid, else the value at the last waypoint
        self.turned = False
        for t1 in np.arange(0, t[-1] + SCAN_STEP, SCAN_STEP):
            th = wrap_heading(float(np.arctan2(float(self.vy_interp(t1)), float(self.vx_interp(t1)))))
            if abs(th - self.start_theta) > 45 * np.pi / 180:
                self.turned = True
                break
        t_pick = (t[-1] + t[-2]) / 2 if self.turned else self.t_end
        self.end_vx = float(self.vx_interp(t_pick))
        self.end_vy = float(self.vy_interp(t_pick))
        self.end_theta = wrap_heading(float(np.arctan2(self.end_vy, self.end_vx)))

    def velocity(self, t: float) -> Tuple[float, float]:
        """(vx_ref, vy_ref) of ``get_reference`` (:235-250)."""
        if t <= self.t_end:
            vx = float(self.vx_interp(t))
            vy = float(self.vy_interp(t))
            if abs(float(np.arctan2(vy, vx)) - self.start_theta) > 90 * np.pi / 180:
                vx, vy = self.end_vx, self.end_vy
            return vx, vy
        return self.end_vx, self.end_vy

    def heading(self, t: float) -> float:
        """``get_reference_heading`` (:264-277) without its diagnostic print."""
        if t > self.t_end:
            return wrap_heading(self.end_theta)
        vx, vy = self.velocity(t)
        return wrap_heading(np.arctan2(vy, vx))

    def window(self, current_time: float, dt: float, horizon: int) -> np.ndarray:
        """``(horizon + 1, 2)`` ``[theta_ref, v_ref]`` of one controller call (:465-478): below 0.1 m/s the heading
        of the previous row is held, and the first row starts from 0.0."""
        ref = np.zeros((horizon + 1, 2))
        held = 0.0
        for i in range(horizon + 1):
            t_ref = current_time + i * dt
            vx, vy = self.velocity(t_ref)
            v_ref = np.sqrt(vx ** 2 + vy ** 2)
            if v_ref >= V_THRESHOLD:
                held = self.heading(t_ref)
            ref[i] = [held, v_ref]
        return ref


def _clip(x: float, lim: float) -> float:
    return min(max(x, -lim), lim)


def mpc_cost(u_flat, theta0: float, v0: float, ref, last_control, dt: float, horizon: int, blocks: int) -> float:
    """The objective of ``solve_mpc`` (:329-373) on Python floats.  ``u_flat`` = ``blocks`` rows of (a, delta); the
    last row is held for the rest of the horizon (:337-339); the rollout is the explicit Euler bicycle model with the
    controls clipped inside the dynamics (:55-56, :59-62, :84); only theta and v enter the cost."""
    theta, v = theta0, v0
    cost = 0.0
    for i in range(horizon + 1):
        e_th = theta - ref[i][0]
        e_v = v - ref[i][1]
        cost += e_th * Q_THETA * e_th + e_v * Q_V * e_v
        if i < horizon:
            k = min(i, blocks - 1)
            a = _clip(u_flat[2 * k], MAX_ACCEL)
            delta = _clip(u_flat[2 * k + 1], MAX_STEER)
            theta, v = theta + v * math.tan(delta) / WHEELBASE * dt, v + a * dt
    for k in range(blocks):
        if k == 0:
            if last_control is None:
                continue                                     # no previous control: the first increment is free (:360-362)
            da, dd = u_flat[0] - last_control[0], u_flat[1] - last_control[1]
        else:
            da, dd = u_flat[2 * k] - u_flat[2 * k - 2], u_flat[2 * k + 1] - u_flat[2 * k - 1]
        cost += da * R_ACCEL * da + dd * R_STEER * dd
    return cost


def mpc_cost_grad(u_flat, theta0, v0, ref, last_control, dt, horizon, blocks):
    """Value and analytic gradient of ``mpc_cost`` inside the control bounds (adjoint sweep); used by ``solve_exact``."""
    u = np.asarray(u_flat, dtype=np.float64).reshape(blocks, 2)
    c = dt / WHEELBASE
    th = np.empty(horizon + 1)
    v = np.empty(horizon + 1)
    th[0], v[0] = theta0, v0
    for i in range(horizon):
        k = min(i, blocks - 1)
        th[i + 1] = th[i] + v[i] * math.tan(u[k, 1]) / WHEELBASE * dt
        v[i + 1] = v[i] + u[k, 0] * dt
    e_th = th - ref[:, 0]
    e_v = v - ref[:, 1]
    cost = float(np.sum(Q_THETA * e_th * e_th + Q_V * e_v * e_v))
    grad = np.zeros((blocks, 2))
    lam_th, lam_v = 2 * Q_THETA * e_th[horizon], 2 * Q_V * e_v[horizon]
    for i in range(horizon - 1, -1, -1):
        k = min(i, blocks - 1)
        tan_d = math.tan(u[k, 1])
        grad[k, 0] += lam_v * dt
        grad[k, 1] += lam_th * c * v[i] * (1 + tan_d * tan_d)
        lam_v = 2 * Q_V * e_v[i] + lam_v + lam_th * c * tan_d
        lam_th = 2 * Q_THETA * e_th[i] + lam_th
    prev = None if last_control is None else np.asarray(last_control, dtype=np.float64)
    for k in range(blocks):
        if k == 0 and prev is None:
            continue
        d = u[k] - (prev if k == 0 else u[k - 1])
        cost += R_ACCEL * d[0] * d[0] + R_STEER * d[1] * d[1]
        g = np.array([2 * R_ACCEL * d[0], 2 * R_STEER * d[1]])
        grad[k] += g
        if k > 0:
            grad[k - 1] -= g
    return cost, grad.reshape(-1)


def _bounds_and_constraint(blocks: int):
    # the reference lists the bounds as [accel] * blocks + [steer] * blocks against a variable vector that interleaves
    # (a, delta) per row (:390-394, :398): the first `blocks` entries get +-7 and the rest +-0.5 whatever they hold.
    # The inequality constraint (:376-387) applies the intended limits, so the feasible set is the intersection.
    bounds = [(-MAX_ACCEL, MAX_ACCEL)] * blocks + [(-MAX_STEER, MAX_STEER)] * blocks

    def constraint(u_flat):
        u = np.asarray(u_flat).reshape(blocks, 2)
        out = np.empty(4 * blocks)
        out[0::4] = MAX_ACCEL - u[:, 0]
        out[1::4] = u[:, 0] + MAX_ACCEL
        out[2::4] = MAX_STEER - u[:, 1]
        out[3::4] = u[:, 1] + MAX_STEER
        return out

    return bounds, constraint


def accel_limit(k: int, blocks: int) -> float:
    """The acceleration limit that row ``k`` of the control sequence really gets in the reference: +-7 while its flat
    index 2k lies in the first ``blocks`` bounds, +-0.5 (the steering bound) after that (see above)."""
    return MAX_ACCEL if 2 * k < blocks else MAX_STEER


def solve_slsqp(theta0, v0, ref, last_control, dt, horizon=30, blocks=20):
    """One controller call (:311-415).  Returns ``(control_sequence (blocks, 2), new_last_control)``."""
    u0 = np.zeros((blocks, 2))
    if last_control is not None:
        u0[0] = last_control
    ref_rows = [(float(r[0]), float(r[1])) for r in ref]
    bounds, constraint = _bounds_and_constraint(blocks)
    res = minimize(lambda z: mpc_cost(z, theta0, v0, ref_rows, last_control, dt, horizon, blocks), u0.flatten(),
                   method="SLSQP", bounds=bounds, constraints={"type": "ineq", "fun": constraint},
                   options={"maxiter": 100, "ftol": 1e-6})
    if res.success:
        seq = res.x.reshape(blocks, 2)
        return seq, seq[0].copy()
    # failure: the initial guess is returned; last_control keeps its value (it already equals u0[0]) or stays None
    return u0, (None if last_control is None else u0[0].copy())


def solve_exact(theta0, v0, ref, last_control, dt, horizon=30, blocks=20, warm=None):
    """The same problem converged to machine precision (not in the reference; see the module docstring)."""
    u0 = np.zeros((blocks, 2)) if warm is None else np.array(warm, dtype=np.float64)
    if warm is None and last_control is not None:
        u0[0] = last_control
    ref = np.asarray(ref, dtype=np.float64)
    bounds = [b for k in range(blocks) for b in ((-accel_limit(k, blocks), accel_limit(k, blocks)), (-MAX_STEER, MAX_STEER))]
    fun = lambda z: mpc_cost_grad(z, theta0, v0, ref, last_control, dt, horizon, blocks)   # noqa: E731
    z = u0.flatten()
    for _ in range(3):                                # restarts shake off an early 'iteration limit' / stall
        res = minimize(fun, z, jac=True, method="SLSQP", bounds=bounds, options={"maxiter": 500, "ftol": 1e-16})
        z = res.x
    seq = z.reshape(blocks, 2)
    return seq, seq[0].copy()


def tracker_steps(total_time, dt: float) -> int:
    """``int(total_time / dt)`` (:505) on the types the reference computes it on (``total_time`` = the float32
    ``waypoints[-1, -1]`` of ``Distribution.py:104`` when the waypoints are float32)."""
    return int(total_time / dt)


def track(waypoints: np.ndarray, initial_state: np.ndarray, dt: float, total_time=None, horizon: int = 30,
          blocks: int = 20, solver: str = "slsqp", max_steps: Optional[int] = None):
    """``PathTracker(...).run_simulation(total_time)`` (:421-452, :454-523).  Returns ``(times (S + 1,), states
    (S + 1, 4) [x, y, theta, v], controls (S, 2) [a, delta])``; ``initial_state`` ``[x, y, theta, vx, vy]`` is not
    modified (the reference wraps its heading in place, :435-436)."""
    init = np.array(initial_state, dtype=np.float64)
    if init[2] < HEADING_WRAP:
        init[2] += 2 * np.pi
    profile = SpeedProfile(waypoints, init)
    state = np.array([init[0], init[1], init[2], math.sqrt(init[3] ** 2 + init[4] ** 2)])
    steps = tracker_steps(waypoints[-1, -1] if total_time is None else total_time, dt)
    if max_steps is not None:
        steps = min(steps, max_steps)
    times, states, controls = [0.0], [state.copy()], []
    last = None
    warm = None
    for i in range(steps):
        current_time = i * dt
        ref = profile.window(current_time, dt, horizon)
        if solver == "slsqp":
            seq, last = solve_slsqp(float(state[2]), float(state[3]), ref, last, dt, horizon, blocks)
        else:
            seq, last = solve_exact(float(state[2]), float(state[3]), ref, last, dt, horizon, blocks, warm)
            warm = np.vstack([seq[1:], seq[-1:]])
        a = _clip(float(seq[0, 0]), MAX_ACCEL)
        delta = _clip(float(seq[0, 1]), MAX_STEER)
        x, y, theta, v = state
        deriv = np.array([v * np.cos(theta), v * np.sin(theta), v * np.tan(delta) / WHEELBASE, a])   # :59-64
        state = state + deriv * dt
        times.append(current_time + dt)
        states.append(state.copy())
        controls.append(np.array(seq[0], dtype=np.float64))
    return np.array(times), np.array(states), np.array(controls).reshape(-1, 2)
