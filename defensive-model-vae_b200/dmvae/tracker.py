"""Batched MPC path tracker: the GPU counterpart of ``MPC/MPC_Tracking.py``.

The reference tracks ONE waypoint set per ``PathTracker`` object, one SLSQP solve per 15-25 ms time step (about a second
of host time each, 200-600 steps per trajectory): ``Distribution.py:91-105`` builds
``PathTracker(waypoints, initial_state, 2.8, 30, 20, dt)`` and calls ``run_simulation(waypoints[-1, -1])``.  Here all
trajectories of a batch are tracked at once, one GPU thread each (``dmvae_mpc_prepare`` / ``dmvae_mpc_track``,
``csrc/dmvae_mpc.cu``), the controller's optimisation problem solved to convergence by a Newton-type method instead
of SLSQP's early stop at ftol = 1e-6: trajectories agree with the reference's to that early-stopping noise
(``tests/test_mpc_gpu.py`` states the tolerances), not bit for bit.

* ``track_batch``     waypoints ``(n, T, 3)`` ``[x, y, t]`` + initial states ``(n, 5)`` ``[x, y, theta, vx, vy]`` ->
  ``TrackResult`` (states ``(n, S + 1, 4)`` ``[x, y, theta, v]``, controls ``(n, S, 2)``, steps per trajectory);
* ``PathTracker``     the reference's class surface for one trajectory (same constructor, ``run_simulation``, the
  recorded ``trajectory`` / ``controls`` / ``times``), so that ``Distribution.process_single_trajectory`` runs unchanged;
* ``run_tracker_jobs`` the second half of ``Distribution.batch_process_trajectories`` for the jobs of
  ``handoff.generate_tracker_jobs``: tracks them in one batch and writes ``tracked_trajectory_*.npy``.

No CPU path: raises without a CUDA device.  Two to 64 waypoints per trajectory (cubic interpolants from four on, the
reference's quadratic / linear fall-backs for three / two).
"""
from __future__ import annotations

import ctypes
import os
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import DmvaeMpcCfg, check, ptr, stream_ptr


def mpc_config(n_way: int, way_f32: bool, prediction_horizon: int = 10, control_horizon: int = 5, wheelbase: float = 2.8,
               max_steer: float = 0.5, max_accel: float = 7.0, max_iter: int = 50, tol: float = 1e-6) -> DmvaeMpcCfg:
    """``VehicleModel`` limits (``MPC_Tracking.py:26``) and ``MPCController`` weights (``:304-306``); the horizons default
    to the class defaults (``:283-284``) - ``Distribution.py:98-99`` passes 30 and 20."""
    if control_horizon > prediction_horizon:
        raise ValueError("the control horizon cannot exceed the prediction horizon")      # MPC_Tracking.py:300-301
    return DmvaeMpcCfg(n_way=n_way, way_f32=int(way_f32), horizon=prediction_horizon, blocks=control_horizon, max_iter=max_iter,
                       reserved=0, wheelbase=wheelbase, max_steer=max_steer, max_accel=max_accel, q_theta=20.0, q_v=5.0,
                       r_accel=1.0, r_steer=50.0, tol=tol)


@dataclass
class TrackResult:
    states: torch.Tensor        # (n, S + 1, 4) float64 [x, y, theta, v]; rows beyond a trajectory's own steps repeat its last state
    controls: torch.Tensor      # (n, S, 2) float64 [a, delta]; zero beyond a trajectory's own steps
    n_steps: np.ndarray         # (n,) steps of every trajectory: int(total_time / dt)
    trackable: np.ndarray       # (n,) bool: waypoint times increase strictly (else the reference raises; rows are NaN)
    iterations: np.ndarray      # (n,) solver iterations summed over the steps
    dt: float

    def trajectory(self, j: int) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        """(times, states, controls) of trajectory j as ``PathTracker.run_simulation`` returns them (``:523``)."""
        s = int(self.n_steps[j])
        times = np.array([0.0] + [i * self.dt + self.dt for i in range(s)])        # :452, :491
        return times, self.states[j, :s + 1].cpu().numpy(), self.controls[j, :s].cpu().numpy()


def steps_of(total_time, dt: float) -> np.ndarray:
    """``int(total_time / dt)`` (``MPC_Tracking.py:505``) element-wise, on the dtype ``total_time`` comes in - the
    float32 last waypoint time when the waypoints are the VAE's (``Distribution.py:104``)."""
    return np.atleast_1d(np.asarray(total_time) / dt).astype(np.int64)


class BatchTracker:
    """Set-up once (``dmvae_mpc_prepare``), then ``advance`` in chunks of steps: state, previous control and previous
    solution of every trajectory live in the workspace between the calls."""

    def __init__(self, waypoints, initial_states, dt: float, prediction_horizon: int = 10, control_horizon: int = 5,
                 wheelbase: float = 2.8, total_time=None, max_iter: int = 50, tol: float = 1e-6):
        if not torch.cuda.is_available():
            raise _lib.DmvaeError("no CUDA device is visible and dmvae has no CPU path")
        dev = torch.device("cuda", torch.cuda.current_device())
        w = torch.as_tensor(waypoints)
        if w.dim() != 3 or w.shape[2] != 3 or w.shape[0] < 1:
            raise ValueError(f"expected (n, T, 3) [x, y, t] waypoints, got {tuple(w.shape)}")
        if w.dtype not in (torch.float32, torch.float64):
            w = w.to(torch.float64)
        self.way = w.detach().to(dev).contiguous()
        self.n, n_way = int(w.shape[0]), int(w.shape[1])
        init = torch.as_tensor(np.asarray(initial_states, dtype=np.float64) if not torch.is_tensor(initial_states) else initial_states)
        if tuple(init.shape) != (self.n, 5):
            raise ValueError(f"expected ({self.n}, 5) initial states [x, y, theta, vx, vy], got {tuple(init.shape)}")
        self.init = init.detach().to(dev, torch.float64).contiguous()
        self.dt = float(dt)
        self.cfg = mpc_config(n_way, self.way.dtype == torch.float32, prediction_horizon, control_horizon, wheelbase,
                              max_iter=max_iter, tol=tol)
        lib = _lib.lib()
        nbytes = lib.dmvae_mpc_workspace_bytes(ctypes.byref(self.cfg), self.n)
        if nbytes < 0:
            raise _lib.DmvaeError(lib.dmvae_last_error().decode())
        self.workspace = torch.empty((nbytes + 7) // 8, dtype=torch.float64, device=dev)
        self.state = torch.empty(self.n, 4, dtype=torch.float64, device=dev)
        self.status = torch.empty(self.n, dtype=torch.int32, device=dev)
        self.profile = torch.empty(self.n, 5, dtype=torch.float64, device=dev)
        with torch.cuda.device(dev):
            check(lib.dmvae_mpc_prepare(ctypes.byref(self.cfg), ptr(self.way), ptr(self.init), self.n, ptr(self.workspace),
                                        ptr(self.state), ptr(self.status), ptr(self.profile), stream_ptr()), "dmvae_mpc_prepare")
        if total_time is None:
            total_time = self.way[:, -1, 2].cpu().numpy()          # waypoints[-1, -1], in the waypoints' dtype
        steps = steps_of(total_time, self.dt)
        if steps.shape[0] == 1 and self.n > 1:
            steps = np.repeat(steps, self.n)
        self.n_steps = np.maximum(steps, 0)
        self.n_steps_dev = torch.from_numpy(self.n_steps.astype(np.int32)).to(dev)
        self.iters = torch.zeros(self.n, dtype=torch.int32, device=dev)
        self.step = 0

    def advance(self, step_count: int, states_out: Optional[torch.Tensor] = None, controls_out: Optional[torch.Tensor] = None) -> None:
        """Steps ``[self.step, self.step + step_count)`` of every trajectory that has them.  ``states_out``
        ``(n, rows, 4)`` / ``controls_out`` ``(n, rows - 1, 2)`` are indexed by the global step."""
        rows = 0
        if states_out is not None:
            rows = int(states_out.shape[1])
            if controls_out is not None and int(controls_out.shape[1]) != rows - 1:
                raise ValueError("controls_out must have one row less than states_out")
        elif controls_out is not None:
            rows = int(controls_out.shape[1]) + 1
        with torch.cuda.device(self.state.device):
            check(_lib.lib().dmvae_mpc_track(ctypes.byref(self.cfg), ptr(self.workspace), self.n, self.dt, ptr(self.n_steps_dev),
                                             ptr(self.status), self.step, int(step_count), ptr(self.state),
                                             ptr(states_out) if states_out is not None else None,
                                             ptr(controls_out) if controls_out is not None else None, rows, ptr(self.iters),
                                             stream_ptr()), "dmvae_mpc_track")
        self.step += int(step_count)

    def set_previous_control(self, controls) -> None:
        """Previous applied control ``(n, 2)`` ``[a, delta]`` of every trajectory (the controller's ``last_control``,
        ``MPC_Tracking.py:308-309``): for resuming a run from recorded data; the previous solution is cleared."""
        c = torch.as_tensor(np.asarray(controls, dtype=np.float64)).to(self.state.device)
        if tuple(c.shape) != (self.n, 2):
            raise ValueError(f"expected ({self.n}, 2) controls, got {tuple(c.shape)}")
        n_way, n = self.cfg.n_way, self.n
        scal = n_way + 8 * (n_way - 1)              # workspace fields: knots, 2 x 4 (n_way - 1) coefficients, 8 scalars, warm start
        ws = self.workspace
        ws[(scal + 5) * n:(scal + 6) * n] = c[:, 0]
        ws[(scal + 6) * n:(scal + 7) * n] = c[:, 1]
        ws[(scal + 7) * n:(scal + 8) * n] = 1.0
        ws[(scal + 8) * n:(scal + 8 + 2 * self.cfg.blocks) * n] = 0.0

    def windows(self, times: Sequence[float]) -> torch.Tensor:
        """The reference windows ``[theta_ref, v_ref]`` that ``PathTracker.step`` would assemble at the given current times
        (``MPC_Tracking.py:465-478``): ``(n, len(times), horizon + 1, 2)``."""
        t = np.ascontiguousarray(np.asarray(times, dtype=np.float64))
        out = torch.empty(self.n, len(t), self.cfg.horizon + 1, 2, dtype=torch.float64, device=self.state.device)
        with torch.cuda.device(self.state.device):
            check(_lib.lib().dmvae_mpc_windows(ctypes.byref(self.cfg), ptr(self.workspace), self.n, self.dt,
                                               t.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), len(t), ptr(self.status), ptr(out),
                                               stream_ptr()), "dmvae_mpc_windows")
        return out


def track_batch(waypoints, initial_states, dt: float, total_time=None, prediction_horizon: int = 30, control_horizon: int = 20,
                wheelbase: float = 2.8, max_steps: Optional[int] = None) -> TrackResult:
    """``PathTracker(w, s, wheelbase, prediction_horizon, control_horizon, dt).run_simulation(total_time)`` for every
    row of the batch (horizons default to what ``Distribution.py:98-99`` passes).  ``total_time``: scalar or ``(n,)``,
    default each trajectory's last waypoint time."""
    # Rows are tracked in the order of their step counts (a stable sort, undone on the way out): one thread tracks one
    # trajectory from its first step to its last, so the lanes of a warp should finish together - the synthetic sets
    # of the bench have 150 to 990 steps per trajectory.  A row's result does not depend on where it sits in the batch.
    w = torch.as_tensor(waypoints)
    init = initial_states if torch.is_tensor(initial_states) else torch.as_tensor(np.asarray(initial_states, dtype=np.float64))
    n = int(w.shape[0])
    tt = w[:, -1, 2].cpu().numpy() if total_time is None else total_time
    steps = steps_of(tt, float(dt))
    if steps.shape[0] == 1 and n > 1:
        steps = np.repeat(steps, n)
    order = np.argsort(-steps, kind="stable")
    permuted = bool((order != np.arange(n)).any())
    if permuted:
        sel = torch.from_numpy(order)
        w, init = w[sel.to(w.device)], init[sel.to(init.device)]
        if np.ndim(tt) > 0 and np.shape(tt)[0] == n:
            tt = np.asarray(tt)[order]
    bt = BatchTracker(w, init, dt, prediction_horizon, control_horizon, wheelbase, tt)
    if max_steps is not None:
        bt.n_steps = np.minimum(bt.n_steps, max_steps)
        bt.n_steps_dev = torch.from_numpy(bt.n_steps.astype(np.int32)).to(bt.state.device)
    status = bt.status.cpu().numpy()
    ok = status == 0
    S = int(bt.n_steps[ok].max()) if ok.any() else 0
    states = torch.full((bt.n, S + 1, 4), float("nan"), dtype=torch.float64, device=bt.state.device)
    controls = torch.zeros((bt.n, S, 2), dtype=torch.float64, device=bt.state.device)
    bt.advance(S, states, controls)
    # rows beyond a trajectory's own last step: its final state (a rectangular array is more useful than NaN tails)
    steps_dev = torch.from_numpy(bt.n_steps).to(bt.state.device).clamp(max=S)
    idx = torch.minimum(torch.arange(S + 1, device=bt.state.device)[None, :], steps_dev[:, None])
    states = torch.gather(states, 1, idx[:, :, None].expand(-1, -1, 4))
    n_steps, iters = bt.n_steps.copy(), bt.iters.cpu().numpy()
    if permuted:
        inv = np.empty(n, dtype=np.int64)
        inv[order] = np.arange(n)
        back = torch.from_numpy(inv).to(states.device)
        states, controls = states[back], controls[back]
        n_steps, iters, ok = n_steps[inv], iters[inv], ok[inv]
    return TrackResult(states=states, controls=controls, n_steps=n_steps, trackable=ok, iterations=iters, dt=float(dt))


class PathTracker:
    """The reference's ``PathTracker`` (``MPC_Tracking.py:418-523``) for one waypoint set, on the GPU.

    Same constructor and ``run_simulation`` / ``step`` surface and the same recorded ``trajectory``, ``controls``,
    ``times`` lists.  ``initial_state`` ``[x, y, theta, vx, vy]``: like the reference, a heading below -2.8 rad is
    wrapped IN the caller's array (``:435-436``)."""

    def __init__(self, waypoints: np.ndarray, initial_state: np.ndarray, wheelbase: float = 2.8, prediction_horizon: int = 10,
                 control_horizon: int = 5, dt: float = 0.01):
        if initial_state[2] < -2.8:
            initial_state[2] = initial_state[2] + 2 * np.pi
        self.waypoints = waypoints
        self.dt = dt
        way = np.asarray(waypoints)
        if way.ndim != 2 or way.shape[1] != 3:
            raise ValueError(f"expected (N, 3) [x, y, t] waypoints, got {way.shape}")
        if len(way) < 2:
            raise ValueError("at least two waypoints are needed")                       # :114-115
        if not np.all(np.diff(way[:, 2]) > 0):
            raise ValueError("waypoint times must increase strictly")                   # :118-119
        self._bt = BatchTracker(way[None], np.asarray(initial_state, dtype=np.float64)[None], dt, prediction_horizon,
                                control_horizon, wheelbase, total_time=0.0)
        self._bt.n_steps_dev.fill_(2 ** 31 - 1)       # step() may be called any number of times
        self.current_state = self._bt.state[0].cpu().numpy()
        self.trajectory = [self.current_state.copy()]
        self.controls: List[np.ndarray] = []
        self.times = [0.0]

    def step(self, current_time: float) -> Tuple[np.ndarray, np.ndarray]:
        """One tracking step at ``current_time`` (``:454-493``).  The kernel derives the time from the step index
        (``i * dt``, as ``run_simulation`` does), so ``current_time`` must be a multiple of ``dt``."""
        i = int(round(current_time / self.dt))
        if abs(i * self.dt - current_time) > 1e-9 * max(1.0, abs(current_time)):
            raise ValueError("step(): current_time must be a whole number of time steps")
        bt = self._bt
        bt.state[0].copy_(torch.from_numpy(np.asarray(self.current_state, dtype=np.float64)))
        bt.step = i
        ctrl = torch.zeros(1, i + 1, 2, dtype=torch.float64, device=bt.state.device)
        bt.advance(1, None, ctrl)
        self.current_state = bt.state[0].cpu().numpy()
        control = ctrl[0, i].cpu().numpy()
        self.trajectory.append(self.current_state.copy())
        self.controls.append(control.copy())
        self.times.append(current_time + self.dt)
        return self.current_state.copy(), control

    def run_simulation(self, total_time: float) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        """All ``int(total_time / dt)`` steps in one launch (``:495-523``).  Like the reference's loop, the run starts at
        time 0 (``current_time = i * dt`` for ``i`` from 0) from the CURRENT state and previous control, whatever was
        tracked before, and appends to the recorded lists."""
        num_steps = int(total_time / self.dt)
        bt = self._bt
        bt.state[0].copy_(torch.from_numpy(np.asarray(self.current_state, dtype=np.float64)))
        bt.step = 0
        states = torch.empty(1, num_steps + 1, 4, dtype=torch.float64, device=bt.state.device)
        controls = torch.empty(1, num_steps, 2, dtype=torch.float64, device=bt.state.device)
        bt.advance(num_steps, states, controls)
        st = states[0, 1:].cpu().numpy()
        ct = controls[0].cpu().numpy()
        for i in range(num_steps):
            self.trajectory.append(st[i].copy())
            self.controls.append(ct[i].copy())
            self.times.append(i * self.dt + self.dt)
        if num_steps:
            self.current_state = st[-1].copy()
        return np.array(self.times), np.array(self.trajectory), np.array(self.controls)


def run_tracker_jobs(jobs, save_dir: Optional[str] = "results/GeneratedData", prediction_horizon: int = 30,
                     control_horizon: int = 20):
    """The tracking half of ``Distribution.batch_process_trajectories`` (``:114-166``) for the jobs of
    ``handoff.generate_tracker_jobs``: one batched launch per time step value, then ``np.save`` of every state sequence
    ``(S + 1, 4)`` under the reference's file name.  Untrackable jobs are skipped with the reference's message.
    Returns ``(all_trajectories, all_times, saved_files)`` like the reference."""
    all_traj, all_times, saved = [], [], []
    usable = [j for j in jobs if j.trackable]
    for j in jobs:
        if not j.trackable:
            print(f"Error processing {j.csv_path}: waypoint times must increase strictly")
    for dt in sorted({j.time_step for j in usable}):
        group = [j for j in usable if j.time_step == dt]
        way = np.stack([j.waypoints for j in group])
        init = np.stack([np.asarray(j.initial_state, dtype=np.float64) for j in group])
        res = track_batch(way, init, dt, prediction_horizon=prediction_horizon, control_horizon=control_horizon)
        for k, j in enumerate(group):
            times, states, _ = res.trajectory(k)
            all_traj.append(states)
            all_times.append(times)
            if save_dir is not None:
                os.makedirs(save_dir, exist_ok=True)
                path = os.path.join(save_dir, j.save_name)
                np.save(path, states)
                saved.append(path)
    return all_traj, all_times, saved
