"""Generate -> track hand-off: everything between the CARLA logs and the MPC
tracker of the reference's batch driver, with the VAE part done ONCE for all logs of a scenario.

Reference flow, per CSV (``Distribution.py:51-111`` called from ``:114-166``): start conditions from the log
(``Tools.get_start_conditions_from_csv``) -> one generated trajectory (``Tools.load_model_and_generate_trajectory``,
a checkpoint load + one B = 1 decode) -> columns reordered ``[t, x, y] -> [x, y, t]`` and ``t0 := 0``
(``:77-78``) -> ``PathTracker`` (SLSQP MPC; its batched GPU counterpart is ``dmvae/tracker.py``) -> ``np.save`` under
``results/GeneratedData/tracked_trajectory_{sce}_exp{n}_{k}.npy`` (``:157``).

Here: the start conditions of all CSVs are read first, ONE batched decode with a per-row start point produces
every waypoint set (``dmvae_decode``), and each result carries what the tracker call needs (initial state,
scenario time step, output name) plus the check the tracker would fail on (``MPC/MPC_Tracking.py:117-119``
raises unless the waypoint times increase strictly).  The tracking itself and the ``tracked_trajectory_*`` files are
``dmvae.tracker.run_tracker_jobs(jobs)``.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np
import torch


def scenario_time_step(model_name: str) -> float:
    """Tracker time step of the scenario a checkpoint belongs to (Distribution.py:80-89)."""
    if "sce1" in model_name:
        return 0.02
    if "sce2" in model_name:
        return 0.025
    if "sce3" in model_name:
        return 0.015
    if "sce4" in model_name:
        return 0.02
    return 0.02


def tracked_name(model_path: str, csv_path: str) -> str:
    """File name the reference saves the tracked states under (Distribution.py:124-125, :144-145, :157):
    ``tracked_trajectory_{model.split('_')[2]}_exp{csv.split('_')[1]}_{csv.split('_')[-1] without extension}.npy``."""
    model_parts = os.path.basename(model_path).split('_')
    csv_parts = os.path.basename(csv_path).split('_')
    return f"tracked_trajectory_{model_parts[2]}_exp{csv_parts[1]}_{csv_parts[-1].split('.')[0]}.npy"


def to_tracker_waypoints(traj_txy: np.ndarray) -> np.ndarray:
    """``[t, x, y]`` trajectories ``(..., T, 3)`` -> the tracker's ``[x, y, t]`` with the first time forced to 0
    (Distribution.py:77-78).  Returns a new array of the same dtype."""
    traj = np.asarray(traj_txy)
    if traj.shape[-1] != 3 or traj.ndim < 2:
        raise ValueError(f"expected (..., T, 3) [t, x, y], got {traj.shape}")
    way = traj[..., [1, 2, 0]].copy()
    way[..., 0, 2] = 0.0
    return way


def times_increase_strictly(waypoints_xyt: np.ndarray) -> np.ndarray:
    """Per trajectory: would ``PathInterpolator`` accept it (MPC/MPC_Tracking.py:113-119: at least two points,
    ``np.all(np.diff(t) > 0)``)?  Input ``(..., T, 3)`` ``[x, y, t]``; returns a bool array of shape ``(...)``."""
    t = np.asarray(waypoints_xyt)[..., 2]
    if t.shape[-1] < 2:
        return np.zeros(t.shape[:-1], dtype=bool)
    return np.all(np.diff(t, axis=-1) > 0, axis=-1)


@dataclass
class TrackerJob:
    """Everything ``PathTracker(waypoints, initial_state, 2.8, 30, 20, dt)`` (Distribution.py:94-101) and the
    save step need for one CSV."""
    csv_path: str
    save_name: str               # tracked_trajectory_*.npy, relative to results/GeneratedData
    waypoints: np.ndarray        # (T, 3) float32 [x, y, t], t0 = 0
    initial_state: np.ndarray    # [x, y, theta, vx, vy] (Distribution.py:80)
    time_step: float
    total_time: float            # waypoints[-1, -1] (Distribution.py:104)
    trackable: bool              # waypoint times increase strictly


def generate_tracker_jobs(model_path: str, csv_files: Sequence[str], seq_len: int = 10, dim: int = 3,
                          latent_dim: int = 8, z: Optional[torch.Tensor] = None) -> List[TrackerJob]:
    """The VAE half of ``Distribution.batch_process_trajectories`` for all ``csv_files`` at once.

    Latents: ``z`` ``(n, latent_dim)`` if given; otherwise one ``torch.randn(1, latent_dim)`` per CSV from the
    host generator, in file order - the draws the reference loop makes (``Tools.py:46``), without the module
    re-initialisation it also does per call.  CSVs whose start condition cannot be read get the scenario's
    default start like in the reference (``Tools.py:101-108``); the reference then fails on them while unpacking
    (a 3-tuple into 5 names) and skips the file - they are skipped here too."""
    import Tools   # the drop-in module at the repository root (start conditions, cached checkpoint)
    model_name = os.path.basename(model_path)
    rows, starts = [], []
    for path in csv_files:
        cond = Tools.get_start_conditions_from_csv(path, model_name)
        if len(cond) != 5:           # the reference's fallback: process_single_trajectory raises and skips the file
            print(f"Error processing {path}: no start condition")
            continue
        rows.append((path, cond))
        starts.append([cond[0], cond[1]])
    if not rows:
        return []
    n = len(rows)
    if z is None:
        z = torch.cat([torch.randn(1, latent_dim) for _ in range(n)], 0)
    elif tuple(z.shape) != (n, latent_dim):
        raise ValueError(f"z must be ({n}, {latent_dim}) for the {n} usable CSVs, got {tuple(z.shape)}")
    model = Tools._cached_model(model_path, seq_len, dim, latent_dim)
    start = torch.from_numpy(np.asarray(starts, dtype=np.float64)).float()      # fp32(start), as Tools.py:49-51
    traj = model.generate(start, z=z, add_start=True).cpu().numpy()             # (n, T, 3) [t, x, y], global
    way = to_tracker_waypoints(traj)
    ok = times_increase_strictly(way)
    dt = scenario_time_step(model_name)
    jobs = []
    for i, (path, cond) in enumerate(rows):
        jobs.append(TrackerJob(csv_path=path, save_name=tracked_name(model_path, path), waypoints=way[i],
                               initial_state=np.array([cond[0], cond[1], cond[2], cond[3], cond[4]]),
                               time_step=dt, total_time=float(way[i, -1, -1]), trackable=bool(ok[i])))
    return jobs
