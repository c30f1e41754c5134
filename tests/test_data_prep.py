"""Data preparation drop-in: ``Traj_Data_Process.process_csv`` with its vectorised cut against
(1) a row-by-row restatement of the reference lines on synthetic logs (runs everywhere), (2) the reference module
itself on the same logs and on real logs, and the shipped ``trajectory_sce1_cond.npy`` (only where
``/root/reference`` is mounted)."""
import importlib.util
import os
import random

import numpy as np
import pandas as pd
import pytest

import Traj_Data_Process as P

REF = "/root/reference"
SCENES = list(P.SCENE_CONFIG)


def restated_process_csv(csv_path, scene, target_points, point_mode, time_interval):
    """Traj_Data_Process.py:72-122 of the reference, row by row (test oracle)."""
    df = pd.read_csv(csv_path)
    cfg = P.SCENE_CONFIG[scene]
    start_idx = None
    for idx, val in enumerate(cfg['start_cond'](df)):
        if val:
            start_idx = idx
            break
    if start_idx is None:
        return None
    sub = df.iloc[start_idx:].copy()
    end_idx = None
    for i, row in sub.iterrows():
        if i == sub.index[0]:
            continue
        if cfg['end_cond'](row):
            end_idx = i
            break
    if end_idx is not None:
        sub = sub.loc[:end_idx - 1]
    traj = sub[['ego_x', 'ego_y']].values
    if len(traj) < target_points:
        return None
    idx = np.linspace(0, len(traj) - 1, target_points, dtype=int)
    if point_mode == 'normal':
        traj = traj[idx]
    else:
        mid = np.ceil((idx[:-1] + idx[1:]) / 2).astype(int)
        traj = traj[np.append(np.insert(mid[:-1], 0, idx[0]), idx[-1])]
    times = np.arange(target_points) * time_interval * ((len(sub) - 1) / (target_points - 1))
    return np.column_stack((times, traj))


def synthetic_log(path, scene, seed, n=400, start_at=37, end_at=301):
    """A log whose start rule first holds at row `start_at` and whose end rule first holds (after it) at `end_at`
    (None: never); the columns every rule reads, plus noise rows that satisfy the end rule BEFORE the start."""
    rng = np.random.default_rng(seed)
    k = np.arange(n)
    df = pd.DataFrame({"frame": k, "ego_x": rng.normal(size=n).cumsum(), "ego_y": rng.normal(size=n).cumsum(),
                       "sv1_x": np.zeros(n), "sv1_y": np.zeros(n), "sv1_yaw": np.zeros(n), "sv1_vx": np.zeros(n),
                       "sv1_vy": np.zeros(n), "sv2_vx": np.zeros(n), "sv2_vy": np.zeros(n)})
    started = k >= start_at
    ended = (k >= end_at) if end_at is not None else np.zeros(n, bool)
    early = k < 5                                        # end rule true before the start: must be ignored
    if scene == 'StaticBlindTown05':
        df["ego_y"] = np.where(ended | early, 96.0, np.where(started, 20.0 + 0.1 * k, 10.0))
        df["sv2_vx"] = np.where(started, 1.0, 0.0); df["sv2_vy"] = np.where(started, 1.0, 0.0)
    elif scene == 'DynamicBlindTown05':
        df["sv1_yaw"] = np.where(started, -175.0, -100.0)
        df["ego_x"] = np.where(ended | early, -190.0, -150.0 - 0.05 * k)
    elif scene == 'PredictableMovementTown05':
        df["sv1_vx"] = np.where(started, 2.0, 0.0); df["sv1_vy"] = np.where(started, 2.0, 0.0)
        df["ego_y"] = np.where(ended | early, -81.0, 39.0 - 0.2 * k)
    else:
        df["sv1_yaw"] = np.where(ended | early, -86.0, np.where(started, -45.0, -120.0))
        df["sv1_x"] = np.where(ended | early, 16.0, 10.0)
        df["ego_x"] = 12.0 + 0.01 * k; df["ego_y"] = 5.0 + 0.3 * k * 0.1
    df.to_csv(path, index=False)


@pytest.mark.parametrize("scene", SCENES)
@pytest.mark.parametrize("mode", ["normal", "extend_mid"])
def test_vectorised_cut_equals_the_row_walk(tmp_path, scene, mode):
    cases = [dict(), dict(end_at=None), dict(start_at=0, end_at=1), dict(start_at=395, end_at=None),
             dict(start_at=100, end_at=105), dict(start_at=10_000)]
    for i, kw in enumerate(cases):
        path = str(tmp_path / f"log_{i}.csv")
        synthetic_log(path, scene, seed=i, **kw)
        for T in (5, 10, 12):
            got = P.process_csv(path, scene, "x", T, mode, 0.02)
            ref = restated_process_csv(path, scene, T, mode, 0.02)
            if ref is None:
                assert got is None, (scene, kw, T)
            else:
                assert got is not None and got.dtype == np.float64 and got.shape == (T, 3)
                np.testing.assert_array_equal(got, ref)


def test_random_walks_reproduce_the_sequential_definition():
    random.seed(3)
    got = P.generate_random_trajectories(4, 30, 5.0)
    random.seed(3)
    ref = []
    for _ in range(4):                                   # reference :44-67
        traj = np.zeros((30, 2)); ang = 0.0
        for i in range(1, 30):
            ang += random.uniform(-np.radians(5.0), np.radians(5.0))
            traj[i, 0] = traj[i - 1, 0] + 1.0 * np.cos(ang)
            traj[i, 1] = traj[i - 1, 1] + 1.0 * np.sin(ang)
        ref.append(traj)
    np.testing.assert_array_equal(got, np.array(ref))
    assert P.generate_random_trajectories(2, 50, 0.0)[:, :, 1].max() == 0.0


def _reference_module():
    spec = importlib.util.spec_from_file_location("ref_traj_data_process", os.path.join(REF, "Traj_Data_Process.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "DefensiveData")), reason="reference tree not mounted")
def test_against_the_reference_module_and_its_shipped_dataset(tmp_path, golden_dir, capsys):
    R = _reference_module()
    # synthetic logs through both
    for scene in SCENES:
        path = str(tmp_path / f"{scene}.csv")
        synthetic_log(path, scene, seed=11)
        np.testing.assert_array_equal(P.process_csv(path, scene, "a", 10, "normal", 0.02),
                                      R.process_csv(path, scene, "a", 10, "normal", 0.02))
    # the real sce1 logs: same trajectories as the reference, and the shipped .npy (row order = os.listdir order)
    root = os.path.join(REF, "DefensiveData")
    actions = ['减速', '转向', '减速+转向']
    mine = P.collect_trajectories(root, ['StaticBlindTown05'], actions, 10, 'normal', 0.02)
    theirs = R.collect_trajectories(root, ['StaticBlindTown05'], actions, 10, 'normal', 0.02)
    capsys.readouterr()
    assert len(mine) == len(theirs) == 38
    np.testing.assert_array_equal(np.array(mine), np.array(theirs))
    shipped = np.load(os.path.join(golden_dir, "data_sce1_cond.npy"))
    key = lambda a: a[np.lexsort(a.reshape(len(a), -1).T[::-1])]
    np.testing.assert_allclose(key(np.array(mine)), key(shipped), rtol=0, atol=1e-9)
