"""Drop-in for the reference's ``Traj_Data_Process.py``: CARLA logs ->
``(N, target_points, 3)`` float64 ``[t, x, y]`` arrays under ``training/DefensiveDataProcessed`` - the input
format of the hot path (``TrajectoryDataset``, Training_VAE.py:105-115).

Same names, signatures, defaults, prints and results as the reference (``SCENE_CONFIG`` :8-26,
``generate_random_trajectories`` :31-69, ``process_csv`` :72-122, ``collect_trajectories`` :125-141,
``pad_and_save`` :144-151, ``__main__`` :154-202).  What changes is how ``process_csv`` finds the cut: the
reference walks the sub-frame with ``DataFrame.iterrows`` (one Series per log row, the cost of this script);
here the end rule is evaluated on whole columns (``end_cond_vec``) and the first hit after the start row is one
``argmax``.  This is host code (pandas / numpy): the job is 254 small files once per dataset, nothing for a GPU.
``tests/test_data_prep.py`` holds it bit-equal to a row-by-row restatement of the reference lines, and to the
reference module itself and its shipped ``trajectory_*_cond.npy`` where the reference tree is mounted.
"""
import os
import random

import numpy as np
import pandas as pd

# ===================== scenes and actions (reference :8-28) =====================
# One start rule and one end rule per scenario, written once on whole columns; the row-wise 'end_cond' of the
# reference's table is the same rule applied to a single row (comparisons and `&` work on scalars too).
def _sce1_start(df):
    return (df['ego_y'] >= 18) & (df['sv2_vx'] != 0) & (df['sv2_vy'] != 0)


def _sce1_end(df):
    return df['ego_y'] >= 95


def _sce2_start(df):
    return df['sv1_yaw'] < -170


def _sce2_end(df):
    return df['ego_x'] < -186


def _sce3_start(df):
    moving = (df['sv1_vx'] != 0) & (df['sv1_vy'] != 0)
    return moving & (df['ego_y'] <= 40) & (df['ego_y'] != 0)


def _sce3_end(df):
    return df['ego_y'] <= -80


def _sce4_start(df):
    gap2 = (df['ego_x'] - df['sv1_x']) ** 2 + (df['ego_y'] - df['sv1_y']) ** 2
    return (gap2 <= 40 ** 2) & (df['sv1_yaw'] >= -89.9)


def _sce4_end(df):
    return (df['sv1_x'] > 15) & (df['sv1_yaw'] < -85)


def _scene(start_rule, end_rule):
    return {'start_cond': start_rule, 'end_cond': lambda row: bool(end_rule(row)), 'end_cond_vec': end_rule}


SCENE_CONFIG = {
    'StaticBlindTown05': _scene(_sce1_start, _sce1_end),
    'DynamicBlindTown05': _scene(_sce2_start, _sce2_end),
    'PredictableMovementTown05': _scene(_sce3_start, _sce3_end),
    'UnpredictableMovementTown04': _scene(_sce4_start, _sce4_end),
}

ACTIONS = ['减速', '减速+转向', '转向']


# ===================== random-walk toy set (reference :31-69) =====================
def generate_random_trajectories(num_trajs, traj_length, max_angle_deviation=5.0):
    """``(num_trajs, traj_length, 2)`` unit-step random walks from the origin.  The heading changes by one
    ``random.uniform(-max, max)`` per step, drawn in the reference's order (trajectory by trajectory, step by
    step), so a seeded ``random`` module reproduces the reference's array; headings and positions are running
    sums in that same order (``np.cumsum`` accumulates sequentially), hence bit-equal."""
    bound = np.radians(max_angle_deviation)
    steps = max(traj_length - 1, 0)
    out = np.zeros((num_trajs, traj_length, 2))
    for n in range(num_trajs):
        turns = np.array([random.uniform(-bound, bound) for _ in range(steps)], dtype=np.float64)
        heading = np.cumsum(turns)
        out[n, 1:, 0] = np.cumsum(np.cos(heading))
        out[n, 1:, 1] = np.cumsum(np.sin(heading))
    return out


# ===================== one log -> one trajectory (reference :72-122), all logs (:125-141) =====================
def _first_true(mask):
    """Index of the first True of a boolean array, or None."""
    mask = np.asarray(mask, dtype=bool)
    return int(np.argmax(mask)) if mask.any() else None


def process_csv(csv_path, scene, action, target_points=5, point_mode='normal', time_interval=0.015):
    """One log -> ``(target_points, 3)`` float64 ``[t, x, y]`` or None (same rules and prints as the reference)."""
    log = pd.read_csv(csv_path)
    rules = SCENE_CONFIG[scene]
    first = _first_true(rules['start_cond'](log))          # :75-81 first row that satisfies the start rule
    if first is None:
        return None
    # :85-95 the cut ends before the first LATER row that satisfies the end rule (that row is excluded)
    hit = _first_true(np.asarray(rules['end_cond_vec'](log), dtype=bool)[first + 1:])
    last = len(log) if hit is None else first + 1 + hit
    cut = log.iloc[first:last]
    if not {'ego_x', 'ego_y'} <= set(cut.columns):
        return None
    xy = cut[['ego_x', 'ego_y']].values
    n = len(xy)
    if n < target_points:
        return None
    # :106 equidistant picks including both ends; dtype=int truncates
    picks = np.linspace(0, n - 1, target_points, dtype=int)
    print(((n - 1) * time_interval) / (target_points - 1))
    if point_mode == 'extend_mid':     # :110-115 the midpoints (rounded up) between neighbouring picks, ends kept
        mids = np.ceil((picks[:-1] + picks[1:]) / 2).astype(int)
        xy = xy[np.append(np.insert(mids[:-1], 0, picks[0]), picks[-1])]
    elif point_mode == 'normal':
        xy = xy[picks]
    # :118 sample k sits at k * tick * (n - 1) / (T - 1)
    stamps = np.arange(target_points) * time_interval * ((n - 1) / (target_points - 1))
    return np.column_stack((stamps, xy))


def _logs_of(data_root, scenes, actions):
    """(scene, action, file name, path) of every CSV, in the reference's visiting order (os.listdir order inside a
    scene / action directory: it fixes the row order of the saved array)."""
    for scene in scenes:
        for action in actions:
            folder = os.path.join(data_root, scene, action)
            if not os.path.exists(folder):
                continue
            for fname in os.listdir(folder):
                if fname.endswith('.csv'):
                    yield scene, action, fname, os.path.join(folder, fname)


def collect_trajectories(data_root, scenes, actions, target_points=5, point_mode='normal', time_interval=0.015):
    kept = []
    for scene, action, fname, path in _logs_of(data_root, scenes, actions):
        traj = process_csv(path, scene, action, target_points, point_mode, time_interval)
        if traj is None or len(traj) != target_points:
            print(f"No trajectory found for {scene}, {action}, {fname}")
            continue
        kept.append(traj)
    return kept


def pad_and_save(trajs, save_path):
    if save_path is not None:
        trajs_array = np.array(trajs)  # (num_samples, target_points, 3): time + trajectory
        np.save(save_path, trajs_array)
        print(f"saved {trajs_array.shape[0]} trajectories of {trajs_array.shape[1]} points each to {save_path}")
    else:
        print("No saving path, mode error")


# ===================== script (reference :154-202) =====================
# The reference edits these literals in place; they are kept as one table per run mode.
DATASET_RUN = {
    'data_root': 'DefensiveData',
    # any of 'StaticBlindTown05', 'DynamicBlindTown05', 'PredictableMovementTown05', 'UnpredictableMovementTown04'
    'scenes': ['UnpredictableMovementTown04'],
    'actions': ['减速', '转向', '减速+转向'],
    'target_points': 10,
    # log tick per scenario (reference :165-166): sce1 0.02 s (38 logs), sce2 0.025 s (16), sce3 0.015 s (66), sce4 0.02 s (135)
    'time_interval': 0.02,
    'point_mode': 'normal',          # 'normal' or 'extend_mid' (normal picks + midpoint picks, concatenated)
    'save_path': {'normal': 'training/DefensiveDataProcessed/trajectory_sce4_cond.npy',
                  'extend_mid': 'training/DefensiveDataProcessed/trajectory_sce2_extend1.npy'},
}
RANDOM_RUN = {'num_trajs': 100, 'traj_length': 50, 'max_angle_deviation': 0.0,
              'save_path': 'training/DefensiveDataProcessed/straight_trajectory.npy'}


def run_dataset(cfg=DATASET_RUN):
    mode = cfg['point_mode']
    args = (cfg['data_root'], cfg['scenes'], cfg['actions'], cfg['target_points'])
    if mode == 'normal':
        trajs = collect_trajectories(*args, 'normal', cfg['time_interval'])
    elif mode == 'extend_mid':
        trajs = (collect_trajectories(*args, 'normal', cfg['time_interval']) +
                 collect_trajectories(*args, 'extend_mid', cfg['time_interval']))
    else:
        print("Unknown point mode...")
        trajs, mode = [], None
    if not trajs:
        print('no trajectory extracted: check the parameters and the data directory')
        return None
    pad_and_save(trajs, cfg['save_path'].get(mode))
    return trajs


def run_random(cfg=RANDOM_RUN):
    os.makedirs(os.path.dirname(cfg['save_path']), exist_ok=True)
    walks = generate_random_trajectories(cfg['num_trajs'], cfg['traj_length'], cfg['max_angle_deviation'])
    np.save(cfg['save_path'], walks)
    print(f"saved {walks.shape[0]} random trajectories of length {walks.shape[1]} to {cfg['save_path']}")
    return walks


if __name__ == "__main__":
    mode = 'dataset'  # 'dataset', 'random'
    if mode == 'dataset':
        run_dataset()
    elif mode == 'random':
        run_random()
    else:
        print("mode must be 'dataset' or 'random'")
