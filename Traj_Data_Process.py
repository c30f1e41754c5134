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
# 'end_cond' keeps the reference's row-wise form; 'end_cond_vec' is the same rule on whole columns.
SCENE_CONFIG = {
    'StaticBlindTown05': {
        'start_cond': lambda df: (df['ego_y'] >= 18) & (df['sv2_vx'] != 0) & (df['sv2_vy'] != 0),
        'end_cond': lambda row: row['ego_y'] >= 95,
        'end_cond_vec': lambda df: df['ego_y'] >= 95,
    },
    'DynamicBlindTown05': {
        'start_cond': lambda df: df['sv1_yaw'] < -170,
        'end_cond': lambda row: row['ego_x'] < -186,
        'end_cond_vec': lambda df: df['ego_x'] < -186,
    },
    'PredictableMovementTown05': {
        'start_cond': lambda df: (df['sv1_vx'] != 0) & (df['sv1_vy'] != 0) & (df['ego_y'] <= 40) & (df['ego_y'] != 0),
        'end_cond': lambda row: row['ego_y'] <= -80,
        'end_cond_vec': lambda df: df['ego_y'] <= -80,
    },
    'UnpredictableMovementTown04': {
        'start_cond': lambda df: ((df['ego_x'] - df['sv1_x']) ** 2 + (df['ego_y'] - df['sv1_y']) ** 2 <= 40 ** 2)
                                 & (df['sv1_yaw'] >= -89.9),
        'end_cond': lambda row: (row['sv1_x'] > 15) and (row['sv1_yaw'] < -85),
        'end_cond_vec': lambda df: (df['sv1_x'] > 15) & (df['sv1_yaw'] < -85),
    },
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


# ===================== one log -> one trajectory (reference :72-122) =====================
def process_csv(csv_path, scene, action, target_points=5, point_mode='normal', time_interval=0.015):
    df = pd.read_csv(csv_path)
    config = SCENE_CONFIG[scene]
    # first row that satisfies the start rule (:75-81)
    start_mask = np.asarray(config['start_cond'](df), dtype=bool)
    if not start_mask.any():
        return None
    start_idx = int(np.argmax(start_mask))
    # first row AFTER the start row that satisfies the end rule; it is excluded (:85-95)
    end_mask = np.asarray(config['end_cond_vec'](df), dtype=bool)
    after = end_mask[start_idx + 1:]
    stop = start_idx + 1 + int(np.argmax(after)) if after.any() else len(df)
    sub_df = df.iloc[start_idx:stop]
    if 'ego_x' not in sub_df.columns or 'ego_y' not in sub_df.columns:
        return None
    traj = sub_df[['ego_x', 'ego_y']].values
    if len(traj) < target_points:
        return None
    # equidistant picks including both ends; dtype=int truncates (:106)
    indices = np.linspace(0, len(traj) - 1, target_points, dtype=int)
    print(((len(traj) - 1) * time_interval) / (target_points - 1))
    if point_mode == 'normal':
        traj = traj[indices]
    elif point_mode == 'extend_mid':
        part1 = indices[:-1]
        part2 = indices[1:]
        indices1 = np.ceil((part1 + part2) / 2).astype(int)
        indices_new = np.append(np.insert(indices1[:-1], 0, indices[0]), indices[-1])
        traj = traj[indices_new]
    # time column: sample k at k * tick * (n - 1) / (T - 1) (:118)
    times = np.arange(target_points) * time_interval * ((len(sub_df) - 1) / (target_points - 1))
    return np.column_stack((times, traj))


def collect_trajectories(data_root, scenes, actions, target_points=5, point_mode='normal', time_interval=0.015):
    all_trajs = []
    for scene in scenes:
        scene_path = os.path.join(data_root, scene)
        for action in actions:
            action_path = os.path.join(scene_path, action)
            if not os.path.exists(action_path):
                continue
            for fname in os.listdir(action_path):
                if fname.endswith('.csv'):
                    csv_path = os.path.join(action_path, fname)
                    traj = process_csv(csv_path, scene, action, target_points, point_mode, time_interval)
                    if traj is not None and len(traj) == target_points:
                        all_trajs.append(traj)
                    else:
                        print(f"No trajectory found for {scene}, {action}, {fname}")
    return all_trajs


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
