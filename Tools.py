"""Drop-in for the generate helpers of the reference's ``Tools.py`` on B200 kernels.

In scope: ``load_model_and_generate_trajectory``
(reference ``Tools.py:18-65``) and the batched decode inside
``visualize_trajectories`` (``Tools.py:862-912``); plus the two pure-Python helpers
the entry points need, ``get_start_conditions_from_csv`` (``:69-134``) and
``plot_losses`` (``:662-771``).  New, additive: ``generate_trajectories`` - the
batched sample-and-decode the reference runs one trajectory (and one checkpoint
load) at a time.

The rest of the reference's ``Tools.py`` (``get_human_and_bv_trajectories``, GIF
animation, ``create_smooth_curve``, CSV re-stamping: plotting / CSV glue outside the
accelerated path) is NOT re-implemented.  Its callers (``Distribution.py:10``,
``Plot_case.py:11``, ``Traj_Tracking_Intact.py:5-6``, ``Plot_Gif.py:24``) keep working
when the reference's own file stays next to this one under the name ``Tools_host.py``:
every public name this module does not define is taken from there (INTEGRATION.md
section 2).  Without that file those names raise an ImportError that says so.
"""
import csv
import math
import os
import sys

import numpy as np
import pandas as pd  # noqa: F401
import torch
# names the reference's ``from Tools import *`` hands to its importers (Tools.py:1-14, star-imported at
# Training_VAE.py:102): kept importable from here
import torch.nn as nn  # noqa: F401
import torch.optim as optim  # noqa: F401
from torch.utils.data import DataLoader, Dataset  # noqa: F401
from tqdm import tqdm  # noqa: F401

_PKG = os.path.join(os.path.dirname(os.path.abspath(__file__)), "defensive-model-vae_b200")
if _PKG not in sys.path:
    sys.path.insert(0, _PKG)

from dmvae.model import ConditionalTrajectoryVAE  # noqa: E402

# ------------------------------------------------------------------------------------------
# model cache: the reference rebuilds the module and re-reads the checkpoint for every single
# trajectory (Tools.py:39-41); here a checkpoint is loaded once per (path, mtime, shape)
# ------------------------------------------------------------------------------------------
_MODEL_CACHE = {}


def _cached_model(model_path, seq_len, dim, latent_dim):
    st = os.stat(model_path)
    key = (os.path.abspath(model_path), st.st_mtime_ns, st.st_size, seq_len, dim, latent_dim)
    model = _MODEL_CACHE.get(key)
    if model is None:
        # the default initialisation draws from the host generator: forked, so that a cache miss and a cache hit
        # leave the caller's RNG stream in the same state (the draws the reference makes are burnt explicitly,
        # _burn_init_draws)
        with torch.random.fork_rng(devices=[]):
            model = ConditionalTrajectoryVAE(seq_len, dim, latent_dim)
        model.load_state_dict(torch.load(model_path, map_location='cpu'))
        model.eval()
        if len(_MODEL_CACHE) >= 16:
            _MODEL_CACHE.clear()
        _MODEL_CACHE[key] = model
    return model


def load_model_and_generate_trajectory(model_path, start_x, start_y, seq_len=12, dim=3, latent_dim=8, device='cpu'):
    """One trajectory ``(seq_len, 3)`` float32 ``[t, x, y]`` in global coordinates.

    Same contract as the reference (``Tools.py:18-65``): z ~ N(0, I) of shape
    ``(1, latent_dim)`` drawn with ``torch.randn`` on the host generator, decode
    conditioned on the start point, then ``x = start_x + dx``, ``y = start_y + dy``
    as one fp32 add of fp32(start) (time column untouched).  ``device`` is accepted
    for compatibility (the reference passes 'cpu'); the kernels run on the current
    CUDA device.

    RNG parity: the reference constructs a fresh, randomly initialised module before
    ``torch.randn`` and so advances the generator by the initialisation draws; the
    same draws are burnt here, so a seeded call returns the reference's trajectory
    (to fp32 tolerance)."""
    model = _cached_model(model_path, seq_len, dim, latent_dim)
    _burn_init_draws(seq_len, dim, latent_dim)
    z = torch.randn(1, latent_dim)
    start = np.array([start_x, start_y], ndmin=2)
    out = model.generate(torch.from_numpy(start).float(), z=z, add_start=True)
    return out.cpu().numpy()[0]


def _burn_init_draws(seq_len, dim, latent_dim, hidden=128):
    """Advance the host generator exactly as ``ConditionalTrajectoryVAE(...)``'s default
    initialisation does (one uniform_ per weight and per bias, constructor order)."""
    shapes = [(hidden, 2), (hidden, hidden), (hidden, seq_len * dim), (hidden, hidden), (hidden, hidden),
              (hidden, hidden), (latent_dim, 2 * hidden), (latent_dim, 2 * hidden), (hidden, latent_dim + hidden),
              (hidden, hidden), (hidden, hidden), (seq_len * dim, hidden)]
    for out_f, in_f in shapes:
        torch.empty(out_f, in_f).uniform_(-1, 1)
        torch.empty(out_f).uniform_(-1, 1)


def generate_trajectories(model_path, start_points, num_samples=None, seq_len=10, dim=3, latent_dim=8, seed=0,
                          sample_offset=0, z=None, as_numpy=True):
    """Batched generation: ``num_samples`` trajectories for one shared start point
    ``(x, y)`` or one per row of ``start_points`` ``(n, 2)``.  Latents are drawn in the
    kernel with Philox keyed by (``seed``, ``sample_offset`` + row), so any contiguous
    sharding of the rows over GPUs reproduces the single-GPU result bit for bit.
    Returns ``(n, seq_len, 3)`` float32 ``[t, x, y]``."""
    model = _cached_model(model_path, seq_len, dim, latent_dim)
    sp = np.asarray(start_points, dtype=np.float64).reshape(-1, 2)
    out = model.generate(torch.from_numpy(sp).float(), z=z, n=num_samples, seed=seed, sample_offset=sample_offset)
    return out.cpu().numpy() if as_numpy else out


# ------------------------------------------------------------------------------------------
# start conditions from a CARLA log (pure pandas; reference Tools.py:69-134)
# ------------------------------------------------------------------------------------------
_DEFAULT_START = {"sce1": (-193.3, 50.0), "sce2": (-155.0, -5.0), "sce4": (11.0, 0.0), None: (155.0, -15.0)}


def _scenario_key(model_name):
    for key in ("sce1", "sce2", "sce4"):
        if key in model_name:
            return key
    return None  # sce3 and anything else share the last rule


def _start_mask(df, key):
    if key == "sce1":
        return (df['ego_y'] >= 18) & (df['sv2_vx'] != 0) & (df['sv2_vy'] != 0)
    if key == "sce2":
        return df['sv1_yaw'] < -170
    if key == "sce4":
        d2 = (df['ego_x'] - df['sv1_x']) ** 2 + (df['ego_y'] - df['sv1_y']) ** 2
        return (d2 <= 40 ** 2) & (df['sv1_yaw'] >= -89.9)
    return (df['sv1_vx'] != 0) & (df['sv1_vy'] != 0) & (df['ego_y'] <= 40) & (df['ego_y'] != 0)


def get_start_conditions_from_csv(csv_path, model_name):
    """First log row satisfying the scenario's start rule ->
    ``(start_x, start_y, start_angle_rad, start_vx, start_vy)``.

    As in the reference, when no row qualifies or the file cannot be read the
    scenario default is returned as a THREE-tuple ``(x, y, -pi/2)`` (its callers
    unpack five values - a reference defect kept as is, not silently changed)."""
    key = _scenario_key(model_name)
    dx, dy = _DEFAULT_START[key]
    fallback = (dx, dy, -90 * math.pi / 180)
    try:
        df = pd.read_csv(csv_path)
        mask = _start_mask(df, key)
        if not mask.any():
            print("warning: no row satisfies the start condition, using the scenario default")
            return fallback
        row = df[mask].iloc[0]
        start_x, start_y = row['ego_x'], row['ego_y']
        start_angle = row['ego_yaw'] * math.pi / 180
        start_vx, start_vy = row['ego_vx'], row['ego_vy']
        print(f"start condition from CSV: x={start_x:.2f}, y={start_y:.2f}, angle={start_angle:.2f}rad,"
              f"vx={start_vx:.2f}, vy={start_vy:.2f}")
        return start_x, start_y, start_angle, start_vx, start_vy
    except Exception as e:  # same print-and-continue style as the reference
        print(f"failed to read CSV: {e}")
        return fallback


# ------------------------------------------------------------------------------------------
# loss curves (reference Tools.py:662-771): PNG when matplotlib is present, CSV always
# ------------------------------------------------------------------------------------------
def plot_losses(loss_history, epochs, save_path="training/loss/loss.png"):
    for key, values in loss_history.items():
        if len(values) != epochs:
            raise ValueError(f"Length of loss_history['{key}'] ({len(values)}) does not match epochs ({epochs})")
    os.makedirs(os.path.dirname(save_path) or ".", exist_ok=True)
    try:
        import matplotlib
        matplotlib.use("Agg")
        import matplotlib.pyplot as plt
        xs = list(range(1, epochs + 1))
        fig, (ax1, ax2) = plt.subplots(1, 2, figsize=(14, 6), constrained_layout=True)
        ax1.plot(xs, loss_history['total_loss'], label='Total Loss', linewidth=2.0)
        ax1.set_xlabel('Epoch'); ax1.set_ylabel('Loss'); ax1.set_title('Total Loss'); ax1.grid(True, linestyle='--')
        for key, label in (('recon_loss', 'Reconstruction Loss'), ('kld_loss', 'KL Divergence Loss'),
                           ('start_loss', 'Starting Point Loss'), ('time_loss', 'Time Loss')):
            ax2.plot(xs, loss_history[key], label=label, linewidth=1.8)
        ax2.set_xlabel('Epoch'); ax2.set_ylabel('Loss'); ax2.set_title('Component Losses'); ax2.grid(True, linestyle='--')
        ax2.legend()
        fig.savefig(save_path, dpi=300, bbox_inches='tight')
        plt.close(fig)
        print(f"Loss plots saved to: {save_path}")
    except ImportError:
        print("matplotlib is not installed: skipping the PNG, writing the CSV only")
    csv_path = os.path.splitext(save_path)[0] + ".csv"
    keys = list(loss_history.keys())
    with open(csv_path, mode="w", newline="", encoding="utf-8") as f:
        writer = csv.writer(f)
        writer.writerow(keys)  # header = the five loss keys
        for i in range(epochs):
            writer.writerow([loss_history[k][i] for k in keys])
    print(f"Loss history saved to CSV: {csv_path}")


# ------------------------------------------------------------------------------------------
# batched decode for the visual comparison (reference Tools.py:834-912)
# ------------------------------------------------------------------------------------------
def generate_for_visualization(model, dataset, use_training_start_end=True, custom_start_end=None,
                               train_traj_start=0, train_traj_end=9):
    """The decode block of ``visualize_trajectories``: row i of the result is
    conditioned on the start point of training trajectory ``train_traj_start + i``
    (or on the one custom start), z = ``torch.randn(num_samples, latent_dim)`` on the
    host generator.  Returns (train_data, generated_samples), both ``(n, T, 3)``."""
    model.eval()
    num_samples = train_traj_end - train_traj_start
    train_data = dataset.data[train_traj_start:train_traj_end]
    if not use_training_start_end and custom_start_end is not None:
        sx, sy = custom_start_end[0]
        start_points = torch.tensor([[sx, sy]] * num_samples, dtype=torch.float32)
    else:
        start_points = torch.from_numpy(np.ascontiguousarray(train_data[:, 0, 1:3])).float()
    with torch.no_grad():
        h_condition = model.condition_encoder(start_points)                 # Tools.py:898
        z = torch.randn(num_samples, model.latent_dim)                      # Tools.py:901
        rel = model.decode(z, h_condition).cpu().numpy()                    # Tools.py:904
    sp = start_points.cpu().numpy()
    generated = rel.copy()
    generated[:, :, 1] = sp[:, 0:1] + rel[:, :, 1]                          # Tools.py:908-912, fp32
    generated[:, :, 2] = sp[:, 1:2] + rel[:, :, 2]
    return train_data, generated


def visualize_trajectories(model, dataset, model_save_path, axis_flip='none', use_training_start_end=True,
                           custom_start_end=None, train_traj_start=0, train_traj_end=9):
    """Generate ``train_traj_end - train_traj_start`` trajectories next to the training
    ones and, when matplotlib is available, draw them.  (The plotting tail of the
    reference calls ``create_smooth_curve`` with a mismatching signature and cannot
    run as written - SURVEY.md section 3D; a plain line plot is drawn instead.)"""
    train_data, generated = generate_for_visualization(model, dataset, use_training_start_end, custom_start_end,
                                                       train_traj_start, train_traj_end)
    print("\n=== time column of the generated trajectories ===")
    for i, g in enumerate(generated[:3]):
        print(f"  generated {i + 1}: t = {np.array2string(g[:, 0], precision=2)}")
    try:
        import matplotlib
        matplotlib.use("Agg")
        import matplotlib.pyplot as plt
    except ImportError:
        print("matplotlib is not installed: returning the arrays without plotting")
        return train_data, generated
    fig, ax = plt.subplots(figsize=(8, 8))
    for i in range(len(generated)):
        ax.plot(train_data[i, :, 1], train_data[i, :, 2], 'b-', alpha=0.6, label='human' if i == 0 else None)
        ax.plot(generated[i, :, 1], generated[i, :, 2], 'r--', alpha=0.8, label='generated' if i == 0 else None)
    if 'x' in axis_flip:
        ax.invert_xaxis()
    if 'y' in axis_flip:
        ax.invert_yaxis()
    ax.set_xlabel('x [m]'); ax.set_ylabel('y [m]'); ax.legend()
    out = os.path.splitext(model_save_path)[0] + "_samples.png"
    fig.savefig(out, dpi=200, bbox_inches='tight')
    plt.close(fig)
    print(f"figure saved to {out}")
    return train_data, generated


# ------------------------------------------------------------------------------------------
# the rest of the reference's Tools.py: served from the reference's own file, kept as Tools_host.py
# ------------------------------------------------------------------------------------------
_HOST_NAMES = ("get_human_and_bv_trajectories", "process_model_trajectory", "create_vehicle_rectangle",
               "plot_gif_human_vs_model", "save_animation_as_gif", "create_smooth_curve")   # reference Tools.py:138-830


def _merge_host_module():
    """Every public name of ``Tools_host`` (the reference's unmodified Tools.py under that name) that this module
    does not define itself becomes importable from here; the accelerated entry points above always win."""
    try:
        import importlib
        host = importlib.import_module("Tools_host")
    except ImportError:
        return False
    mine = globals()
    for name, value in vars(host).items():
        if not name.startswith("_") and name not in mine:
            mine[name] = value
    return True


_HOST_MERGED = _merge_host_module()


def __getattr__(name):
    if name in _HOST_NAMES:
        raise ImportError(
            f"Tools.{name} is plotting / CSV glue of the reference's Tools.py that the B200 drop-in does not "
            "re-implement: keep the reference's file next to this one as Tools_host.py (INTEGRATION.md section 2)")
    raise AttributeError(f"module 'Tools' has no attribute {name!r}")
