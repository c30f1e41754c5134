"""Golden fixtures for the ENTRY POINTS of the path, produced by RUNNING THE REFERENCE'S OWN CODE.

TEST INFRASTRUCTURE.  Run once in the build container (where ``/root/reference`` is mounted):

    python -m oracle.make_golden_entry

``oracle/make_golden.py`` pins the model, the loss and the optimiser; this script pins what sits around
them - the code a user of the reference actually runs:

  train_entry.npz      the training mode of ``Training_VAE.py`` (its ``__main__`` block, lines 316-394,
        executed verbatim from the reference's source text with only the literals data_path / epochs /
        batch_size / save paths replaced) under ``torch.manual_seed``: DataLoader shuffle, default
        initialisation, ``randn_like`` noise, the five ``.item()`` sums per step, the weight scaling of the
        component histories, ``plot_losses`` arguments and the saved ``state_dict``.  Two runs: the
        reference configuration (batch 38 = the whole sce1 set, one step per epoch) and batch 16 (three
        steps per epoch, the last one ragged: 16 + 16 + 6 rows).
  visualize_entry.npz  ``Tools.visualize_trajectories`` (lines 834-912: the batched decode next to the
        training trajectories) on a shipped checkpoint under a seed, both start-point modes.  The plotting
        tail of that function cannot run; the arrays are taken from the function's
        frame when the stubbed matplotlib stops it.
  reg157.npz           ``Driver_Models.Reg157`` on a grid of inputs (None recorded as NaN).
"""
from __future__ import annotations

import os
import re
import sys
import tempfile
import textwrap

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)

from oracle.make_golden import digest, reduced  # noqa: E402
from oracle.ref_loader import REFERENCE_ROOT, load_reference  # noqa: E402


def reference_main_block() -> str:
    """The body of ``if __name__ == "__main__":`` of the reference's Training_VAE.py, dedented."""
    text = open(os.path.join(REFERENCE_ROOT, "Training_VAE.py"), encoding="utf-8").read()
    head = 'if __name__ == "__main__":\n'
    body = text[text.index(head) + len(head):]
    return textwrap.dedent(body)


def run_reference_training(ref, seed, data_path, epochs, batch_size, out_dir):
    """Executes the reference's training mode in the reference module's own namespace."""
    src = reference_main_block()
    literals = {
        "data_path": repr(data_path), "epochs": str(epochs), "batch_size": str(batch_size),
        "model_save_path": repr(os.path.join(out_dir, "model.pth")),
        "loss_save_path": repr(os.path.join(out_dir, "loss.png")),
    }
    for name, value in literals.items():
        src, n = re.subn(rf"^{name} = .*$", f"{name} = {value}", src, count=1, flags=re.M)
        assert n == 1, name
    ns = dict(vars(ref.Training_VAE))
    seen = {}

    def plot_losses(loss_history, epochs_arg, save_path):     # the real one needs matplotlib; record the call
        seen["plot"] = ({k: list(v) for k, v in loss_history.items()}, epochs_arg, save_path)

    ns["plot_losses"] = plot_losses
    ns["tqdm"] = lambda it, **kw: it
    torch.manual_seed(seed)
    exec(compile(src, "<reference Training_VAE.py __main__>", "exec"), ns)
    assert ns["mode"] == "training"
    hist, ep, path = seen["plot"]
    assert ep == epochs and path == literals["loss_save_path"].strip("'")
    sd = torch.load(os.path.join(out_dir, "model.pth"), map_location="cpu")
    return hist, sd


def frame_locals_of(exc, func_name):
    tb = exc.__traceback__
    while tb is not None:
        if tb.tb_frame.f_code.co_name == func_name:
            return tb.tb_frame.f_locals
        tb = tb.tb_next
    raise KeyError(func_name)


def main() -> None:
    torch.set_num_threads(1)
    ref = load_reference()
    data_path = os.path.join(GOLD, "data_sce1_cond.npy")

    # ---- Training_VAE.py training mode -------------------------------------------------------------------
    out = {"keys": np.array(["total_loss", "recon_loss", "kld_loss", "start_loss", "time_loss"])}
    for tag, seed, epochs, bs in (("b38", 2024, 12, 38), ("b16", 7, 6, 16)):
        with tempfile.TemporaryDirectory() as tmp:
            hist, sd = run_reference_training(ref, seed, data_path, epochs, bs, tmp)
        assert list(hist.keys()) == list(out["keys"])
        out[f"{tag}/seed"], out[f"{tag}/epochs"], out[f"{tag}/batch_size"] = np.array(seed), np.array(epochs), np.array(bs)
        out[f"{tag}/hist"] = np.array([hist[k] for k in out["keys"]], dtype=np.float64)      # (5, epochs), weight-scaled
        out[f"{tag}/state_keys"] = np.array(list(sd.keys()))
        for k, v in sd.items():
            out[f"{tag}/final/{k}"] = reduced(v)
            out[f"{tag}/final_digest/{k}"] = digest(v)
    np.savez(os.path.join(GOLD, "train_entry.npz"), **out)

    # ---- Tools.visualize_trajectories: the decode block ---------------------------------------------------
    ckpt = os.path.join(REFERENCE_ROOT, "training", "models", "vae_offset_sce1_cond_ld8_epoch3000.pth")
    model = ref.Training_VAE.ConditionalTrajectoryVAE(10, 3, 8)
    model.load_state_dict(torch.load(ckpt, map_location="cpu"))
    dataset = ref.Training_VAE.TrajectoryDataset(data_path)
    vis = {}
    for tag, seed, kwargs in (("train_starts", 31, dict(use_training_start_end=True, train_traj_start=3, train_traj_end=12)),
                              ("custom_start", 32, dict(use_training_start_end=False, custom_start_end=[(-194.0, 19.1), (0.0, 0.0)],
                                                        train_traj_start=0, train_traj_end=9))):
        torch.manual_seed(seed)
        try:
            ref.Tools.visualize_trajectories(model, dataset, "unused.pth", axis_flip="y", **kwargs)
        except Exception as e:  # noqa: BLE001 - the stubbed matplotlib ends the function after the decode block
            loc = frame_locals_of(e, "visualize_trajectories")
        else:
            raise AssertionError("visualize_trajectories was expected to stop at its plotting tail")
        vis[f"{tag}/seed"] = np.array(seed)
        vis[f"{tag}/train_data"] = np.asarray(loc["train_data"])
        vis[f"{tag}/generated"] = np.asarray(loc["generated_samples"])
        vis[f"{tag}/z"] = loc["z"].numpy()
        assert vis[f"{tag}/generated"].dtype == np.float32
    np.savez(os.path.join(GOLD, "visualize_entry.npz"), **vis)

    # ---- Driver_Models.Reg157 -----------------------------------------------------------------------------
    g = np.random.default_rng(157)
    args = np.concatenate([g.uniform([-50, 0.5, -50, 0], [50, 40, 50, 40], size=(200, 4)),
                           np.array([[0.0, 10.0, 30.0, 4.0], [0.0, 10.0, 1.0, 4.0], [5.0, 3.0, 9.0, 8.0], [0.0, 20.0, 23.0, 8.0]])])
    outv = []
    for x_e, v_e, x_f, v_f in args:
        if v_e == v_f:
            continue
        r = ref.Driver_Models.Reg157(x_e, v_e, x_f, v_f)
        outv.append(np.nan if r is None else float(r))
    np.savez(os.path.join(GOLD, "reg157.npz"), args=args, out=np.array(outv))
    print("wrote train_entry.npz, visualize_entry.npz, reg157.npz")


if __name__ == "__main__":
    main()
