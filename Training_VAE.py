"""Drop-in for the reference's ``Training_VAE.py`` on B200 kernels.

Same public names (``TrajectoryDataset``, ``ConditionalTrajectoryVAE``,
``conditional_vae_loss``), same ``__main__`` literals, save-name scheme and
loss-history layout (reference ``Training_VAE.py:105-115, :118-226, :229-268,
:273-431``).  The arithmetic runs in ``libdmvae.so`` (hand-written sm_100a
kernels, ``include/dmvae.h``); there is no CPU path - ``device = 'cpu'`` below is
kept because the reference hard-codes it (``:282``) and is accepted as a value,
but the model lives on the current CUDA device.

Unlike the reference (whose ``import Training_VAE`` fails unless ``Tools`` was
imported first - a circular import), both import orders work here.
"""
import os
import sys

import numpy as np
import torch
from torch.utils.data import DataLoader, Dataset

_PKG = os.path.join(os.path.dirname(os.path.abspath(__file__)), "defensive-model-vae_b200")
if _PKG not in sys.path:
    sys.path.insert(0, _PKG)

from dmvae.autograd import conditional_vae_loss  # noqa: E402,F401
from dmvae.model import ConditionalTrajectoryVAE  # noqa: E402,F401
from dmvae.train import LOSS_KEYS, FusedTrainer, LossMeter  # noqa: E402,F401


class TrajectoryDataset(Dataset):
    """``(N, seq_len, 3)`` float64 ``.npy`` of absolute ``[t, x, y]`` waypoints, cast
    once to float32 (reference ``Training_VAE.py:105-115``)."""

    def __init__(self, data_path):
        self.data = np.load(data_path).astype(np.float32)

    def __len__(self):
        return len(self.data)

    def __getitem__(self, idx):
        return self.data[idx]


def train(data_path, seq_len=10, dim=3, latent_dim=8, batch_size=38, lr=1e-3, epochs=3000, device='cpu',
          recon_weight=0.1, kld_weight=0.1, start_weight=1.0, time_weight=1.0, model_save_path=None,
          loss_save_path=None, verbose=True, philox_seed=None):
    """The reference's training mode (``Training_VAE.py:316-394``) on the fused
    step: per batch one ``FusedTrainer.step`` (offset transform, forward, loss,
    backward, Adam) instead of ~110 torch ops and five ``.item()`` syncs.

    Batches come from the same ``DataLoader(shuffle=True)`` as in the reference and
    the reparameterisation noise from ``torch.randn`` on the host generator, so a
    seeded run consumes the reference's RNG stream (set ``philox_seed`` to draw the
    noise in-kernel instead).  Returns (model, loss_history)."""
    dataset = TrajectoryDataset(data_path)
    dataloader = DataLoader(dataset, batch_size=batch_size, shuffle=True)
    if verbose:
        print(f"dataset: {len(dataset)} trajectories; seq_len={seq_len}, latent_dim={latent_dim}, "
              f"batch_size={batch_size}, lr={lr}")
    model = ConditionalTrajectoryVAE(seq_len, dim, latent_dim).to(device)
    trainer = FusedTrainer(model, lr=lr, weights=(recon_weight, kld_weight, start_weight, time_weight),
                           seed=0 if philox_seed is None else philox_seed)
    model.train()
    meter = LossMeter(trainer.device)
    loss_history = {k: [] for k in LOSS_KEYS}
    for epoch in range(epochs):
        for batch in dataloader:
            eps = None if philox_seed is not None else torch.randn(batch.shape[0], latent_dim)
            losses = trainer.step(batch, eps=eps)
            meter.update(losses, batch.shape[0])
        means = meter.means()  # the epoch's one device->host read
        if verbose:
            print(f"Epoch {epoch+1}: Loss={means[0]:.4f}, Recon={means[1]:.4f}, KLD={means[2]:.4f}, "
                  f"Start={means[3]:.4f}, Time={means[4]:.4f}")
        for k, v in zip(LOSS_KEYS, means):
            loss_history[k].append(v)
    # component histories are reported multiplied by their weights (reference :385-388)
    for k, w in (('recon_loss', recon_weight), ('kld_loss', kld_weight), ('start_loss', start_weight),
                 ('time_loss', time_weight)):
        loss_history[k] = [x * w for x in loss_history[k]]
    if loss_save_path is not None:
        from Tools import plot_losses
        plot_losses(loss_history, epochs, loss_save_path)
    if model_save_path is not None:
        os.makedirs(os.path.dirname(model_save_path) or ".", exist_ok=True)
        torch.save(model.state_dict(), model_save_path)
        if verbose:
            print(f"model saved to {model_save_path}")
    return model, loss_history


from Tools import *  # noqa: E402,F401,F403  (the reference star-imports Tools, Training_VAE.py:102)

if __name__ == "__main__":
    # ====== editable parameters: same literals as the reference (Training_VAE.py:273-313) ======
    mode = 'training'  # 'training', 'visualization'
    data_path = 'training/DefensiveDataProcessed/trajectory_sce1_cond.npy'
    seq_len = 10
    dim = 3
    latent_dim = 8
    batch_size = 38  # sce1 = 38, sce2 = 16, sce3 = 66, sce4 = 135
    lr = 1e-3
    epochs = 3000
    device = 'cpu'  # accepted for compatibility; the kernels run on the current CUDA device
    model_name = data_path.split('/')[-1].split('.')[0].replace("trajectory_", "", 1)
    model_save_path = 'training/models/vae_offset_' + model_name + '_ld' + str(latent_dim) + '_epoch' + str(epochs) + '_loss2.pth'
    loss_save_path = 'training/loss/vae_offset_' + model_name + '_ld' + str(latent_dim) + '_epoch' + str(epochs) + '_loss2.png'
    use_training_start_end = True
    custom_start_end = [(155.0, -15.0), (155.0, 40.0)]
    recon_weight = 0.1
    kld_weight = 0.1
    start_weight = 1.0
    time_weight = 1.0
    train_traj_start = 0
    train_traj_end = 9
    axis_flip = 'y'

    if mode == 'training':
        train(data_path, seq_len, dim, latent_dim, batch_size, lr, epochs, device, recon_weight, kld_weight,
              start_weight, time_weight, model_save_path, loss_save_path)
    elif mode == 'visualization':
        model = ConditionalTrajectoryVAE(seq_len, dim, latent_dim).to(device)
        model.load_state_dict(torch.load(model_save_path, map_location=device))
        dataset = TrajectoryDataset(data_path)
        visualize_trajectories(model, dataset, model_save_path, axis_flip=axis_flip,  # noqa: F405
                               use_training_start_end=use_training_start_end, custom_start_end=custom_start_end,
                               train_traj_start=train_traj_start, train_traj_end=train_traj_end)
    else:
        print("mode must be 'training' or 'visualization'")
