"""Generate the golden fixtures under ``tests/golden`` by RUNNING THE REFERENCE.

TEST INFRASTRUCTURE.  Run once in the build container (where
``/root/reference`` is mounted):

    python -m oracle.make_golden

Everything written here is produced by the reference's own classes
(``Training_VAE.ConditionalTrajectoryVAE``, ``conditional_vae_loss``,
``torch.optim.Adam``, ``Tools.load_model_and_generate_trajectory``) through
``oracle.ref_loader`` - never by the restatement in ``oracle.vae_oracle`` and
never by the CUDA path.  The fixtures travel to the GPU box (the reference does
not) and pin both the restatement and the kernels.

Fixtures:
  ckpt_sce1_cond.npz, ckpt_sce4_cond.npz  weights of two shipped checkpoints
        (training/models/vae_offset_sce{1,4}_cond_ld8_epoch3000.pth), fp32.
  data_sce1_cond.npy   the shipped 38x10x3 float64 dataset
        (training/DefensiveDataProcessed/trajectory_sce1_cond.npy).
  decode_kat.npz       fixed-latent decode known answers on those checkpoints.
  generate_api.npz     Tools.load_model_and_generate_trajectory under a seed.
  init_seed.npz        per-tensor digests of the seeded default initialisation.
  train_sce1.npz       50 reference train steps (seed-0 init, real sce1 data,
        injected eps): loss history, step-0 gradients, final parameter digests.
  train_small.npz      5 steps of a T=12, L=4, B=16 synthetic configuration.
  loss_kat.npz         conditional_vae_loss on fixed tensors, both weight sets.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)

from oracle.ref_loader import REFERENCE_ROOT, load_reference  # noqa: E402

BIG = 128 * 128


def digest(t: torch.Tensor) -> np.ndarray:
    """(sum, abs-sum, first, last) in float64 - enough to detect any drift."""
    d = t.detach().double().reshape(-1)
    return np.array([d.sum().item(), d.abs().sum().item(), d[0].item(), d[-1].item()], dtype=np.float64)


def reduced(t: torch.Tensor) -> np.ndarray:
    """Full tensor when small; the first four rows when it is a 128x128 block."""
    a = t.detach().cpu().numpy()
    return a[:4].copy() if a.size >= BIG else a.copy()


class InjectedNoise:
    """Replace torch.randn_like inside the reference's reparameterize
    (Training_VAE.py:205) by a queue of pre-drawn tensors."""

    def __init__(self, queue):
        self.queue = list(queue)
        self._orig = None

    def __enter__(self):
        self._orig = torch.randn_like

        def fake(t, *a, **k):
            e = self.queue.pop(0)
            assert e.shape == t.shape, (e.shape, t.shape)
            return e

        torch.randn_like = fake
        return self

    def __exit__(self, *exc):
        torch.randn_like = self._orig


def reference_train(ref, model, batch, eps_steps, weights, lr=1e-3):
    """Training_VAE.py:338-363 for a full-batch loader (one step per epoch)."""
    opt = torch.optim.Adam(model.parameters(), lr=lr)
    model.train()
    hist, grads0 = [], None
    with InjectedNoise(eps_steps):
        for s in range(len(eps_steps)):
            start_points = batch[:, 0, 1:3]
            batch_rel = batch.clone()
            batch_rel[:, :, 1:3] = batch_rel[:, :, 1:3] - start_points.unsqueeze(1)
            opt.zero_grad()
            recon, mu, logvar, cond = model(batch_rel, start_points)
            out = ref.Training_VAE.conditional_vae_loss(
                recon, batch_rel, mu, logvar, cond, recon_weight=weights[0], kld_weight=weights[1],
                start_weight=weights[2], time_weight=weights[3])
            out[0].backward()
            if s == 0:
                grads0 = {k: v.grad.detach().clone() for k, v in model.named_parameters()}
            opt.step()
            hist.append([float(o) for o in out])
    return np.array(hist, dtype=np.float64), grads0


def main() -> None:
    torch.set_num_threads(1)  # one summation order for the record
    os.makedirs(GOLD, exist_ok=True)
    ref = load_reference()
    VAE = ref.Training_VAE.ConditionalTrajectoryVAE

    # ---- shipped checkpoints + dataset as fixtures ---------------------------------
    ckpts = {}
    for sce in ("sce1", "sce4"):
        path = os.path.join(REFERENCE_ROOT, "training", "models", f"vae_offset_{sce}_cond_ld8_epoch3000.pth")
        sd = torch.load(path, map_location="cpu")
        assert all(v.dtype == torch.float32 for v in sd.values())
        np.savez(os.path.join(GOLD, f"ckpt_{sce}_cond.npz"), **{k: v.numpy() for k, v in sd.items()})
        ckpts[sce] = sd
    data = np.load(os.path.join(REFERENCE_ROOT, "training", "DefensiveDataProcessed", "trajectory_sce1_cond.npy"))
    assert data.shape == (38, 10, 3) and data.dtype == np.float64
    np.save(os.path.join(GOLD, "data_sce1_cond.npy"), data)

    # ---- fixed-latent decode KATs (Tools.py:898-912 semantics, batched, pure fp32) -----
    g = torch.Generator().manual_seed(1234)
    z = torch.randn(16, 8, generator=g)
    z[0] = 0.0
    kat = {"z": z.numpy()}
    starts = {
        "sce1": torch.tensor(data[:16, 0, 1:3]).float(),                       # per-row starts from the dataset
        "sce4": torch.tensor([[11.0, 0.0]] * 16, dtype=torch.float32),        # shared default start (Tools.py:106)
    }
    for sce, sd in ckpts.items():
        model = VAE(10, 3, 8)
        model.load_state_dict(sd)
        model.eval()
        with torch.no_grad():
            sp = starts[sce]
            h = model.condition_encoder(sp)
            rel = model.decode(z, h).cpu().numpy()
            sp_np = sp.cpu().numpy()
            glob = rel.copy()
            for i in range(rel.shape[0]):
                glob[i, :, 1] = sp_np[i, 0] + rel[i, :, 1]
                glob[i, :, 2] = sp_np[i, 1] + rel[i, :, 2]
        assert glob.dtype == np.float32
        kat[f"{sce}_start"] = sp_np
        kat[f"{sce}_rel"] = rel
        kat[f"{sce}_global"] = glob
        kat[f"{sce}_h_c"] = h.numpy()
    np.savez(os.path.join(GOLD, "decode_kat.npz"), **kat)

    # ---- the single-trajectory API under a seed (Tools.py:18-65) ---------------------
    path = os.path.join(REFERENCE_ROOT, "training", "models", "vae_offset_sce1_cond_ld8_epoch3000.pth")
    api = {}
    for seed, (sx, sy) in ((123, (-194.25, 19.0)), (7, (np.float32(-193.77), np.float32(18.82)))):
        torch.manual_seed(seed)
        traj = ref.Tools.load_model_and_generate_trajectory(path, sx, sy, seq_len=10, dim=3, latent_dim=8, device="cpu")
        assert traj.shape == (10, 3) and traj.dtype == np.float32
        api[f"seed{seed}"] = traj
        api[f"seed{seed}_start"] = np.array([sx, sy], dtype=np.float64)
    np.savez(os.path.join(GOLD, "generate_api.npz"), **api)

    # ---- seeded default init digests (constructor order, Training_VAE.py:124-167) ----
    init = {}
    for (T, L, seed) in ((10, 8, 0), (12, 8, 5), (50, 16, 1)):
        torch.manual_seed(seed)
        m = VAE(T, 3, L)
        for k, v in m.state_dict().items():
            init[f"T{T}_L{L}_s{seed}/{k}"] = digest(v)
    np.savez(os.path.join(GOLD, "init_seed.npz"), **init)

    # ---- 50 reference train steps on the real sce1 data --------------------------------
    weights = (0.1, 0.1, 1.0, 1.0)  # Training_VAE.py:300-306
    torch.manual_seed(0)
    model = VAE(10, 3, 8)
    batch = torch.from_numpy(data.astype(np.float32))
    g = torch.Generator().manual_seed(99)
    eps = torch.randn(50, 38, 8, generator=g)
    hist, grads0 = reference_train(ref, model, batch, list(eps), weights)
    out = {"eps_seed": np.array(99), "init_seed": np.array(0), "weights": np.array(weights), "loss_hist": hist}
    for k, v in grads0.items():
        out[f"grad0/{k}"] = reduced(v)
        out[f"grad0_digest/{k}"] = digest(v)
    for k, v in model.state_dict().items():
        out[f"final/{k}"] = reduced(v)
        out[f"final_digest/{k}"] = digest(v)
    np.savez(os.path.join(GOLD, "train_sce1.npz"), **out)

    # ---- a small non-default shape: T=12, L=4, B=16, function-default weights ---------
    weights2 = (0.1, 0.1, 1.0, 0.5)
    torch.manual_seed(3)
    model = VAE(12, 3, 4)
    g = torch.Generator().manual_seed(17)
    t = torch.cumsum(torch.rand(16, 12, generator=g) * 0.8 + 0.1, dim=1)
    xy = torch.cumsum(torch.randn(16, 12, 2, generator=g), dim=1) + torch.tensor([12.0, -30.0])
    batch2 = torch.cat([t.unsqueeze(-1), xy], dim=-1).contiguous()
    eps2 = torch.randn(5, 16, 4, generator=g)
    hist2, grads02 = reference_train(ref, model, batch2, list(eps2), weights2)
    out = {"batch": batch2.numpy(), "eps": eps2.numpy(), "init_seed": np.array(3), "weights": np.array(weights2),
           "loss_hist": hist2}
    for k, v in grads02.items():
        out[f"grad0/{k}"] = reduced(v)
        out[f"grad0_digest/{k}"] = digest(v)
    for k, v in model.state_dict().items():
        out[f"final/{k}"] = reduced(v)
        out[f"final_digest/{k}"] = digest(v)
    np.savez(os.path.join(GOLD, "train_small.npz"), **out)

    # ---- the loss alone on fixed tensors, incl. a non-monotone time column -------------
    g = torch.Generator().manual_seed(5)
    x = torch.randn(9, 10, 3, generator=g)
    r = x + 0.3 * torch.randn(9, 10, 3, generator=g)
    r[:, :, 0] = torch.randn(9, 10, generator=g)  # time differences of both signs
    r[2, 4, 0] = r[2, 3, 0]  # an exactly-zero difference: relu'(0) = 0
    mu = torch.randn(9, 8, generator=g)
    lv = torch.randn(9, 8, generator=g)
    lk = {"x": x.numpy(), "recon": r.numpy(), "mu": mu.numpy(), "logvar": lv.numpy()}
    for name, w in (("script", weights), ("default", weights2)):
        rr, mm, ll = (t.clone().requires_grad_(True) for t in (r, mu, lv))
        o = ref.Training_VAE.conditional_vae_loss(rr, x, mm, ll, None, *w)
        o[0].backward()
        lk[f"{name}_losses"] = np.array([float(v) for v in o], dtype=np.float64)
        lk[f"{name}_g_recon"] = rr.grad.numpy()
        lk[f"{name}_g_mu"] = mm.grad.numpy()
        lk[f"{name}_g_logvar"] = ll.grad.numpy()
    np.savez(os.path.join(GOLD, "loss_kat.npz"), **lk)

    total = sum(os.path.getsize(os.path.join(GOLD, f)) for f in os.listdir(GOLD))
    print(f"wrote {len(os.listdir(GOLD))} fixtures, {total/1e6:.2f} MB, to {GOLD}")


if __name__ == "__main__":
    main()
