"""Development aid: the tensor-core training pass stage by stage against the oracle.

    python tests/dev/debug_train_tc.py [B] [T] [L]

Runs dmvae_train_fwd_bwd with the tensor-core kernels, reads the stash back from the workspace,
and reports (1) every stashed layer input X against the oracle's activations, (2) the weight
gradients recomputed on the host as G^T X from the stashed images (isolates the chain kernel),
(3) the gradients / losses the kernels returned, against the oracle and the FFMA kernels.
"""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "defensive-model-vae_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

from dmvae import ConditionalTrajectoryVAE, _lib  # noqa: E402
from dmvae.train import FusedTrainer  # noqa: E402
from oracle import vae_oracle as O  # noqa: E402

SLOTS = ["SX_START", "SX_HC1", "SX_HC", "SX_X", "SX_E1", "SX_E2", "SX_E3", "SX_E4", "SX_Z", "SX_D1", "SX_D2", "SX_D3",
         "SG_REC", "SG_D3", "SG_D2", "SG_D1", "SG_ML", "SG_HC", "SG_HC1", "SG_E4", "SG_E3", "SG_E2", "SG_E1"]


def pad_width(n):
    return 32 if n <= 32 else (64 if n <= 64 else 128)


def rel(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    T = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    L = int(sys.argv[3]) if len(sys.argv) > 3 else 8
    I, Ip = 3 * T, pad_width(3 * T)
    Lp16 = (L + 15) // 16 * 16
    NH = 16 if 2 * L <= 16 else pad_width(2 * L)
    widths = {s: 128 for s in SLOTS}
    widths.update(SX_START=16, SX_X=Ip, SG_REC=Ip, SX_Z=Lp16, SG_ML=NH)
    offs, o = {}, 0
    wmem = {s: (widths[s] + 31) // 32 * 32 for s in SLOTS}
    for s in SLOTS:
        offs[s] = o
        o += 128 * wmem[s]
    tile_stash = o

    p = O.init_params(T, L, seed=0)
    model = ConditionalTrajectoryVAE(T, 3, L)
    model.load_state_dict({k: v.clone() for k, v in p.items()})
    model.to("cuda")
    g = torch.Generator().manual_seed(1)
    start = torch.rand(B, 2, generator=g) * 200 - 100
    t = torch.cumsum(torch.rand(B, T, generator=g) + 0.3, 1) - 0.3
    xy = torch.cumsum(torch.randn(B, T, 2, generator=g), 1) + start[:, None, :]
    batch = torch.cat([t[..., None], xy], -1).contiguous()
    eps = torch.randn(B, L, generator=g)
    losses_ref, grads_ref, keep = O.loss_and_grads(p, batch, eps, O.SCRIPT_WEIGHTS)
    gref = torch.cat([v.reshape(-1) for v in grads_ref.values()]).numpy()

    lib = _lib.lib()
    trainer = FusedTrainer(model, lr=1e-3, weights=O.SCRIPT_WEIGHTS)
    lib.dmvae_set_train_impl(1)
    l1, g1 = trainer.loss_and_grads(batch.cuda(), eps=eps.cuda())
    l1, g1 = l1.cpu().numpy().copy(), g1.cpu().numpy().copy()
    print(f"FFMA  : grad rel {rel(g1, gref):.2e}  losses {l1}  ref {np.array(losses_ref)}")
    lib.dmvae_set_train_impl(0)
    l0, g0 = trainer.loss_and_grads(batch.cuda(), eps=eps.cuda())
    torch.cuda.synchronize()
    l0, g0 = l0.cpu().numpy().copy(), g0.cpu().numpy().copy()
    print(f"TC    : grad rel {rel(g0, gref):.2e}  losses {l0}")

    ws = trainer._ws[B].view(torch.float32).cpu().numpy()
    n_tiles = (B + 127) // 128

    def slot(name):
        w, wm = widths[name], wmem[name]
        r = np.arange(128)[:, None]
        f = np.arange(w)[None, :]
        # MN-major image, 32-byte swizzle (dmvae_tc.cuh: mn_image_index(f, r, 128, wm * 4))
        idx = (f >> 5) * 128 + (r >> 2) * (wm * 4) + (r & 3) * 32 + ((((f >> 3) & 3) ^ (r & 3)) << 3) + (f & 7)
        out = np.zeros((n_tiles * 128, w), np.float32)
        for ti in range(n_tiles):
            img = ws[ti * tile_stash + offs[name]: ti * tile_stash + offs[name] + 128 * wm]
            out[ti * 128:(ti + 1) * 128] = img[idx]
        return out[:B]

    hc1 = torch.relu(torch.nn.functional.linear(keep["start"], p["condition_encoder.0.weight"], p["condition_encoder.0.bias"]))
    refs = {"SX_HC1": hc1, "SX_HC": keep["h_c"], "SX_X": keep["x_rel"].reshape(B, -1), "SX_E1": keep["enc1"],
            "SX_E2": keep["enc3"], "SX_E3": keep["enc5"], "SX_E4": keep["enc7"], "SX_Z": keep["z"], "SX_D1": keep["dec0"],
            "SX_D2": keep["dec2"], "SX_D3": keep["dec4"]}
    for name, r in refs.items():
        got = slot(name)[:, : r.shape[1]]
        print(f"  {name:8s} rel {rel(got, r.numpy()):.2e}")

    # weight gradients from the stash, on the host (float64)
    X = {k: slot(k).astype(np.float64) for k in SLOTS}
    def chk(name, got, key):
        r = grads_ref[key].numpy()
        print(f"  host G^T X {name:10s} rel {rel(got, r):.2e}   kernel {rel(g0[off_of[key]:off_of[key] + r.size].reshape(r.shape), r):.2e}")
    off_of, o = {}, 0
    for k, v in grads_ref.items():
        off_of[k] = o
        o += v.numel()
    chk("cond0.w", (X["SG_HC1"].T @ X["SX_START"])[:, :2], "condition_encoder.0.weight")
    chk("cond0.b", X["SG_HC1"].sum(0), "condition_encoder.0.bias")
    chk("cond1.w", X["SG_HC"].T @ X["SX_HC1"], "condition_encoder.2.weight")
    chk("cond1.b", X["SG_HC"].sum(0), "condition_encoder.2.bias")
    chk("enc0.w", (X["SG_E1"].T @ X["SX_X"])[:, :I], "encoder.1.weight")
    chk("enc0.b", X["SG_E1"].sum(0), "encoder.1.bias")
    chk("enc1.w", X["SG_E2"].T @ X["SX_E1"], "encoder.3.weight")
    chk("enc2.w", X["SG_E3"].T @ X["SX_E2"], "encoder.5.weight")
    chk("enc3.w", X["SG_E4"].T @ X["SX_E3"], "encoder.7.weight")
    chk("enc3.b", X["SG_E4"].sum(0), "encoder.7.bias")
    ml = X["SG_ML"]
    chk("fc_mu.w", np.concatenate([ml[:, :L].T @ X["SX_E4"], ml[:, :L].T @ X["SX_HC"]], 1), "fc_mu.weight")
    chk("fc_mu.b", ml[:, :L].sum(0), "fc_mu.bias")
    chk("fc_lv.w", np.concatenate([ml[:, L:2 * L].T @ X["SX_E4"], ml[:, L:2 * L].T @ X["SX_HC"]], 1), "fc_logvar.weight")
    chk("fc_lv.b", ml[:, L:2 * L].sum(0), "fc_logvar.bias")
    chk("dec0.w", np.concatenate([(X["SG_D1"].T @ X["SX_Z"])[:, :L], X["SG_D1"].T @ X["SX_HC"]], 1), "decoder.0.weight")
    chk("dec0.b", X["SG_D1"].sum(0), "decoder.0.bias")
    chk("dec1.w", X["SG_D2"].T @ X["SX_D1"], "decoder.2.weight")
    chk("dec2.w", X["SG_D3"].T @ X["SX_D2"], "decoder.4.weight")
    chk("dec2.b", X["SG_D3"].sum(0), "decoder.4.bias")
    chk("dec3.w", (X["SG_REC"].T @ X["SX_D3"])[:I], "decoder.6.weight")
    chk("dec3.b", X["SG_REC"].sum(0)[:I], "decoder.6.bias")
    print("losses TC", l0, "ref", losses_ref)


if __name__ == "__main__":
    main()
