"""Development aid: per-tensor gradient errors of the fused train kernel vs the oracle."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "defensive-model-vae_b200"))
import numpy as np, torch
from oracle import vae_oracle as O
from dmvae import ConditionalTrajectoryVAE
from dmvae.train import FusedTrainer

def run(T, L, B, weights=O.SCRIPT_WEIGHTS):
    p = O.init_params(T, L, seed=7)
    m = ConditionalTrajectoryVAE(T, 3, L); m.load_state_dict({k: v.clone() for k, v in p.items()}); m.to("cuda")
    g = torch.Generator().manual_seed(B)
    t = torch.cumsum(torch.rand(B, T, generator=g) * 1.5 + 0.3, 1); t = t - t[:, :1]
    xy = torch.cumsum(torch.randn(B, T, 2, generator=g), 1) + (torch.rand(B, 1, 2, generator=g) - 0.5) * 100
    batch = torch.cat([t[..., None], xy], -1).contiguous()
    eps = torch.randn(B, L, generator=g)
    lr, gr, keep = O.loss_and_grads(p, batch, eps, weights)
    tr = FusedTrainer(m, weights=weights)
    losses, grads = tr.loss_and_grads(batch.cuda(), eps=eps.cuda())
    torch.cuda.synchronize()
    print(f"--- T={T} L={L} B={B}: losses got {[round(float(v),6) for v in losses.cpu()]} ref {[round(v,6) for v in lr]}")
    gnp = grads.cpu().numpy(); off = 0
    gmax = max(float(v.abs().max()) for v in gr.values())
    for k, v in gr.items():
        n = v.numel(); mine = gnp[off:off+n]; ref = v.reshape(-1).numpy(); off += n
        err = np.abs(mine - ref).max(); sc = np.abs(ref).max()
        flag = "" if err <= 2e-5 * gmax + 2e-4 * sc else "  <<<<<<"
        print(f"   {k:28s} |ref|max {sc:10.3e}  abs err {err:10.3e}  rel {err/max(sc,1e-30):9.2e}  nan={np.isnan(mine).any()}{flag}")

for cfg in [(10, 8, 38), (10, 8, 300), (10, 8, 9500), (42, 64, 70), (10, 5, 64)]:
    run(*cfg)
print("debug done")
