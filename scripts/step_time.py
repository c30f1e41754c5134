"""Development aid: per-kernel time of a training step (event pairs around every launch, dmvae_profile_*) and the
graph-replayed step time at a few batch sizes.  python scripts/step_time.py [B ...]"""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "defensive-model-vae_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)
from dmvae import ConditionalTrajectoryVAE, _lib  # noqa: E402
from dmvae.train import FusedTrainer  # noqa: E402

lib = _lib.lib()
IMPL = int(os.environ.get("IMPL", "0"))     # dmvae_set_train_impl: 0 default, 2 tensor cores with two launches, 3 tensor cores always
_lib.check(lib.dmvae_set_train_impl(IMPL), "dmvae_set_train_impl")
T, L = int(os.environ.get("T", "10")), int(os.environ.get("L", "8"))      # trajectory length, latent dim
sizes = [int(a) for a in sys.argv[1:]] or [4096, 65536]
for B in sizes:
    torch.manual_seed(0)
    model = ConditionalTrajectoryVAE(T, 3, L).to("cuda")
    tr = FusedTrainer(model, lr=1e-4)
    x = torch.randn(B, T, 3, device="cuda").cumsum(1)
    for _ in range(5):
        tr.step(x)
    n = _lib.KERNEL_COUNT
    ms = (ctypes.c_double * n)()
    cnt = (ctypes.c_int64 * n)()
    _lib.check(lib.dmvae_profile_begin(), "begin")
    iters = 50
    for _ in range(iters):
        tr.step(x)
    torch.cuda.synchronize()
    _lib.check(lib.dmvae_profile_end(ms, cnt, n), "end")
    per = {lib.dmvae_kernel_name(i).decode(): ms[i] / cnt[i] * 1e3 for i in range(n) if cnt[i]}
    try:
        gs = tr.capture(B)
        gs.batch.copy_(x)
        run = gs.replay
    except Exception:            # outside the tensor-core envelope there is no device-side step counter: host-driven steps
        run = lambda: tr.step(x)
    for _ in range(20):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(200):
        run()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 200 * 1e3
    print(f"T={T} L={L} B={B}: step {us:.2f} us ({B / us:.2f} M samples/s)  kernels: " + ", ".join(f"{k} {v:.2f}" for k, v in per.items()), flush=True)
