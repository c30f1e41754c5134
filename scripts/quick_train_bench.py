"""Development aid: step time of the fused training pass, tensor-core vs FFMA kernels.

    python scripts/quick_train_bench.py [B ...]
"""
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

T, L = 10, 8
FLOP = 758272


def profile(lib, fn, iters):
    n = _lib.KERNEL_COUNT
    ms = (ctypes.c_double * n)()
    cnt = (ctypes.c_int64 * n)()
    lib.dmvae_profile_begin()
    for i in range(iters):
        fn()
    torch.cuda.synchronize()
    lib.dmvae_profile_end(ms, cnt, n)
    return {lib.dmvae_kernel_name(i).decode(): ms[i] / cnt[i] * 1e3 for i in range(n) if cnt[i]}


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [4096, 18944, 65536]
    lib = _lib.lib()
    torch.manual_seed(0)
    model = ConditionalTrajectoryVAE(T, 3, L).to("cuda")
    for B in sizes:
        x = torch.randn(B, T, 3, device="cuda").cumsum(1)
        for impl, name in ((1, "FFMA     "), (2, "TC 2-launch"), (0, "TC       ")):
            lib.dmvae_set_train_impl(impl)
            tr = FusedTrainer(model, lr=1e-4)
            for _ in range(5):
                tr.step(x)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            K = 50
            e0.record()
            for _ in range(K):
                tr.step(x)
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) / K * 1e3
            prof = profile(lib, lambda: tr.step(x), 20)
            print(f"B={B:6d} {name}: {us:8.1f} us/step  {B / us:7.2f} M samples/s  {B * FLOP / us / 1e6:6.1f} TFLOP/s   kernels(us): "
                  + "  ".join(f"{k}={v:.1f}" for k, v in prof.items()), flush=True)


if __name__ == "__main__":
    main()
