"""Scratch timing of the decode kernel (development aid, not the bench contract)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "defensive-model-vae_b200"))
import torch
from dmvae import ConditionalTrajectoryVAE

torch.manual_seed(0)
m = ConditionalTrajectoryVAE(10, 3, 8).to("cuda").eval()
for B in (4096, 148 * 128, 1 << 17, 1 << 20, 4 << 20):
    for mode in ("per-row", "shared", "philox-shared", "philox-per-row"):
        start = (torch.rand(B if "per-row" in mode else 1, 2, device="cuda") * 100)
        z = None if "philox" in mode else torch.randn(B, 8, device="cuda")
        out = torch.empty(B, 10, 3, device="cuda")
        for _ in range(3):
            m.generate(start, z=z, n=B, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        it = 10
        e0.record()
        for _ in range(it):
            m.generate(start, z=z, n=B, out=out)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / it
        fl = 141312 if "per-row" in mode else 75264
        print(f"B={B:8d} {mode:15s} {ms*1e3:9.1f} us  {B/ms/1e3:8.2f} M traj/s  {B*fl/ms/1e9:7.2f} TFLOP/s", flush=True)
