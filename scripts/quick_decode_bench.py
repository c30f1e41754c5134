"""Scratch timing of the two decode kernels (development aid, not the bench contract)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "defensive-model-vae_b200"))
import torch
from dmvae import ConditionalTrajectoryVAE, _lib

torch.manual_seed(0)
m = ConditionalTrajectoryVAE(10, 3, 8).to("cuda").eval()
lib = _lib.lib()
# accuracy of the tensor-core kernel against the FFMA kernel on the same latents
B = 1 << 16
z = torch.randn(B, 8, device="cuda")
for mode in ("shared", "per-row"):
    start = torch.rand(1 if mode == "shared" else B, 2, device="cuda") * 300 - 150
    lib.dmvae_set_decode_impl(1); ref = m.generate(start, z=z).clone()
    lib.dmvae_set_decode_impl(0); got = m.generate(start, z=z).clone()
    torch.cuda.synchronize()
    err = (got - ref).abs().max().item() / ref.abs().max().item()
    rel = (got - start.view(-1, 1, 2).expand(-1, 10, 2).reshape(-1, 10, 2).new_zeros(1)).abs().max().item()
    print(f"{mode}: tc vs ffma max|diff|/max|ref| = {err:.3e}", flush=True)
for impl in (0, 1):
    lib.dmvae_set_decode_impl(impl)
    for B in (4096, 148 * 128, 1 << 17, 1 << 20, 4 << 20):
        for mode in ("philox-shared", "philox-per-row", "z-shared"):
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
            print(f"impl={'tc' if impl == 0 else 'ffma'} B={B:8d} {mode:15s} {ms*1e3:9.1f} us  {B/ms/1e3:8.2f} M traj/s  {B*fl/ms/1e9:7.2f} TFLOP/s", flush=True)
