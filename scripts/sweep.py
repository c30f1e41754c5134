"""BASELINE configs[4]: scaling sweep over latent dim, trajectory length and batch on one GPU.

    python scripts/sweep.py [--out gpurun_out/sweep.jsonl] [--quick]

For every (seq_len T, latent L, batch B): decoded trajectories/s (shared start, in-kernel Philox) and
training samples/s (fused step: offset transform + forward + loss + backward + Adam), CUDA-event timed,
inputs resident in HBM; with the achieved FLOP rate (SURVEY.md 8d formulas) and which kernel family ran
(tensor cores inside their envelope, FP32 FFMA outside it).
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "defensive-model-vae_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)
from dmvae import ConditionalTrajectoryVAE, _lib  # noqa: E402
from dmvae.train import FusedTrainer  # noqa: E402

H = 128


def flops(T, L):
    I = 3 * T
    cond = 2 * H + H * H
    enc = I * H + 3 * H * H
    heads = 4 * H * L
    dec = (L + H) * H + 2 * H * H + I * H
    fwd = 2 * (cond + enc + heads + dec)
    return {"decode_shared": 2 * dec, "train": 3 * fwd - 2 * (2 * H + I * H)}


def timed(fn, reps):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "sweep.jsonl"))
    ap.add_argument("--quick", action="store_true")
    args = ap.parse_args()
    lib = _lib.lib()
    Ts = [10, 50, 100, 200, 400]
    Ls = [8, 16, 32, 64]
    Bs = [1 << 10, 1 << 14, 1 << 17, 1 << 20]
    if args.quick:
        Ts, Ls, Bs = [10, 50], [8, 64], [1 << 10, 1 << 14]
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    out = open(args.out, "w")
    names = [lib.dmvae_kernel_name(i).decode() for i in range(_lib.KERNEL_COUNT)]
    print(f"{'T':>4} {'L':>3} {'B':>8} | {'decode traj/s':>14} {'TFLOP/s':>8} {'kernel':>18} | {'train samples/s':>16} {'TFLOP/s':>8} {'kernel':>22}")
    for T in Ts:
        for L in Ls:
            torch.manual_seed(0)
            model = ConditionalTrajectoryVAE(T, 3, L).to("cuda")
            fl = flops(T, L)
            start = torch.tensor([[11.0, 0.0]])
            for B in Bs:
                train_B = min(B, 1 << 17) if T >= 200 else B          # keeps the (B, T, 3) input under 1.3 GB
                reps = 3 if B >= (1 << 17) else 10
                c0 = [lib.dmvae_launch_count(i) for i in range(_lib.KERNEL_COUNT)]
                dt = timed(lambda: model.generate(start, n=B, seed=1), reps)
                c1 = [lib.dmvae_launch_count(i) for i in range(_lib.KERNEL_COUNT)]
                dk = "+".join(n for n, a, b in zip(names, c0, c1) if b > a and n != "pack_kernel")
                x = torch.randn(train_B, T, 3, device="cuda").cumsum(1)
                tr = FusedTrainer(model, lr=1e-4)
                tt = timed(lambda: tr.step(x), reps)
                c2 = [lib.dmvae_launch_count(i) for i in range(_lib.KERNEL_COUNT)]
                tk = "+".join(n for n, a, b in zip(names, c1, c2) if b > a and n != "pack_kernel")
                del tr, x
                row = {"seq_len": T, "latent_dim": L, "batch": B, "decode_traj_per_s": B / dt,
                       "decode_tflops": B * fl["decode_shared"] / dt / 1e12, "decode_kernels": dk,
                       "train_batch": train_B, "train_samples_per_s": train_B / tt,
                       "train_tflops": train_B * fl["train"] / tt / 1e12, "train_kernels": tk}
                out.write(json.dumps(row) + "\n")
                out.flush()
                print(f"{T:4d} {L:3d} {B:8d} | {B / dt:14.4g} {row['decode_tflops']:8.1f} {dk:>18} | "
                      f"{train_B / tt:16.4g} {row['train_tflops']:8.1f} {tk:>22}", flush=True)
            del model
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
