"""Development aid: throughput of the batched MPC tracker.  python scripts/quick_mpc_bench.py [n] [steps]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "defensive-model-vae_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)
import bench  # noqa: E402  (synthetic trajectories of SURVEY.md 8d)
from dmvae.tracker import BatchTracker  # noqa: E402


def synthetic_jobs(n, seed=0, device="cuda"):
    """[x, y, t] float32 waypoints + [x, y, theta, vx, vy] initial states from the bench's synthetic trajectories."""
    traj = bench.synth_trajectories(n, seed, device)                  # (n, T, 3) [t, x, y]
    way = traj[:, :, [1, 2, 0]].contiguous()
    d = (way[:, 1] - way[:, 0]).double()
    vx, vy = d[:, 0] / d[:, 2], d[:, 1] / d[:, 2]
    init = torch.stack([way[:, 0, 0].double(), way[:, 0, 1].double(), torch.atan2(vy, vx), vx, vy], 1)
    return way, init


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
    K = int(sys.argv[2]) if len(sys.argv) > 2 else 50
    way, init = synthetic_jobs(n)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    bt = BatchTracker(way, init, 0.02, 30, 20)
    torch.cuda.synchronize()
    print(f"n={n}: prepare {1e3 * (time.perf_counter() - t0):.2f} ms; steps per trajectory min/mean/max "
          f"{bt.n_steps.min()}/{bt.n_steps.mean():.0f}/{bt.n_steps.max()}; untrackable {(bt.status != 0).sum().item()}")
    bt.advance(10)                                     # cold start of the solver (no previous solution)
    torch.cuda.synchronize()
    for rep in range(3):
        it0 = bt.iters.clone()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        bt.advance(K)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        active = int((bt.n_steps >= bt.step).sum())
        its = (bt.iters - it0).double().mean().item() / K
        print(f"  steps {bt.step - K}..{bt.step}: {ms:.2f} ms, {n * K / ms / 1e3:.2f} M controller calls/s, {its:.2f} solver iterations per call, "
              f"{active} of {n} trajectories still running")


if __name__ == "__main__":
    main()
