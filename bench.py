#!/usr/bin/env python
"""bench.py - the trajectory-VAE hot path on B200 (driver contract: one JSON line).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's PyTorch-CPU path (port)

Workload (BASELINE.json configs[1]): the conditional trajectory VAE,
seq_len 10, latent 8, hidden 128, all four scenarios jointly, batch 4096 per GPU; one "step"
= relative-offset transform + forward + 5-term loss + backward + Adam on one batch
(reference Training_VAE.py:345-363).  N GPUs = data parallel, one SUM all-reduce of the flat
gradient buffer per step (weak scaling: 4096 rows per GPU per step).

  value      training samples/s, whole job, batches already resident in HBM
  e2e        the same through the public API with the batches in pinned HOST memory: every step
             uploads its batch (on a copy stream, while the previous step computes) and the host
             reads its five loss terms before it launches the next step
  roofline   the dominant kernel of the step (chain_kernel: forward + loss + data-gradient
             chain on the tcgen05 tensor cores): its algorithmic FLOP / its live CUDA-event
             duration, against the measured dense bf16 tensor peak of MEASURED_PEAKS.json.
             The path computes in fp32-equivalent 3xTF32 (TF32 runs at half the bf16 rate and
             every product takes three passes), so the ceiling of this arithmetic is peak / 6:
             reported as tf32x3_ceiling / frac_of_tf32x3_ceiling.  The FP32 FFMA rate measured
             by dmvae_ffma_probe in this run and the FFMA kernels' step rate are beside it.
  decode     second half of the metric (configs[2]): 4 scenarios x 2^20 latents per GPU,
             in-kernel Philox, shared scenario start -> (rows, 10, 3) fp32; same sub-keys
  cpu_baseline  the oracle (port of the reference's PyTorch-CPU path) on this host's cores

Inputs are synthetic (no dataset/checkpoint can be fetched): trajectories drawn from the
shipped per-scenario start boxes, seed 0, default-initialised weights, seed 0.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "defensive-model-vae_b200")
for _p in (ROOT, PKG):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import torch  # noqa: E402

T, D, L, H = 10, 3, 8, 128
WEIGHTS = (0.1, 0.1, 1.0, 1.0)            # Training_VAE.py:300-306
LR = 1e-3                                 # Training_VAE.py:279
METRIC = "vae_train_samples_per_sec"
UNIT = "samples/s"
# the same description in both arms (the driver compares the `config` objects of the two lines)
WORKLOAD = ("configs[1]: four scenarios jointly, batch 4096 per GPU, full train step (offset transform + forward + "
            "5-term loss + backward + Adam), seq_len 10, latent 8, hidden 128")


def workload_config(batch_per_gpu: int) -> dict:
    return {"workload": WORKLOAD, "batch_per_gpu": batch_per_gpu, "seq_len": T, "latent_dim": L, "hidden_dim": H}

# per-scenario start boxes of the shipped datasets and direction of travel
SCENARIOS = (
    # x_lo,    x_hi,    y_lo,   y_hi,  axis, sign
    (-195.71, -193.77, 18.82, 19.40, 1, +1.0),    # sce1 StaticBlindTown05
    (-150.63, -108.70, -3.41, 0.78, 0, -1.0),     # sce2 DynamicBlindTown05
    (153.83, 156.25, 39.32, 39.70, 1, -1.0),      # sce3 PredictableMovementTown05
    (13.17, 16.73, -45.39, 107.30, 1, -1.0),      # sce4 UnpredictableMovementTown04
)
SCENARIO_DEFAULT_START = ((-193.3, 50.0), (-155.0, -5.0), (155.0, -15.0), (11.0, 0.0))  # Tools.py:101-108


_REAL_STDOUT = None


def claim_stdout():
    """stdout carries exactly ONE JSON line (driver contract): everything else that native libraries print
    there (NCCL's version banner, for one) is sent to stderr for the life of the process."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line: dict) -> None:
    out = _REAL_STDOUT if _REAL_STDOUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def traffic_from_profile(kernel: str):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of `kernel`, from the newest committed
    `ncu --set full` summary under profiles/ (scripts/ncu_summary.py writes them); None when there is none."""
    import glob
    import re
    unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "*_prof_*.txt")), reverse=True):
        if path.endswith("_long.txt"):      # captures of the long-trajectory instantiations (T = 100): another workload
            continue
        try:
            text = open(path).read()
        except OSError:
            continue
        if not re.search(r"^--- .*\b" + re.escape(kernel) + r"\b", text, flags=re.M):
            continue
        total, seen = 0.0, 0
        for name in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            m = re.search(name + r"\s+([0-9.]+)\s+(\w+)", text)
            if m and m.group(2) in unit:
                total += float(m.group(1)) * unit[m.group(2)]
                seen += 1
        if seen == 2:
            return {"bytes_per_launch": total, "source": os.path.relpath(path, ROOT)}
    return None


def flops_per_unit():
    I = 3 * T
    cond = 2 * H + H * H
    enc = I * H + 3 * H * H
    heads = 4 * H * L
    dec = (L + H) * H + 2 * H * H + I * H
    fwd = cond + enc + heads + dec
    return {"train": 2 * (3 * fwd - (2 * H + I * H)), "decode_per_row_start": 2 * (cond + dec),
            "decode_shared_start": 2 * (dec - H * H),
            # the three GEMM families of a training pass: forward, data gradient (none for the two input
            # layers), weight gradient
            "train_fwd": 2 * fwd, "train_dgrad": 2 * (fwd - (2 * H + I * H)), "train_wgrad": 2 * fwd}


def synth_trajectories(n: int, seed: int, device) -> torch.Tensor:
    """(n, T, 3) fp32 absolute [t, x, y]: scenario uniform over the four, start uniform in the
    scenario's box, t_k = k * dt with dt ~ U(0.33, 2.20) s, speed 5-20 m/s decaying linearly,
    lateral offset a cumulative N(0, 0.15^2) walk."""
    g = torch.Generator(device=device).manual_seed(seed)
    box = torch.tensor(SCENARIOS, dtype=torch.float32, device=device)
    s = torch.randint(0, 4, (n,), generator=g, device=device)
    b = box[s]
    u = torch.rand(n, 6, generator=g, device=device)
    x0 = b[:, 0] + (b[:, 1] - b[:, 0]) * u[:, 0]
    y0 = b[:, 2] + (b[:, 3] - b[:, 2]) * u[:, 1]
    dt = 0.33 + (2.20 - 0.33) * u[:, 2]
    v0 = 5.0 + 15.0 * u[:, 3]
    v_end = u[:, 4] * v0
    k = torch.arange(T, dtype=torch.float32, device=device)
    t = k[None, :] * dt[:, None]
    frac = k[None, :] / (T - 1)
    speed = v0[:, None] + (v_end - v0)[:, None] * frac
    seg = torch.zeros(n, T, device=device)
    seg[:, 1:] = 0.5 * (speed[:, 1:] + speed[:, :-1]) * dt[:, None]
    along = torch.cumsum(seg, 1) * b[:, 5:6]
    lat = torch.cumsum(torch.randn(n, T, generator=g, device=device) * 0.15, 1)
    lat[:, 0] = 0.0
    is_y = b[:, 4:5] > 0.5
    x = x0[:, None] + torch.where(is_y, lat, along)
    y = y0[:, None] + torch.where(is_y, along, lat)
    return torch.stack([t, x, y], -1).contiguous()


# ---------------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc, self.thread = index, [], None, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return self
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def mark(self):
        return time.perf_counter()

    def stop(self, t0=None, t1=None):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        rows = [r for (ts, r) in self.rows if (t0 is None or ts >= t0) and (t1 is None or ts <= t1 + 0.15)] or \
               [r for (_, r) in self.rows]
        sm, mx, pw, reasons = [], [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for r in rows:
            if len(r) < 7:
                continue
            try:
                sm.append(float(r[0])); mx.append(float(r[1])); pw.append(float(r[2]))
            except ValueError:
                continue
            for nme, val in zip(names, r[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------- CPU arms
def cpu_train_rate(batch_cpu: torch.Tensor, budget_s: float, max_steps: int = 10 ** 9):
    """The oracle's full training step (port of Training_VAE.py:345-363: offset transform,
    forward, loss, autograd backward, torch-style Adam) on all host cores."""
    from oracle import vae_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    p = O.init_params(T, L, seed=0)
    adam = O.AdamState(p, lr=LR)
    B = batch_cpu.shape[0]

    def one():
        eps = torch.randn(B, L)
        _, grads, _ = O.loss_and_grads(p, batch_cpu, eps, WEIGHTS)
        adam.step(p, grads)

    one(); one()
    t0 = time.perf_counter()
    n = 0
    while n < max_steps and (n < 3 or time.perf_counter() - t0 < budget_s):
        one()
        n += 1
    dt = time.perf_counter() - t0
    return n * B / dt, n, dt


def cpu_decode_rate(rows: int, budget_s: float):
    from oracle import vae_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    p = O.init_params(T, L, seed=0)
    start = torch.tensor([SCENARIO_DEFAULT_START[0]], dtype=torch.float32)
    z = torch.randn(rows, L)
    O.generate(p, z, start)
    t0 = time.perf_counter()
    n = 0
    while n < 2 or time.perf_counter() - t0 < budget_s:
        z = torch.randn(rows, L)           # the reference draws the latents on the host too (Tools.py:46)
        O.generate(p, z, start)
        n += 1
    dt = time.perf_counter() - t0
    return n * rows / dt, n, dt


def tracker_jobs(n: int, seed: int, device):
    """Inputs of the MPC tracker from the synthetic trajectories: [x, y, t] float32 waypoints as
    Distribution.py:74-78 hands them over and initial states [x, y, theta, vx, vy] along the first segment."""
    traj = synth_trajectories(n, seed, device)                  # (n, T, 3) [t, x, y], t0 = 0
    way = traj[:, :, [1, 2, 0]].contiguous()
    d = (way[:, 1] - way[:, 0]).double()
    vx, vy = d[:, 0] / d[:, 2], d[:, 1] / d[:, 2]
    init = torch.stack([way[:, 0, 0].double(), way[:, 0, 1].double(), torch.atan2(vy, vx), vx, vy], 1).contiguous()
    return way, init


def cpu_tracker_rate(budget_s: float) -> dict:
    """Controller calls per second of the reference's tracker on ONE host core (SLSQP is sequential; the port
    oracle/mpc_oracle.py of MPC/MPC_Tracking.py, scenario time step 0.02 s, horizons 30 / 20 as Distribution.py:94-101)."""
    from oracle import mpc_oracle as MO
    way, init = tracker_jobs(4, 11, "cpu")
    w, s0 = way[0].numpy(), init[0].numpy()
    MO.track(w, s0, 0.02, max_steps=1)
    steps = 0
    t0 = time.perf_counter()
    while time.perf_counter() - t0 < budget_s:
        MO.track(w, s0, 0.02, max_steps=4)
        steps += 4
    dt = time.perf_counter() - t0
    return {"value": steps / dt, "unit": "controller calls/s", "cores": 1, "kind": "port",
            "sample": f"{steps} controller calls (SLSQP, 40 variables, finite-difference gradients) of one synthetic trajectory in {dt:.1f} s; "
                      "the reference's own class needs ~1 s per call on the same problem (NumPy objective)"}


def reference_config_legs(gpu: bool) -> dict:
    """BASELINE configs[0] (the reference's own configuration): training on the shipped
    StaticBlindTown05 set with batch 38 = the whole set (Training_VAE.py:275-280: one step per epoch), and the
    generate helper at batch 1, with and without the checkpoint load the reference repeats on every call
    (Tools.py:39-41).  gpu: this repo's drop-in modules; else the oracle port on the host."""
    import tempfile

    import numpy as np
    from oracle import vae_oracle as O
    gold = os.path.join(ROOT, "tests", "golden")
    data = np.load(os.path.join(gold, "data_sce1_cond.npy")).astype(np.float32)        # (38, 10, 3)
    ck = np.load(os.path.join(gold, "ckpt_sce1_cond.npz"))
    out = {"train": {"batch": int(data.shape[0]), "data": "tests/golden/data_sce1_cond.npy (the shipped trajectory_sce1_cond.npy)"},
           "decode_b1": {"checkpoint": "the shipped vae_offset_sce1_cond_ld8_epoch3000.pth (tests/golden/ckpt_sce1_cond.npz)"}}
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "vae_offset_sce1_cond_ld8_epoch3000.pth")
        torch.save({k: torch.from_numpy(ck[k]) for k in ck.files}, path)
        sx, sy = float(data[0, 0, 1]), float(data[0, 0, 2])
        if gpu:
            import Tools
            from dmvae import ConditionalTrajectoryVAE
            from dmvae.train import FusedTrainer
            torch.manual_seed(0)
            model = ConditionalTrajectoryVAE(T, D, L).to("cuda")
            tr = FusedTrainer(model, lr=LR, weights=WEIGHTS, seed=0)
            batch = torch.from_numpy(data).cuda()
            for _ in range(20):
                tr.step(batch)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            n = 400
            for _ in range(n):
                float(tr.step(batch)[0])              # the loop reads its loss every step (Training_VAE.py:366-370)
            dt = time.perf_counter() - t0
            out["train"].update(steps_per_s=n / dt, samples_per_s=n * data.shape[0] / dt,
                                how="FusedTrainer.step per epoch + one loss read per step, host-driven launches (FFMA kernels: batches "
                                    "of at most 128 rows), wall clock")
            for tag, clear, n in (("cached_checkpoint", False, 300), ("with_checkpoint_load", True, 30)):
                Tools.load_model_and_generate_trajectory(path, sx, sy, T, D, L)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                for _ in range(n):
                    if clear:
                        Tools._MODEL_CACHE.clear()
                    Tools.load_model_and_generate_trajectory(path, sx, sy, T, D, L)
                out["decode_b1"][tag + "_calls_per_s"] = n / (time.perf_counter() - t0)
            out["decode_b1"]["how"] = ("Tools.load_model_and_generate_trajectory, one trajectory per call, result on the host; the drop-in "
                                       "caches the checkpoint by (path, mtime, size) - with_checkpoint_load clears that cache before every call")
        else:
            torch.set_num_threads(os.cpu_count() or 1)
            p = O.init_params(T, L, seed=0)
            adam = O.AdamState(p, lr=LR)
            batch = torch.from_numpy(data)

            def step():
                _, grads, _ = O.loss_and_grads(p, batch, torch.randn(batch.shape[0], L), WEIGHTS)
                adam.step(p, grads)

            for _ in range(5):
                step()
            t0 = time.perf_counter()
            n = 0
            while time.perf_counter() - t0 < 3.0:
                step()
                n += 1
            dt = time.perf_counter() - t0
            out["train"].update(steps_per_s=n / dt, samples_per_s=n * data.shape[0] / dt,
                                how=f"oracle port of Training_VAE.py:345-363, {torch.get_num_threads()} threads, wall clock")
            params = {k: torch.from_numpy(ck[k]) for k in ck.files}
            start = torch.tensor([[sx, sy]], dtype=torch.float32)
            for tag, load, budget in (("cached_checkpoint", False, 2.0), ("with_checkpoint_load", True, 3.0)):
                t0 = time.perf_counter()
                n = 0
                while time.perf_counter() - t0 < budget:
                    if load:                                         # Tools.py:39-41: module construction + torch.load per call
                        O.init_params(T, L)
                        params = torch.load(path, map_location="cpu")
                    O.generate(params, torch.randn(1, L), start)
                    n += 1
                out["decode_b1"][tag + "_calls_per_s"] = n / (time.perf_counter() - t0)
            out["decode_b1"]["how"] = "oracle port of Tools.py:18-65 on the host; with_checkpoint_load = what the reference does on every call"
    return out


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path (the oracle port:
    the reference is a flat directory of Python scripts, not installable, and does not travel
    to the GPU box) on all host cores, same config / metric / unit as the CUDA arm."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    B = args.batch
    data = synth_trajectories(B * 4, 0, "cpu")
    cores = os.cpu_count() or 1
    # one "step" = one batch of 4096; warm-up + K steps, each bounded
    steps = max(1, min(args.steps, 200))
    t_budget = 120.0
    from oracle import vae_oracle as O
    torch.set_num_threads(cores)
    p = O.init_params(T, L, seed=0)
    adam = O.AdamState(p, lr=LR)

    def one(i):
        b = data[(i % 4) * B:(i % 4 + 1) * B]
        eps = torch.randn(B, L)
        _, grads, _ = O.loss_and_grads(p, b, eps, WEIGHTS)
        adam.step(p, grads)

    for i in range(min(args.warmup, 5)):
        one(i)
    t0 = time.perf_counter()
    done = 0
    for i in range(steps):
        one(i)
        done += 1
        if time.perf_counter() - t0 > t_budget:
            break
    dt = time.perf_counter() - t0
    value = done * B / dt
    dec_rate, dec_n, dec_dt = cpu_decode_rate(1 << 18, 5.0)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": done,
        "warmup": min(args.warmup, 5), "ms_per_step": 1e3 * dt / done, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {**workload_config(B), "device": "host CPU, torch %s, %d threads" % (torch.__version__, torch.get_num_threads()),
                   "what": "the oracle port of the reference's PyTorch-CPU path (the reference is a flat directory of scripts: "
                           "not installable, and /root/reference does not exist on the GPU box)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                         "sample": f"{done} steps x {B} rows of the same workload"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "decode": {"metric": "decoded_trajectories_per_sec", "value": dec_rate, "unit": "trajectories/s",
                   "sample": f"{dec_n} x {1 << 18} rows, shared start, host randn + cond-encoder + decoder + offset add"},
        "reference_config": reference_config_legs(gpu=False),
        "mpc_tracker": cpu_tracker_rate(8.0),
        "gpu_launches": 0,
    }
    emit(line)


# ---------------------------------------------------------------------------------- CUDA arm
def profile(lib, fn, iters):
    """Run fn() iters times with the library's event-pair profiling on; returns
    {kernel_name: (total_ms, launches)}."""
    from dmvae import _lib
    n = _lib.KERNEL_COUNT
    ms = (ctypes.c_double * n)()
    cnt = (ctypes.c_int64 * n)()
    _lib.check(lib.dmvae_profile_begin(), "dmvae_profile_begin")
    for i in range(iters):
        fn(i)
    torch.cuda.synchronize()
    _lib.check(lib.dmvae_profile_end(ms, cnt, n), "dmvae_profile_end")
    return {lib.dmvae_kernel_name(i).decode(): (ms[i], cnt[i]) for i in range(n) if cnt[i]}


def ffma_peak_tflops(lib) -> float:
    from dmvae import _lib
    sink = torch.zeros(4, device="cuda")
    flop = ctypes.c_double(0.0)
    best = 0.0
    iters = 20000   # ~1.4 ms at 75 TFLOP/s
    for rep in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(lib.dmvae_ffma_probe(iters, _lib.ptr(sink), ctypes.byref(flop), _lib.stream_ptr()), "dmvae_ffma_probe")
        e1.record()
        torch.cuda.synchronize()
        if rep:
            best = max(best, flop.value / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    return best


def tf32_peak_tflops(lib) -> dict:
    """Dense tcgen05.mma kind::tf32 rate measured with dmvae_tf32_probe: operands in shared memory (N = 256) and A in
    tensor memory with N = 128 (the shape the training chain issues)."""
    from dmvae import _lib
    sink = torch.zeros(4, device="cuda")
    out = {}
    for mode, key in ((0, "ss_n256"), (1, "ts_n128")):
        flop = ctypes.c_double(0.0)
        best = 0.0
        iters = 4000 if mode == 0 else 8000        # ~1 ms
        for rep in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _lib.check(lib.dmvae_tf32_probe(iters, mode, _lib.ptr(sink), ctypes.byref(flop), _lib.stream_ptr()), "dmvae_tf32_probe")
            e1.record()
            torch.cuda.synchronize()
            if rep:
                best = max(best, flop.value / (e0.elapsed_time(e1) * 1e-3) / 1e12)
        out[key] = best
    return out


def bind_to_gpu_numa_node(index: int) -> str:
    """Pins this process to the CPUs next to its GPU (NVML's affinity mask) BEFORE any pinned host buffer is
    allocated, so that first-touch places those buffers on the GPU's own NUMA node."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (w >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
            return f"{len(cpus)} CPUs next to GPU {index}"
        return "NVML affinity mask empty"
    except Exception as e:  # noqa: BLE001 - placement is an optimisation only
        return f"unchanged ({type(e).__name__})"


def run_cuda(args):
    import torch.distributed as dist
    from dmvae import ConditionalTrajectoryVAE, _lib
    from dmvae.train import FusedTrainer

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the CUDA arm has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa_note = bind_to_gpu_numa_node(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"    # the version banner would land on stdout next to the one JSON line
        dist.init_process_group("nccl", device_id=dev)
    if world != args.gpus and rank == 0:
        print(f"bench.py: --gpus {args.gpus} but WORLD_SIZE={world}; using {world}", file=sys.stderr)
    lib = _lib.lib()
    fl = flops_per_unit()
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except OSError:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    tensor_peak = float(peaks.get("bf16_tflops_sustained", 1400.0))
    peak_src = "MEASURED_PEAKS.json" if peaks else "fallback"

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---------------------------------------------------------------- model + data
    B = args.batch
    torch.manual_seed(0)
    model = ConditionalTrajectoryVAE(T, D, L).to(dev)
    trainer = FusedTrainer(model, lr=LR, weights=WEIGHTS, seed=0)
    rows = args.dataset_rows                       # per GPU, resident in HBM; 2^21 x 120 B = 252 MB > L2 (126 MB)
    data = synth_trajectories(rows, 1000 + rank, dev)
    n_batches = rows // B
    Bg = B * world

    from dmvae.parallel import DataParallelTrainer
    # N > 1: the gradient exchange runs inside the update kernel over peer memory (NVLink) when the ranks can map
    # each other's buffers, else fwd+bwd, ONE NCCL all-reduce of [grads | 5 losses], replicated Adam
    dp = DataParallelTrainer(trainer, exchange=args.exchange)
    if rank == 0 and dp.exchange_note:
        print("bench.py: " + dp.exchange_note, file=sys.stderr)

    # ---------------------------------------------------------------- N > 1: is the exchanged sum RIGHT?
    # (replicas_identical below only proves that the ranks agree.)  Before anything is timed, every rank takes 3
    # data-parallel steps on its slice of one global batch that all ranks generate identically (injected noise)
    # and, on a clone of the model, the same 3 steps single-rank on the whole global batch; parameters and loss
    # terms must agree within 2e-5 or the run is abandoned.
    dp_parity = None
    if world > 1:
        gdata = synth_trajectories(Bg, 4242, dev)                       # same seed on every rank
        geps = torch.randn(3, Bg, L, generator=torch.Generator(device=dev).manual_seed(4243), device=dev)
        m_dp, m_one = ConditionalTrajectoryVAE(T, D, L).to(dev), ConditionalTrajectoryVAE(T, D, L).to(dev)
        for m_ in (m_dp, m_one):
            m_.load_state_dict(model.state_dict())
        dp_chk = DataParallelTrainer(FusedTrainer(m_dp, lr=LR, weights=WEIGHTS, seed=0), exchange=dp.exchange)
        one_chk = FusedTrainer(m_one, lr=LR, weights=WEIGHTS, seed=0)
        lo_r = rank * B
        loss_err = 0.0
        for s_ in range(3):
            l_dp = dp_chk.step(gdata[lo_r:lo_r + B], eps=geps[s_, lo_r:lo_r + B]).double().cpu()
            l_one = one_chk.step(gdata, eps=geps[s_]).double().cpu()
            loss_err = max(loss_err, float(((l_dp - l_one).abs() / l_one.abs().clamp_min(1e-30)).max()))
        dp_chk.check_exchange()
        p_dp, p_one = m_dp.flat_parameters().double(), m_one.flat_parameters().double()
        g_dp, g_one = dp_chk.engine.grads.double(), one_chk.grads.double()
        dp_parity = {"world": world, "exchange": dp_chk.exchange, "steps": 3, "global_batch": Bg,
                     "param_rel_err": max_over_ranks(float((p_dp - p_one).abs().max() / p_one.abs().max())),
                     "grad_rel_err": max_over_ranks(float((g_dp - g_one).abs().max() / g_one.abs().max())),
                     "loss_rel_err": max_over_ranks(loss_err),
                     "reference": "FusedTrainer.step on the whole global batch, one rank, same injected eps", "tol": 2e-5}
        if max(dp_parity["param_rel_err"], dp_parity["loss_rel_err"], dp_parity["grad_rel_err"]) > 2e-5 or \
                not all(v == v for v in (dp_parity["param_rel_err"], dp_parity["loss_rel_err"])):
            raise SystemExit(f"bench.py: data-parallel step disagrees with the single-rank step: {dp_parity}")
        del gdata, geps, m_dp, m_one, dp_chk, one_chk

    # single GPU: the whole step is one CUDA graph (dmvae_train_step_dev: Adam step index in device memory);
    # the batch of the step is copied device-to-device into the graph's input buffer
    # The set stays in HBM and the kernel picks the batch of update t from the device-side step counter
    # (dmvae_train_step_resident): a step is one graph replay, nothing else.
    resident = not args.no_graph and (world == 1 or dp.exchange == "peer")
    if args.no_graph:
        gstep = None
    elif world == 1:
        gstep = trainer.capture(B, dataset=data)
    elif resident:
        gstep = dp.capture(B, dataset=data)     # + the exchange of [grads | losses] inside the update kernel
    else:
        gstep = dp.capture(B)                   # + the NCCL all-reduce captured in the graph

    def train_step(i):
        b = data[(i % n_batches) * B:(i % n_batches + 1) * B]
        if gstep is not None and resident:
            gstep.replay()                               # chain + wgrad (one launch at this batch), reduce + Adam + packed refresh
        elif gstep is not None:
            gstep.batch.copy_(b, non_blocking=True)
            gstep.replay()
        elif world == 1:
            trainer.step(b, sample_offset=0)
        else:
            dp.step(b)

    K, W = args.steps, max(args.warmup, 3)
    clocks = ClockSampler(local).start() if rank == 0 else None
    for i in range(W):
        train_step(i)
    barrier()
    launches0 = lib.dmvae_launch_count(-1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_mark0 = time.perf_counter()
    e0.record()
    for i in range(K):
        train_step(W + i)
    e1.record()
    barrier()
    t_mark1 = time.perf_counter()
    launches = lib.dmvae_launch_count(-1) - launches0
    if gstep is not None:
        launches += K * gstep.kernels        # kernels inside the replayed graphs (counted once, at capture)
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    value = K * Bg / (ms_total * 1e-3)
    losses_end = [float(v) for v in trainer.losses.cpu()]

    # ---------------------------------------------------------------- per-kernel live timing (separate pass)
    def train_step_host(i):      # host-driven launches: the library brackets each kernel with an event pair
        b = data[(i % n_batches) * B:(i % n_batches + 1) * B]
        if world == 1:
            trainer.step(b, sample_offset=0)
        else:
            dp.step(b)

    prof = profile(lib, lambda i: train_step_host(W + K + i), min(K, 200))
    n_steps_prof = max(max(v[1] for v in prof.values()), 1) if prof else 1
    step_kernel_ms = sum(v[0] for v in prof.values()) / n_steps_prof
    shares = {k: round(v[0] / max(sum(x[0] for x in prof.values()), 1e-12), 4) for k, v in prof.items()}
    kernel_flop = {"chain_kernel": B * (fl["train_fwd"] + fl["train_dgrad"]), "wgrad_kernel": B * fl["train_wgrad"],
                   "train_kernel(fused)": B * fl["train"], "train_tc_fused_kernel": B * fl["train"]}
    dominant = max((k for k in prof if k in kernel_flop), key=lambda k: prof[k][0])
    dom_ms = prof[dominant][0] / max(prof[dominant][1], 1)
    per_kernel_us = {k: round(v[0] / max(v[1], 1) * 1e3, 2) for k, v in prof.items()}

    # the FFMA kernels on the same workload, for the comparison north_star asks for (N > 1: with the NCCL
    # all-reduce - the in-kernel exchange belongs to the tensor-core update kernel)
    _lib.check(lib.dmvae_set_train_impl(1), "dmvae_set_train_impl")
    dp_ffma = dp if dp.exchange != "peer" else DataParallelTrainer(trainer, exchange="nccl")

    def ffma_step(i):
        b = data[(i % n_batches) * B:(i % n_batches + 1) * B]
        if world == 1:
            trainer.step(b, sample_offset=0)
        else:
            dp_ffma.step(b)

    for i in range(5):
        ffma_step(i)
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    Kf = max(10, min(K, 200))
    f0.record()
    for i in range(Kf):
        ffma_step(i)
    f1.record()
    barrier()
    ffma_value = Kf * Bg / (max_over_ranks(f0.elapsed_time(f1)) * 1e-3)
    _lib.check(lib.dmvae_set_train_impl(0), "dmvae_set_train_impl")

    # the same step at a throughput batch (configs[3]: large synthetic set, data parallel): 65536 rows per GPU
    big = None
    Bbig = min(args.big_batch, rows)
    if Bbig > B:
        nb = rows // Bbig

        def big_step(i):
            b = data[(i % nb) * Bbig:(i % nb + 1) * Bbig]
            if world == 1:
                trainer.step(b, sample_offset=0)
            else:
                dp.step(b)

        for i in range(3):
            big_step(i)
        barrier()
        Kb = max(5, min(K, 30))
        f0.record()
        for i in range(Kb):
            big_step(i)
        f1.record()
        barrier()
        big_ms = max_over_ranks(f0.elapsed_time(f1)) / Kb
        bprof = profile(lib, big_step, 5)
        big = {"batch_per_gpu": Bbig, "value": Bbig * world / (big_ms * 1e-3), "unit": UNIT, "ms_per_step": big_ms,
               "kernel_us": {k: round(v[0] / max(v[1], 1) * 1e3, 1) for k, v in bprof.items()},
               "step_tflops": Bbig * world * fl["train"] / (big_ms * 1e-3) / 1e12}

    # ---------------------------------------------------------------- e2e (host buffers, per-step loss read)
    # The public API used the way a training loop would use it: two captured steps (GraphStep) with their own device
    # input buffer and pinned loss slot; the batch of step i + 1 goes host -> device on a copy stream while step i
    # computes; the five loss terms of every step come back to pinned host memory inside the step's graph.
    NSLOT = 8
    host_batches = [torch.empty(B, T, 3, dtype=torch.float32).pin_memory() for _ in range(NSLOT)]
    for j, hb in enumerate(host_batches):
        hb.copy_(data[j * B:(j + 1) * B].cpu())
    host_losses = [torch.empty(5, dtype=torch.float32).pin_memory() for _ in range(2)]
    copy_stream = torch.cuda.Stream()
    main_stream = torch.cuda.current_stream()
    if gstep is not None:
        e2e_graphs = [(trainer.capture(B, host_losses=hl) if world == 1 else dp.capture(B, host_losses=hl))
                      for hl in host_losses]
        dbufs = [g.batch for g in e2e_graphs]
    else:
        e2e_graphs = None
        dbufs = [torch.empty(B, T, 3, dtype=torch.float32, device=dev) for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]    # batch of the slot is on the device
    done = [torch.cuda.Event() for _ in range(2)]     # the step of the slot has finished (its device buffer is free again)
    for ev in done:
        ev.record(main_stream)

    def e2e_upload(i):        # H2D of the batch of step i from pinned host memory, on the copy stream
        j = i % 2
        copy_stream.wait_event(done[j])
        with torch.cuda.stream(copy_stream):
            dbufs[j].copy_(host_batches[i % NSLOT], non_blocking=True)
            ready[j].record(copy_stream)

    def e2e_launch(i):
        j = i % 2
        main_stream.wait_event(ready[j])
        if e2e_graphs is not None:
            e2e_graphs[j].replay()
        else:
            losses = trainer.step(dbufs[j]) if world == 1 else dp.step(dbufs[j])
            host_losses[j].copy_(losses, non_blocking=True)
        done[j].record(main_stream)

    def e2e_read(i):          # the step's result on the host: waits for exactly that step
        done[i % 2].synchronize()
        return float(host_losses[i % 2][0])

    def run_e2e(mode):
        """blocking: the host reads the loss of step i before it launches step i + 1 (the reference loop's .item() per
        step, Training_VAE.py:366-370); the upload of batch i + 1 is already in flight;
        overlapped: the read of step i - 1 follows the launch of step i, so the device never waits for the host;
        serial: no copy/compute overlap at all - upload, step, read, one after the other."""
        for i in range(W):
            e2e_upload(i)
            e2e_launch(i)
            e2e_read(i)
        barrier()
        t0 = time.perf_counter()
        if mode == "serial":
            for i in range(K):
                e2e_upload(i)
                e2e_launch(i)
                e2e_read(i)
        else:
            e2e_upload(0)
            for i in range(K):
                e2e_launch(i)                      # its batch is on the device already
                if i + 1 < K:
                    e2e_upload(i + 1)              # issued while step i computes
                if mode == "blocking":
                    e2e_read(i)
                elif i > 0:
                    e2e_read(i - 1)
            if mode != "blocking":
                e2e_read(K - 1)
        barrier()
        return K * Bg / max_over_ranks(time.perf_counter() - t0)

    e2e = {m: run_e2e(m) for m in ("blocking", "overlapped", "serial")}
    e2e_obj = {"value": e2e["blocking"], "unit": UNIT, "h2d_bytes_per_step": B * T * 3 * 4 * world,
               "d2h_bytes_per_step": 20 * world,
               "overlapped_value": e2e["overlapped"], "serial_value": e2e["serial"],
               "note": "every step: H2D of its batch from pinned host memory, the fused step, D2H of its 5 loss terms, "
                       "all inside the timed region (host wall clock, max over ranks).  value: the host waits for and "
                       "reads the loss of step i before it launches step i + 1, while the batch of step i + 1 is already "
                       "being uploaded on a copy stream; overlapped_value: it reads the loss of step i - 1 right after "
                       "launching step i; serial_value: upload, step and read strictly one after the other"}

    # ---------------------------------------------------------------- decode (second half of the metric)
    R = args.decode_rows
    outs = [torch.empty(R, T, 3, dtype=torch.float32, device=dev) for _ in range(4)]   # 4 x 126 MB > L2
    starts = [torch.tensor([s], dtype=torch.float32, device=dev) for s in SCENARIO_DEFAULT_START]

    def decode_step(i):
        for s in range(4):
            model.generate(starts[s], n=R, seed=s, sample_offset=rank * R, out=outs[s])

    Kd = max(3, min(K, 20))
    for i in range(3):
        decode_step(i)
    barrier()
    d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    d0.record()
    for i in range(Kd):
        decode_step(i)
    d1.record()
    barrier()
    dec_ms = max_over_ranks(d0.elapsed_time(d1))
    dec_value = Kd * 4 * R * world / (dec_ms * 1e-3)
    dprof = profile(lib, decode_step, Kd)
    dk_ms, dk_n = dprof.get("decode_tc_kernel", dprof.get("decode_kernel", (0.0, 0)))
    dec_kernel_ms = dk_ms / max(dk_n, 1)
    # per-row start points (every row runs the condition encoder): the other decode variant
    pr_start = data[:R, 0, 1:3].contiguous() if rows >= R else synth_trajectories(R, 7, dev)[:, 0, 1:3].contiguous()
    for i in range(2):
        model.generate(pr_start, n=R, seed=9, out=outs[0])
    barrier()
    d0.record()
    for i in range(Kd):
        model.generate(pr_start, n=R, seed=9, out=outs[i % 4])
    d1.record()
    barrier()
    pr_ms = max_over_ranks(d0.elapsed_time(d1)) / Kd
    # e2e decode: public API, results land in pinned host memory.  Two device buffers and two pinned host buffers:
    # launch s + 1 decodes into the other buffer while launch s travels device -> host on a copy stream.
    host_outs = [torch.empty(R, T, 3, dtype=torch.float32).pin_memory() for _ in range(2)]
    d2h_stream = torch.cuda.Stream()
    dec_done = [torch.cuda.Event() for _ in range(2)]     # decode into device buffer j finished
    d2h_done = [torch.cuda.Event() for _ in range(2)]     # device buffer j (and host buffer j) free again
    for ev in d2h_done:
        ev.record(torch.cuda.current_stream())
    n_e2e = max(2, Kd // 4) * 4

    def decode_e2e_pass(count):
        cur = torch.cuda.current_stream()
        for i in range(count):
            j, sc = i % 2, i % 4
            cur.wait_event(d2h_done[j])                                  # its previous copy has left the buffer
            model.generate(starts[sc], n=R, seed=sc, sample_offset=rank * R, out=outs[j])
            dec_done[j].record(cur)
            d2h_stream.wait_event(dec_done[j])
            with torch.cuda.stream(d2h_stream):
                host_outs[j].copy_(outs[j], non_blocking=True)
                d2h_done[j].record(d2h_stream)
        d2h_stream.synchronize()

    decode_e2e_pass(2)
    barrier()
    t0 = time.perf_counter()
    decode_e2e_pass(n_e2e)
    barrier()
    dec_e2e_dt = max_over_ranks(time.perf_counter() - t0)
    dec_e2e = n_e2e * R * world / dec_e2e_dt

    # validation metrics over one launch's worth of decoded trajectories: three HBM-bound scans,
    # timed with CUDA events around the C ABI calls (dmvae/validation.py adds a small device -> host read per call)
    from dmvae import validation as V
    metrics = None
    if rank == 0:
        import numpy as np
        tr_m = outs[0]                                          # (R, T, 3) [t, x, y], left there by the decode legs
        v_m, (v_lo, v_hi) = V.waypoint_speeds(tr_m)
        edges_m = np.ascontiguousarray(np.linspace(v_lo, v_hi, 50))
        h_m = V.histogram(v_m, edges_m)
        H_m = V.trajectories_per_cell(tr_m, "vae_offset_sce4_cond")
        gx0, gnx, gy0, gny = V.grid_edges("vae_offset_sce4_cond")
        mm_d = torch.empty(2, device=dev)
        cnt_d = torch.empty(49, dtype=torch.int64, device=dev)
        cells_d = torch.empty((gnx - 1) * (gny - 1), dtype=torch.int64, device=dev)
        calls = (
            ("waypoint_speeds", R * T * 3 * 4 + R * T * 4,
             lambda: lib.dmvae_waypoint_speeds(_lib.ptr(tr_m), R, T, 0, _lib.ptr(v_m), _lib.ptr(mm_d), _lib.stream_ptr())),
            ("histogram_49_bins", R * T * 4,
             lambda: lib.dmvae_histogram(_lib.ptr(v_m), R * T, edges_m.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), 49,
                                         _lib.ptr(cnt_d), _lib.stream_ptr())),
            ("trajectories_per_cell", R * T * 3 * 4,
             lambda: lib.dmvae_trajectories_per_cell(_lib.ptr(tr_m), R, T, 0, gx0, 1.0, gnx, gy0, 1.0, gny, _lib.ptr(cells_d),
                                                     _lib.stream_ptr())))
        metrics = {"trajectories": R, "note": "CUDA events around 20 back-to-back calls of the C entry point; algorithmic bytes: 120 B "
                                              "read per trajectory and pass (+ 40 B of speeds written / read); peak = measured HBM copy rate"}
        for name, nbytes, call in calls:
            _lib.check(call(), name)
            torch.cuda.synchronize()
            m0, m1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            m0.record()
            for _ in range(20):
                call()
            m1.record()
            torch.cuda.synchronize()
            ms_call = m0.elapsed_time(m1) / 20
            metrics[name] = {"ms": ms_call, "gbs": nbytes / ms_call / 1e6, "frac_of_hbm_peak": nbytes / ms_call / 1e6 / hbm_peak}
        metrics["check"] = {"histogram_total": int(h_m.sum()), "expected": R * T, "cells_max": int(H_m.max()),
                            "cells_equal_wrapper": bool((cells_d.cpu().numpy().reshape(H_m.shape) == H_m).all())}

    # batched MPC tracker: controller calls per second over synthetic waypoint sets, one thread per
    # trajectory; e2e = host waypoints in, full state histories out (the tracked_trajectory_* content)
    mpc = None
    if rank == 0:
        from dmvae.tracker import BatchTracker
        n_mpc, k_mpc = args.mpc_rows, 40
        way_m, init_m = tracker_jobs(n_mpc, 11, dev)
        bt = BatchTracker(way_m, init_m, 0.02, 30, 20)
        bt.advance(10)                                          # the first calls start without a previous solution
        torch.cuda.synchronize()
        it0 = bt.iters.clone()
        m0, m1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        m0.record()
        bt.advance(k_mpc)
        m1.record()
        torch.cuda.synchronize()
        mpc_ms = m0.elapsed_time(m1)
        running = int((bt.n_steps >= bt.step).sum())
        # end to end through dmvae.tracker: pinned host inputs -> device, set-up, the first k_mpc steps with their state
        # and control histories, histories back to pinned host memory
        way_h, init_h = way_m.cpu().pin_memory(), init_m.cpu().pin_memory()
        st_h = torch.empty(n_mpc, k_mpc + 1, 4, dtype=torch.float64).pin_memory()
        ct_h = torch.empty(n_mpc, k_mpc, 2, dtype=torch.float64).pin_memory()
        st_d = torch.empty(n_mpc, k_mpc + 1, 4, dtype=torch.float64, device=dev)
        ct_d = torch.empty(n_mpc, k_mpc, 2, dtype=torch.float64, device=dev)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        bt2 = BatchTracker(way_h.to(dev, non_blocking=True), init_h.to(dev, non_blocking=True), 0.02, 30, 20)
        bt2.advance(k_mpc, st_d, ct_d)
        st_h.copy_(st_d, non_blocking=True)
        ct_h.copy_(ct_d, non_blocking=True)
        torch.cuda.synchronize()
        mpc_e2e_dt = time.perf_counter() - t0
        mpc = {"metric": "mpc_controller_calls_per_sec", "value": n_mpc * k_mpc / (mpc_ms * 1e-3), "unit": "controller calls/s",
               "config": {"workload": f"{n_mpc} synthetic waypoint sets (10 waypoints, float32), time step 0.02 s, prediction / control "
                                      "horizon 30 / 20 (Distribution.py:94-101), steps 10..50 of every trajectory",
                          "steps_per_trajectory_mean": float(bt.n_steps.mean()), "trajectories_still_running": running},
               "ms": mpc_ms, "solver_iterations_per_call": float((bt.iters - it0).double().mean()) / k_mpc,
               "complete_trajectories_per_sec_implied": n_mpc * k_mpc / (mpc_ms * 1e-3) / float(bt.n_steps.mean()),
               "e2e": {"value": n_mpc * k_mpc / mpc_e2e_dt, "unit": "controller calls/s",
                       "h2d_bytes_per_step": int(way_h.numel() * 4 + init_h.numel() * 8),
                       "d2h_bytes_per_step": int(st_h.numel() * 8 + ct_h.numel() * 8),
                       "note": "set-up (interpolants, heading scan) + the first 40 steps from a cold solver + histories to the host"},
               "roofline": {"bound": "fp64 issue / local-memory latency (one thread per trajectory; see profiles/*_prof_mpc_track_kernel.txt)",
                            "traffic": (traffic_from_profile("mpc_track_kernel") or {}).get("bytes_per_launch"),
                            "traffic_source": (traffic_from_profile("mpc_track_kernel") or {}).get("source")},
               "cpu_baseline": cpu_tracker_rate(8.0) if (world == 1 and not args.no_cpu) else None}
        del bt, bt2, st_d, ct_d
        # complete runs through the public call (set-up, every step of every trajectory, histories in device memory):
        # the trajectories have 150 .. 990 steps, track_batch orders them so that the lanes of a warp finish together
        from dmvae.tracker import track_batch
        n_full = min(n_mpc, 16384)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        full_res = track_batch(way_m[:n_full], init_m[:n_full], 0.02)
        torch.cuda.synchronize()
        full_dt = time.perf_counter() - t0
        mpc["complete_runs"] = {"trajectories": n_full, "controller_calls": int(full_res.n_steps.sum()), "seconds": full_dt,
                                "trajectories_per_sec": n_full / full_dt, "controller_calls_per_sec": float(full_res.n_steps.sum()) / full_dt,
                                "solver_iterations_per_call": float(full_res.iterations.sum()) / float(full_res.n_steps.sum())}
        del full_res

    # the reference shuffles its data set every epoch (Training_VAE.py:327): the same resident step with the rows of
    # every epoch picked through the keyed permutation (dmvae_train_step_resident, shuffle = 1)
    shuffled = None
    if resident and gstep is not None:
        gsh = trainer.capture(B, dataset=data, shuffle=True, shuffle_seed=7) if world == 1 else \
            dp.capture(B, dataset=data, shuffle=True, shuffle_seed=7)
        for i in range(W):
            gsh.replay()
        barrier()
        f0.record()
        for i in range(K):
            gsh.replay()
        f1.record()
        barrier()
        sh_ms = max_over_ranks(f0.elapsed_time(f1))
        shuffled = {"value": K * Bg / (sh_ms * 1e-3), "unit": UNIT, "ms_per_step": sh_ms / K,
                    "note": "rows of every epoch picked in the kernel through a keyed permutation of the resident set "
                            "(scattered 120-byte reads instead of one bulk copy per tile)"}

    peak_ffma = ffma_peak_tflops(lib)
    tf32 = tf32_peak_tflops(lib)
    clk = clocks.stop(t_mark0, time.perf_counter()) if clocks is not None else None
    # every rank applied the same updates to the same bits (collective: all ranks take part)
    replicas_identical = dp.parameter_checksum(model.flat_parameters()) if world > 1 else None

    def leave():
        """Exit of a multi-rank job: the step graphs hold captured NCCL work and tearing the communicator
        down under them can block, so the ranks meet once more and leave without destroying the group."""
        if world > 1:
            barrier()
            sys.stdout.flush()
            sys.stderr.flush()
            os._exit(0)

    if rank != 0:
        leave()
        return

    # ---------------------------------------------------------------- CPU baseline beside it (rank 0, N = 1 only)
    cpu_obj, dec_cpu_obj = None, None
    if world == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        rate, n, dt = cpu_train_rate(data[:B].cpu(), 12.0)
        cpu_obj = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": f"{n} full train steps x {B} rows of the same workload in {dt:.1f} s (oracle = PyTorch-CPU port of Training_VAE.py:345-363)"}
        drate, dn, ddt = cpu_decode_rate(1 << 18, 6.0)
        dec_cpu_obj = {"value": drate, "unit": "trajectories/s", "cores": cores, "kind": "port",
                       "sample": f"{dn} x {1 << 18} rows in {ddt:.1f} s (oracle generate: host randn + cond-encoder + decoder + offset add)"}

    ref_cfg = None
    if world == 1:
        ref_cfg = {"b200": reference_config_legs(gpu=True)}
        if not args.no_cpu:
            ref_cfg["host_cpu_port"] = reference_config_legs(gpu=False)

    ach = kernel_flop[dominant] / (dom_ms * 1e-3) / 1e12 if dom_ms > 0 else 0.0
    tensor_burst = float(peaks.get("bf16_tflops", 1682.8 if not peaks else tensor_peak))
    on_tensor = dominant in ("chain_kernel", "wgrad_kernel", "train_tc_fused_kernel")
    peak = tensor_burst if on_tensor else peak_ffma
    dach = R * fl["decode_shared_start"] / (dec_kernel_ms * 1e-3) / 1e12 if dec_kernel_ms > 0 else 0.0
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32 (3xTF32 on tcgen05: tf32 hi/lo split operands, fp32 accumulate)", "data": "synthetic",
        "config": {**workload_config(B), "global_batch": Bg, "parallelism": f"dp{world}",
                   "eps": "in-kernel Philox4x32-10", "dataset_rows_per_gpu": rows, "dataset_rows_total": rows * world,
                   "host_placement": numa_note,
                   "launch": ("one CUDA graph per step, batch picked in the kernel from the resident set (device-side step counter" if resident else
                              "one CUDA graph per step (device-side Adam step counter") + (
                              ("; the NCCL all-reduce is captured in it)" if dp.exchange == "nccl" else ")")) if gstep is not None else "host-driven launches",
                   "l2": f"each step reads a different batch of a {rows * T * 3 * 4 / 1e6:.0f} MB resident set (> 126 MB L2); "
                         "weights and the per-step stash / slabs are L2-resident by design",
                   "collective": {"none": "none", "nccl": "NCCL all-reduce SUM of 128947 fp32 per step",
                                  "peer": "no library collective: the update kernel exchanges the 128947 fp32 of every rank "
                                          "over peer memory (NVLink), block by block, and sums them in rank order"}[dp.exchange]},
        "roofline": {"bound": "tensor" if on_tensor else "fp32", "kernel": dominant, "achieved": ach, "peak": peak,
                     "unit": "TFLOP/s", "frac": ach / peak if peak else None,
                     "traffic": (traffic_from_profile(dominant) or {}).get("bytes_per_launch"),
                     "traffic_source": (traffic_from_profile(dominant) or {}).get("source"),
                     "peak_source": ("dmvae_ffma_probe measured in this run (FP32 FFMA, all SMs)" if not on_tensor else
                                     "MEASURED_PEAKS.json bf16_tflops (dense bf16, burst: the kernel is timed alone)" if peaks
                                     else "fallback 1682.8 TFLOP/s: MEASURED_PEAKS.json absent"),
                     "tf32_dense_measured": tf32,
                     "tf32x3_ceiling": max(tf32.values()) / 3.0,
                     "frac_of_tf32x3_ceiling": ach / (max(tf32.values()) / 3.0) if on_tensor and max(tf32.values()) > 0 else None,
                     "tf32x3_ceiling_assumed_from_bf16": tensor_burst / 6.0,
                     "note": "3xTF32: every fp32-class product takes three TF32 passes, so fp32-equivalent work tops out at one "
                             "third of the dense TF32 rate, which dmvae_tf32_probe MEASURES in this run (tf32_dense_measured: "
                             "operands in shared memory with N = 256, and A in tensor memory with N = 128 - the shape the chain "
                             "issues); frac stays against the measured bf16 peak as the contract asks.  The chain walks 128-row "
                             "tiles and at batch 4096 there are only 32 of them for 148 SMs, so train_tc_fused_kernel runs the 32 "
                             "chain CTAs and 96 weight-gradient CTAs side by side in one launch (per-tile ready counters): the "
                             "launch is bound by the latency of one tile's 24-product chain, not by issue rate",
                     "flop_per_launch": kernel_flop[dominant], "kernel_ms": dom_ms, "kernel_us": per_kernel_us,
                     "kernel_share_of_step": shares, "step_kernel_ms_sum": step_kernel_ms,
                     "step_tflops": B * fl["train"] / (step_kernel_ms * 1e-3) / 1e12 if step_kernel_ms else None,
                     "hbm_gbs_achieved": B * T * 3 * 4 / (dom_ms * 1e-3) / 1e9 if dom_ms else None,
                     "hbm_gbs_peak": hbm_peak, "ffma_peak_tflops": peak_ffma,
                     "ffma_kernels_value": ffma_value, "tensor_peak_sustained": tensor_peak,
                     "tensor_peak_source": peak_src},
        "large_batch": big,
        "shuffled_resident_set": shuffled,
        "validation_metrics": metrics,
        "mpc_tracker": mpc,
        "reference_config": ref_cfg,
        "cpu_baseline": cpu_obj,
        "e2e": e2e_obj,
        "gpu_launches": int(launches),
        "clocks": clk,
        "losses_after_run": losses_end,
        "replicas_identical": replicas_identical,
        "dp_parity": dp_parity,
        "decode": {
            "metric": "decoded_trajectories_per_sec", "value": dec_value, "unit": "trajectories/s",
            "config": {"workload": f"configs[2]: 4 scenarios x {R} latents per GPU, in-kernel Philox, shared scenario start, "
                                   "cond-encoder hoisted, (rows,10,3) fp32 written to HBM", "rows_per_launch": R,
                       "l2": "4 output buffers of 126 MB cycle (> L2)"},
            "steps": Kd, "ms_per_step": dec_ms / Kd,
            "per_row_start_value": R * world / (pr_ms * 1e-3),
            "roofline": {"bound": "tensor", "kernel": "decode_tc_kernel", "achieved": dach, "peak": tensor_burst, "unit": "TFLOP/s",
                         "frac": dach / tensor_burst if tensor_burst else None,
                         "traffic": (traffic_from_profile("decode_tc_kernel") or {}).get("bytes_per_launch"),
                         "traffic_source": (traffic_from_profile("decode_tc_kernel") or {}).get("source"),
                         "traffic_note": "captured by scripts/gpu_check.sh at 262144 rows per launch (31 MB of output, which "
                                         "stays in the 126 MB L2 for the length of the kernel); the algorithmic bytes are 120 B per row",
                         "tf32_dense_measured": tf32, "tf32x3_ceiling": max(tf32.values()) / 3.0,
                         "frac_of_tf32x3_ceiling": dach / (max(tf32.values()) / 3.0) if max(tf32.values()) > 0 else None,
                         "tf32x3_ceiling_assumed_from_bf16": tensor_burst / 6.0,
                         "ffma_peak_tflops": peak_ffma,
                         "flop_per_launch": R * fl["decode_shared_start"], "kernel_ms": dec_kernel_ms,
                         "hbm_gbs_achieved": R * T * 3 * 4 / (dec_kernel_ms * 1e-3) / 1e9 if dec_kernel_ms else None,
                         "hbm_gbs_peak": hbm_peak,
                         "per_row_start_tflops": R * fl["decode_per_row_start"] / (pr_ms * 1e-3) / 1e12},
            "e2e": {"value": dec_e2e, "unit": "trajectories/s", "h2d_bytes_per_step": 8 * 4 * world,
                    "d2h_bytes_per_step": 4 * R * T * 3 * 4 * world},
            "cpu_baseline": dec_cpu_obj,
        },
    }
    emit(line)
    leave()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--dataset-rows", type=int, default=1 << 23,
                    help="rows per GPU of the resident set: 2^23 = 1.0 GB per GPU, 64 M rows on eight GPUs (BASELINE configs[3])")
    ap.add_argument("--big-batch", type=int, default=1 << 16, help="rows per GPU of the large-batch throughput leg")
    ap.add_argument("--decode-rows", type=int, default=1 << 20)
    ap.add_argument("--mpc-rows", type=int, default=1 << 16, help="trajectories of the MPC tracker leg")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg")
    ap.add_argument("--no-graph", action="store_true", help="host-driven steps instead of the CUDA-graph step")
    ap.add_argument("--exchange", choices=("auto", "peer", "nccl"), default="auto",
                    help="N > 1: gradient exchange inside the update kernel over peer memory, or one NCCL all-reduce")
    args = ap.parse_args()
    claim_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_cuda(args)


if __name__ == "__main__":
    main()
