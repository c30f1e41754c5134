import sys, ctypes, torch
sys.path.insert(0, "defensive-model-vae_b200")
from dmvae import _lib
lib = _lib.lib()
sink = torch.zeros(4, device="cuda")
for mode, iters in ((1, 8000), (0, 4000), (2, 8000), (3, 8000)):
    flop = ctypes.c_double(0.0)
    best = 0.0
    for rep in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = lib.dmvae_tf32_probe(iters, mode, _lib.ptr(sink), ctypes.byref(flop), _lib.stream_ptr())
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if rep: best = max(best, flop.value / (ms * 1e-3) / 1e12)
    n_mma = 16 * iters
    print(f"mode {mode}: rc={rc} {best:.1f} TFLOP/s, {ms*1e-3/n_mma*1.965e9:.1f} cycles per instruction at 1965 MHz", flush=True)
