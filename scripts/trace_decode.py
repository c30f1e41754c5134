"""Development aid: per-layer-step clock stamps of CTA 0 of decode_tc_kernel."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "defensive-model-vae_b200"))
import torch
from dmvae import ConditionalTrajectoryVAE, _lib
torch.manual_seed(0)
m = ConditionalTrajectoryVAE(10, 3, 8).to("cuda").eval()
lib = _lib.lib()
B = 148 * 128 * 8
for mode in ("shared", "per-row"):
    start = torch.rand(1 if mode == "shared" else B, 2, device="cuda") * 100
    out = torch.empty(B, 10, 3, device="cuda")
    m.generate(start, n=B, out=out)
    tr = torch.zeros(128, dtype=torch.int64, device="cuda")
    lib.dmvae_debug_decode_trace(_lib.ptr(tr))
    m.generate(start, n=B, out=out)
    torch.cuda.synchronize()
    lib.dmvae_debug_decode_trace(None)
    t = tr.cpu().view(4, 8, 4)
    t0 = int(t[0, 0, 0])
    print(f"== {mode}: per tile/op: [mma start, mma issued] [epi start, epi published] (cycles from first MMA start)")
    for ti in range(4):
        for o in range(8):
            if int(t[ti, o, 0]) == 0: continue
            a, b, c, d = (int(x) - t0 for x in t[ti, o])
            print(f"tile {ti} op {o}: mma {a:7d} -> issued {b:7d} (+{b-a:5d}) | epi {c:7d} (+{c-b:5d} after issue) -> published {d:7d} (+{d-c:5d})")
