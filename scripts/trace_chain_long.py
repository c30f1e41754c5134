"""Development aid: clock64 timeline of one tile of chain_kernel<long> for a two-chunk trajectory (3 T in 129..256):
per op [first MMA -> issued], per epilogue [accumulator awaited -> A operand released].  python scripts/trace_chain_long.py [T] [B]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "defensive-model-vae_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)
from dmvae import ConditionalTrajectoryVAE, _lib  # noqa: E402
from dmvae.train import FusedTrainer  # noqa: E402

T = int(sys.argv[1]) if len(sys.argv) > 1 else 85
B = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
assert 128 < 3 * T <= 256, "the trace buffer holds the stamps of two chunks"
# chain_program_long, NC = 2 (dmvae_train_tc.cu)
OPS = ["enc0.c0", "enc0.c1", "enc1", "enc2", "enc3", "heads_e", "cond0", "cond1", "heads_c", "dec0_c", "dec0_z", "dec1", "dec2",
       "dec3.c0", "dec3.c1", "b_dec3.c0", "b_dec3.c1", "b_dec2", "b_dec1", "b_dec0_z", "b_dec0_c", "b_heads_c", "b_cond1",
       "b_heads_e", "b_enc3", "b_enc2", "b_enc1"]
EPIS = ["stage x.c1", "enc0", "enc1", "enc2", "enc3", "cond0", "cond1", "heads", "dec0", "dec1", "dec2", "loss.c0", "loss.c1 (+ stage g.c0)",
        "stage g.c1", "b_dec3", "b_dec2", "b_dec1", "b_dec0_z", "b_heads_c", "b_cond1", "b_heads_e", "b_enc3", "b_enc2", "b_enc1"]
lib = _lib.lib()
torch.manual_seed(0)
model = ConditionalTrajectoryVAE(T, 3, 8).to("cuda")
tr = FusedTrainer(model, lr=1e-4)
x = torch.randn(B, T, 3, device="cuda").cumsum(1)
for _ in range(3):
    tr.step(x)
buf = torch.zeros(256, dtype=torch.int64, device="cuda")
lib.dmvae_debug_train_trace_tile(_lib.ptr(buf), 1)        # the second tile of CTA 0 (steady state)
tr.step(x)
torch.cuda.synchronize()
lib.dmvae_debug_train_trace(None)
t = buf.cpu().tolist()
t0 = t[0]
print(f"T={T} B={B}: second tile of CTA 0, cycles from its first MMA")
for i, n in enumerate(OPS):
    if t[4 * i + 1]:
        print(f"op  {n:10s} mma {t[4 * i] - t0:8d} -> {t[4 * i + 1] - t0:8d} (+{t[4 * i + 1] - t[4 * i]:6d})")
for e, n in enumerate(EPIS):
    if t[128 + 2 * e + 1]:
        print(f"epi {n:24s} {t[128 + 2 * e] - t0:8d} -> {t[128 + 2 * e + 1] - t0:8d} (+{t[128 + 2 * e + 1] - t[128 + 2 * e]:6d})")
