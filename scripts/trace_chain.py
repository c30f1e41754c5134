"""Development aid: per-op clock64 timeline of chain_kernel (CTA 0, first tile)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "defensive-model-vae_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)
from dmvae import ConditionalTrajectoryVAE, _lib  # noqa: E402
from dmvae.train import FusedTrainer  # noqa: E402

OPS = ["enc0", "enc1", "enc2", "enc3", "heads_e", "cond0", "cond1", "heads_c", "dec0_c", "dec0_z", "dec1", "dec2", "dec3",
       "b_dec3", "b_dec2", "b_dec1", "b_dec0_z", "b_dec0_c", "b_heads_c", "b_cond1", "b_heads_e", "b_enc3", "b_enc2", "b_enc1"]
EPI_OF_OP = {}
e = 0
for i, n in enumerate(OPS):
    if n not in ("b_dec0_c", "heads_e", "dec0_c"):     # ops that do not commit: no epilogue of their own
        EPI_OF_OP[i] = e
        e += 1

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
TILE = int(sys.argv[2]) if len(sys.argv) > 2 else 0      # which of CTA 0's tiles (persistent CTAs: B > 148 * 128)
lib = _lib.lib()
torch.manual_seed(0)
model = ConditionalTrajectoryVAE(10, 3, 8).to("cuda")
tr = FusedTrainer(model, lr=1e-4)
x = torch.randn(B, 10, 3, device="cuda").cumsum(1)
for _ in range(3):
    tr.step(x)
buf = torch.zeros(256, dtype=torch.int64, device="cuda")
lib.dmvae_debug_train_trace_tile(_lib.ptr(buf), TILE)
tr.step(x)
torch.cuda.synchronize()
lib.dmvae_debug_train_trace(None)
t = buf.cpu().tolist()
t0 = t[0]
print(f"B={B}, tile {TILE} of CTA 0: per op [mma start -> issued] | [epilogue start -> released]  (cycles from first MMA start)")
for i, n in enumerate(OPS):
    ms, mi = t[4 * i] - t0, t[4 * i + 1] - t0
    line = f"{n:10s} mma {ms:7d} -> {mi:7d} (+{mi - ms:5d})"
    if i in EPI_OF_OP:
        e = EPI_OF_OP[i]
        es, er = t[128 + 2 * e] - t0, t[128 + 2 * e + 1] - t0
        line += f" | epi {es:7d} (+{es - mi:5d} after issue) -> {er:7d} (+{er - es:5d})"
    print(line)

# train_tc_fused_kernel (small batch): %globaltimer stamps of chain CTA 0 and the weight-gradient CTAs of tile 0
if t[176] and t[180]:
    g0 = t[176]
    print(f"fused launch: chain CTA 0 start 0 ns -> first MMA {t[178] - g0} -> last epilogue done {t[179] - g0} -> end {t[177] - g0} ns")
    for r in range(3):
        b = 180 + 16 * r
        ops = [t[b + 1 + o] - g0 for o in range(5) if t[b + 1 + o]]
        print(f"  wgrad role {r}: start {t[b] - g0}  ops ready at {ops}  accumulators done {t[b + 8] - g0}  written {t[b + 9] - g0}")

if t[240]:
    print("bdec0 epilogue (cycles): enter", 0, "accumulator ready", t[241] - t[240], "reparam/KLD backward done", t[242] - t[240],
          "operand + stash written", t[243] - t[240], "released", t[245] - t[240])

if t[246]:
    e = OPS.index("heads_c"); e = EPI_OF_OP[e]
    print("heads epilogue (cycles from its start): mu/logvar in shared memory", t[246] - t[128 + 2 * e], "released", t[128 + 2 * e + 1] - t[128 + 2 * e])
if t[248]:
    e = OPS.index("dec3"); e = EPI_OF_OP[e]
    print("loss epilogue (cycles from its start): recon in shared memory", t[248] - t[128 + 2 * e], "released", t[128 + 2 * e + 1] - t[128 + 2 * e])

if t[249]:
    print(f"start-up (ns from CTA start): barriers + tensor memory {t[249] - t[176]}, x tile + biases in shared memory {t[250] - t[176]}, "
          f"operands staged {t[251] - t[176]}, first MMA {t[178] - t[176]}")
