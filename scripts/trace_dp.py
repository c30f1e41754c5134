"""Development aid (torchrun, 2+ GPUs): %globaltimer stamps of block 0 of the data-parallel update kernel."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "defensive-model-vae_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)
from dmvae import ConditionalTrajectoryVAE, _lib  # noqa: E402
from dmvae.parallel import DataParallelTrainer, init_distributed  # noqa: E402
from dmvae.train import FusedTrainer  # noqa: E402

rank, world, local = init_distributed("nccl")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
lib = _lib.lib()
torch.manual_seed(0)
model = ConditionalTrajectoryVAE(10, 3, 8).to("cuda")
dp = DataParallelTrainer(FusedTrainer(model, lr=1e-4), exchange="peer")
x = torch.randn(B, 10, 3, device="cuda").cumsum(1)
gs = dp.capture(B)
gs.batch.copy_(x)
for _ in range(20):
    gs.replay()
torch.cuda.synchronize()
buf = torch.zeros(256, dtype=torch.int64, device="cuda")
for rep in range(3):
    lib.dmvae_debug_train_trace(_lib.ptr(buf))
    dp.step(x)                      # host-driven step (the trace pointer is a launch parameter)
    torch.cuda.synchronize()
    lib.dmvae_debug_train_trace(None)
    t = buf.cpu().tolist()
    g0 = t[176]
    print(f"rank {rank} rep {rep}: fused start 0, chain end {t[177] - g0}; update kernel block 0: start {t[230] - g0}, "
          f"own sum +{t[231] - t[230]}, pushed + pulled +{t[233] - t[231]}, adam + pack +{t[234] - t[233]} ns", flush=True)
torch.distributed.barrier()
os._exit(0)
