"""Development aid: a few host-driven training steps at one batch size (ncu target).

    python scripts/one_step.py [B] [steps]
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "defensive-model-vae_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)
from dmvae import ConditionalTrajectoryVAE  # noqa: E402
from dmvae.train import FusedTrainer  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 8
torch.manual_seed(0)
model = ConditionalTrajectoryVAE(10, 3, 8).to("cuda")
tr = FusedTrainer(model, lr=1e-4)
x = torch.randn(B, 10, 3, device="cuda").cumsum(1)
for _ in range(steps):
    tr.step(x)
torch.cuda.synchronize()
print("ok", [float(v) for v in tr.losses.cpu()])
