#!/usr/bin/env python
"""Bulk generation for the four defensive scenarios, sharded over the GPUs of one node
(BASELINE configs[2]):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        scripts/generate_sharded.py --n 1000000 --models training/models/vae_offset_sce1_cond_ld8_epoch3000.pth ...

Each rank decodes a contiguous slice of the global row range; rows are keyed by their GLOBAL
index in the Philox stream, so the files are byte-identical for any GPU count."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "defensive-model-vae_b200"))

from dmvae.parallel import generate_scenarios  # noqa: E402

DEFAULT_START = {"sce1": (-193.3, 50.0), "sce2": (-155.0, -5.0), "sce3": (155.0, -15.0), "sce4": (11.0, 0.0)}  # Tools.py:101-108


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--models", nargs="+", required=True, help="state_dict checkpoints (.pth), one per scenario")
    ap.add_argument("--scenarios", nargs="+", default=None, help="names (default: the sceN token of each file name)")
    ap.add_argument("--n", type=int, default=1_000_000)
    ap.add_argument("--out-dir", default="results/GeneratedData")
    ap.add_argument("--seq-len", type=int, default=10)
    ap.add_argument("--latent-dim", type=int, default=8)
    ap.add_argument("--seed", type=int, default=0)
    a = ap.parse_args()
    names = a.scenarios or [next((t for t in os.path.basename(m).split("_") if t.startswith("sce")), f"model{i}")
                            for i, m in enumerate(a.models)]
    starts = [DEFAULT_START.get(n, DEFAULT_START["sce3"]) for n in names]
    for p in generate_scenarios(a.models, names, starts, a.n, a.out_dir, a.seq_len, a.latent_dim, a.seed):
        if int(os.environ.get("RANK", "0")) == 0:
            print(p)


if __name__ == "__main__":
    main()
