"""Host-side multi-GPU logic on CPU: world_size-2 ``gloo`` process groups.

The per-rank compute engine is the oracle here (the CUDA engine needs a GPU; its parity
is tests/test_train_gpu.py).  What is checked is the logic around it: contiguous
sharding, loss means over the GLOBAL batch, one SUM all-reduce of [grads | losses],
replicated Adam, the replica checksum, and the sharded ``.npy`` writer.
"""
import os
import socket
from collections import OrderedDict

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import vae_oracle as O

T, L = 10, 8


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _flat(d):
    return torch.cat([v.reshape(-1) for v in d.values()])


class OracleEngine:
    """GradEngine over the CPU oracle: this rank's SHARE of the global means."""

    def __init__(self, p, weights=O.SCRIPT_WEIGHTS, lr=1e-3):
        self.p = p
        self.weights = weights
        self.adam = O.AdamState(p, lr=lr)
        self.n_params = O.n_params(T, L)
        self.grad_buf = torch.zeros(self.n_params + 5)

    def loss_and_grads(self, batch, eps=None, global_batch=None, sample_offset=0):
        B = batch.shape[0]
        share = B / float(global_batch if global_batch is not None else B)
        losses, grads, _ = O.loss_and_grads(self.p, batch, eps, self.weights)
        self.grad_buf[: self.n_params] = _flat(grads) * share
        self.grad_buf[self.n_params:] = torch.tensor(losses) * share
        return self.grad_buf[self.n_params:], self.grad_buf[: self.n_params]

    def apply(self, grads=None):
        g = self.grad_buf if grads is None else grads
        out, off = OrderedDict(), 0
        for k, v in self.p.items():
            out[k] = g[off:off + v.numel()].view(v.shape)
            off += v.numel()
        self.adam.step(self.p, out)


def _batch(B, seed):
    g = torch.Generator().manual_seed(seed)
    t = torch.cumsum(torch.rand(B, T, generator=g) + 0.3, 1)
    t = t - t[:, :1]
    xy = torch.cumsum(torch.randn(B, T, 2, generator=g), 1) + (torch.rand(B, 1, 2, generator=g) - 0.5) * 80
    return torch.cat([t[..., None], xy], -1).contiguous()


def _dp_worker(rank, world, port, tmp, equal):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    torch.set_num_threads(1)
    from dmvae.parallel import DataParallelTrainer, init_distributed, shard_range, write_sharded_npy
    r, w, _ = init_distributed("gloo")
    assert (r, w) == (rank, world)
    B, steps = 96, 3
    batch = _batch(B, 5)
    eps = torch.randn(steps, B, L, generator=torch.Generator().manual_seed(6))
    cut = B // 2 if equal else 37
    lo, hi = (0, cut) if rank == 0 else (cut, B)
    if equal:
        assert (lo, hi) == shard_range(B, rank, world)
    p = O.init_params(T, L, seed=11)
    dp = DataParallelTrainer(OracleEngine(p))
    hist = []
    for s in range(steps):
        losses = dp.step(batch[lo:hi], eps=eps[s, lo:hi], equal_shards=equal)
        hist.append(losses.clone())
    assert dp.parameter_checksum(_flat(p))
    # a deliberately diverged replica is caught
    bad = _flat(p).clone()
    if rank == 1:
        bad[3] += 1e-3
    assert not dp.parameter_checksum(bad)
    # sharded writer: every rank writes its slab of one file
    rows = np.full((hi - lo, T, 3), float(rank + 1), dtype=np.float32)
    write_sharded_npy(os.path.join(tmp, "shards.npy"), rows, lo, B, rank, world)
    torch.save({"p": p, "hist": torch.stack(hist)}, os.path.join(tmp, f"rank{rank}_{int(equal)}.pt"))
    dist.destroy_process_group()


@pytest.mark.parametrize("equal", [True, False])
def test_data_parallel_two_ranks_match_single_process(tmp_path, equal):
    port = _free_port()
    mp.spawn(_dp_worker, args=(2, port, str(tmp_path), equal), nprocs=2, join=True)
    B, steps = 96, 3
    batch = _batch(B, 5)
    eps = torch.randn(steps, B, L, generator=torch.Generator().manual_seed(6))
    p = O.init_params(T, L, seed=11)
    hist, _ = O.train_steps(p, batch, eps, O.SCRIPT_WEIGHTS, lr=1e-3)
    r0 = torch.load(os.path.join(tmp_path, f"rank0_{int(equal)}.pt"), weights_only=False)
    r1 = torch.load(os.path.join(tmp_path, f"rank1_{int(equal)}.pt"), weights_only=False)
    for k in p:
        assert torch.equal(r0["p"][k], r1["p"][k]), k                      # replicas in lock-step
        scale = p[k].abs().max().item()
        # Adam normalises by sqrt(v): an element whose gradient is ~0 may move by a visible fraction of
        # lr (1e-3) when the summation order differs; 2e-5 absolute = 2 % of one step
        assert (r0["p"][k] - p[k]).abs().max().item() <= 2e-5 * scale + 2e-5, k
    np.testing.assert_allclose(r0["hist"].numpy(), hist, rtol=2e-5, atol=1e-7)  # global-batch means
    shards = np.load(os.path.join(tmp_path, "shards.npy"))
    cut = B // 2 if equal else 37
    assert shards.shape == (B, T, 3) and (shards[:cut] == 1).all() and (shards[cut:] == 2).all()


def test_shard_range_partitions_exactly():
    from dmvae.parallel import shard_range
    for n in (0, 1, 7, 1000, 1 << 20, 1_000_003):
        for world in (1, 2, 3, 4, 8):
            edges = [shard_range(n, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in edges]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def test_single_process_trainer_needs_no_process_group():
    from dmvae.parallel import DataParallelTrainer
    p = O.init_params(T, L, seed=2)
    q = O.clone_params(p)
    dp = DataParallelTrainer(OracleEngine(p))
    batch, eps = _batch(20, 1), torch.randn(1, 20, L, generator=torch.Generator().manual_seed(3))
    dp.step(batch, eps=eps[0])
    O.train_steps(q, batch, eps)
    for k in p:
        assert torch.equal(p[k], q[k])
    assert dp.parameter_checksum(_flat(p))
