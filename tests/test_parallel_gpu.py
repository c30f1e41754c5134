"""Multi-GPU paths on the device.  Single-GPU boxes run the shard-invariance
tests (ranks emulated one after the other); with >= 2 GPUs the NCCL data-parallel step and
the sharded writer run as real 2-rank jobs."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from oracle import vae_oracle as O

pytestmark = pytest.mark.gpu
T, L = 10, 8


def _model(seed=3):
    from dmvae import ConditionalTrajectoryVAE
    p = O.init_params(T, L, seed=seed)
    m = ConditionalTrajectoryVAE(T, 3, L)
    m.load_state_dict({k: v.clone() for k, v in p.items()})
    return m.to("cuda").eval(), p


def _batch(B, seed):
    g = torch.Generator().manual_seed(seed)
    t = torch.cumsum(torch.rand(B, T, generator=g) + 0.3, 1)
    t = t - t[:, :1]
    xy = torch.cumsum(torch.randn(B, T, 2, generator=g), 1) + (torch.rand(B, 1, 2, generator=g) - 0.5) * 80
    return torch.cat([t[..., None], xy], -1).contiguous()


@pytest.mark.parametrize("world", [2, 3, 8])
@pytest.mark.parametrize("shared", [True, False])
def test_generation_is_bit_identical_for_any_sharding(world, shared):
    from dmvae.parallel import generate_shard
    model, _ = _model()
    n = 100_003
    g = torch.Generator().manual_seed(1)
    starts = [[11.0, 0.0]] if shared else (torch.rand(n, 2, generator=g) * 300 - 150).numpy()
    _, _, whole = generate_shard(model, starts, n, seed=42, rank=0, world=1)
    parts = []
    for r in range(world):
        lo, hi, out = generate_shard(model, starts, n, seed=42, rank=r, world=world, chunk=20_000)
        assert out.shape[0] == hi - lo
        parts.append(out)
    assert torch.equal(torch.cat(parts, 0), whole)


def test_generation_philox_moments():
    """In-kernel Philox latents are validated statistically (RNG parity with torch's CPU
    generator is impossible by construction - SURVEY.md section 7)."""
    model, p = _model()
    n = 1 << 20
    out, z = model.generate(torch.tensor([[11.0, 0.0]]), n=n, seed=7, return_z=True)
    z = z.double().cpu()
    assert abs(z.mean().item()) < 5e-3 and abs(z.var().item() - 1.0) < 5e-3
    assert abs((z ** 3).mean().item()) < 2e-2 and abs((z ** 4).mean().item() - 3.0) < 5e-2
    c = np.corrcoef(z[:200_000].numpy().T)
    assert np.abs(c - np.eye(L)).max() < 1e-2
    # and the decoded rows are the oracle's decode of those very latents
    ref = O.generate(p, z[:4096].float(), torch.tensor([[11.0, 0.0]]))
    got = out[:4096].cpu()
    assert (got - ref).abs().max().item() <= 1e-5 * ref.abs().max().item()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _nccl_worker(rank, world, port, tmp, exchange):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import torch.distributed as dist
    from dmvae.parallel import DataParallelTrainer, generate_shard, init_distributed, shard_range, write_sharded_npy
    from dmvae.train import FusedTrainer
    init_distributed("nccl")
    model, _ = _model(seed=5)
    B, steps = 4096, 4
    batch = _batch(B, 9)
    eps = torch.randn(steps, B, L, generator=torch.Generator().manual_seed(10))
    lo, hi = shard_range(B, rank, world)
    dp = DataParallelTrainer(FusedTrainer(model, lr=1e-3, weights=O.SCRIPT_WEIGHTS), exchange=exchange)
    assert dp.exchange == exchange, dp.exchange_note
    hist = []
    for s in range(steps):
        hist.append(dp.step(batch[lo:hi].cuda(), eps=eps[s, lo:hi].cuda()).cpu().clone())
    assert dp.parameter_checksum(model.flat_parameters())
    glo, ghi, out = generate_shard(model, [[11.0, 0.0]], 50_001, seed=3, rank=rank, world=world)
    write_sharded_npy(os.path.join(tmp, "gen.npy"), out.cpu().numpy(), glo, 50_001, rank, world)
    if rank == 0:
        torch.save({"hist": torch.stack(hist), "params": model.flat_parameters().cpu()}, os.path.join(tmp, "dp.pt"))
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
@pytest.mark.parametrize("exchange", ["peer", "nccl"])
def test_nccl_data_parallel_matches_single_gpu(tmp_path, exchange):
    """exchange = "peer": the gradient exchange runs inside the update kernel over peer memory
    (dmvae_train_step_dp); "nccl": one all-reduce between the fused pass and the Adam kernel."""
    from dmvae.parallel import generate_shard
    from dmvae.train import FusedTrainer
    port = _free_port()
    mp.spawn(_nccl_worker, args=(2, port, str(tmp_path), exchange), nprocs=2, join=True)
    got = torch.load(os.path.join(tmp_path, "dp.pt"), weights_only=False)
    model, _ = _model(seed=5)
    B, steps = 4096, 4
    batch = _batch(B, 9).cuda()
    eps = torch.randn(steps, B, L, generator=torch.Generator().manual_seed(10)).cuda()
    tr = FusedTrainer(model, lr=1e-3, weights=O.SCRIPT_WEIGHTS)
    hist = torch.stack([tr.step(batch, eps=eps[s]).cpu().clone() for s in range(steps)])
    np.testing.assert_allclose(got["hist"].numpy(), hist.numpy(), rtol=1e-5, atol=1e-7)
    ref = model.flat_parameters().cpu()
    assert (got["params"] - ref).abs().max().item() <= 2e-5 * ref.abs().max().item() + 2e-5
    # sharded generation written by two ranks == one rank's output (the 2-rank job used the
    # weights after training; regenerate with the same weights here)
    model.flat_parameters().copy_(got["params"].cuda())
    _, _, whole = generate_shard(model, [[11.0, 0.0]], 50_001, seed=3, rank=0, world=1)
    assert np.array_equal(np.load(os.path.join(tmp_path, "gen.npy")), whole.cpu().numpy())


def _nccl_graph_worker(rank, world, port, tmp, exchange):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import torch.distributed as dist
    from dmvae.parallel import DataParallelTrainer, init_distributed, shard_range
    from dmvae.train import FusedTrainer
    init_distributed("nccl")
    if exchange == "peer-owned":
        from dmvae import _lib
        _lib.check(_lib.lib().dmvae_set_dp_owned_from(2), "dmvae_set_dp_owned_from")
        exchange = "peer"
    B, steps = 2048, 4
    batches = [_batch(B, 20 + s) for s in range(steps)]
    lo, hi = shard_range(B, rank, world)
    # host-driven data-parallel steps (Philox noise, keyed by the global row index) ...
    model_a, _ = _model(seed=6)
    dpa = DataParallelTrainer(FusedTrainer(model_a, lr=1e-3, weights=O.SCRIPT_WEIGHTS, seed=77), exchange=exchange)
    assert dpa.exchange == exchange, dpa.exchange_note
    ha = [dpa.step(b[lo:hi].cuda()).cpu().clone() for b in batches]
    # ... against the same steps replayed from one CUDA graph (all-reduce captured inside)
    model_b, _ = _model(seed=6)
    dpb = DataParallelTrainer(FusedTrainer(model_b, lr=1e-3, weights=O.SCRIPT_WEIGHTS, seed=77), exchange=exchange)
    gs = dpb.capture(hi - lo)
    hb = []
    for rep in range(6):      # the same batches again and again: many back-to-back replays exercise the flags
        for b in batches:
            gs.batch.copy_(b[lo:hi])
            out = gs.replay()
            if rep == 0:
                hb.append(out.cpu().clone())
    assert dpb.parameter_checksum(model_b.flat_parameters())
    np.testing.assert_allclose(torch.stack(hb).numpy(), torch.stack(ha).numpy(), rtol=1e-6)
    for rep in range(5):
        for b in batches:
            dpa.step(b[lo:hi].cuda())
    assert dpa.parameter_checksum(model_a.flat_parameters())
    pa, pb = model_a.flat_parameters().cpu(), model_b.flat_parameters().cpu()
    assert (pa - pb).abs().max().item() <= 1e-6 * pa.abs().max().item()
    if rank == 0:
        open(os.path.join(tmp, "ok"), "w").write("ok")
    # the graph holds captured NCCL work: release it before the communicator goes away
    gs.graph.reset()
    del gs
    torch.cuda.synchronize()
    dist.barrier()
    os._exit(0)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
@pytest.mark.parametrize("exchange", ["peer", "peer-owned", "nccl"])
def test_nccl_graph_step_matches_host_driven_step(tmp_path, exchange):
    """"peer-owned": the owner scheme of the exchange (the default from 3 ranks on) forced onto two ranks."""
    port = _free_port()
    mp.spawn(_nccl_graph_worker, args=(2, port, str(tmp_path), exchange), nprocs=2, join=True)
    assert os.path.isfile(os.path.join(tmp_path, "ok"))
