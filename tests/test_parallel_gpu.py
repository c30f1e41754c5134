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


def _exchange_args(exchange):
    """test id -> (DataParallelTrainer exchange, owned_from): "peer" = the library default (all-to-all on two
    ranks, element owners from three on), "peer-owned" / "peer-all" force one scheme at any world size."""
    return {"peer": ("peer", 0), "peer-owned": ("peer", 2), "peer-all": ("peer", 9), "nccl": ("nccl", 0)}[exchange]


def _need_gpus(world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs (gpurun --gpus {world})")


def _dp_worker(rank, world, port, tmp, exchange):
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
    kind, owned_from = _exchange_args(exchange)
    dp = DataParallelTrainer(FusedTrainer(model, lr=1e-3, weights=O.SCRIPT_WEIGHTS), exchange=kind, owned_from=owned_from)
    assert dp.exchange == kind, dp.exchange_note
    hist = []
    for s in range(steps):
        hist.append(dp.step(batch[lo:hi].cuda(), eps=eps[s, lo:hi].cuda()).cpu().clone())
    dp.check_exchange()
    assert dp.parameter_checksum(model.flat_parameters())
    glo, ghi, out = generate_shard(model, [[11.0, 0.0]], 50_001, seed=3, rank=rank, world=world)
    write_sharded_npy(os.path.join(tmp, "gen.npy"), out.cpu().numpy(), glo, 50_001, rank, world)
    if rank == 0:
        torch.save({"hist": torch.stack(hist), "params": model.flat_parameters().cpu(),
                    "grads": dp.engine.grad_buf.cpu()}, os.path.join(tmp, "dp.pt"))
    dist.destroy_process_group()


@pytest.mark.parametrize("exchange", ["peer", "peer-owned", "peer-all", "nccl"])
@pytest.mark.parametrize("world", [2, 4, 8])
def test_data_parallel_matches_single_gpu(tmp_path, world, exchange):
    """The sum is RIGHT, not merely the same on every rank: `world` ranks on slices of one global batch against one
    rank on the whole batch (same injected noise), losses of every step, the gradient of the last step and the
    parameters after four updates.  exchange = "peer*": the gradient exchange runs inside the update kernel over
    peer memory (dmvae_train_step_dp; default scheme, element owners forced, all-to-all forced); "nccl": one
    all-reduce between the fused pass and the Adam kernel."""
    _need_gpus(world)
    from dmvae.parallel import generate_shard
    from dmvae.train import FusedTrainer
    port = _free_port()
    mp.spawn(_dp_worker, args=(world, port, str(tmp_path), exchange), nprocs=world, join=True)
    got = torch.load(os.path.join(tmp_path, "dp.pt"), weights_only=False)
    model, _ = _model(seed=5)
    B, steps = 4096, 4
    batch = _batch(B, 9).cuda()
    eps = torch.randn(steps, B, L, generator=torch.Generator().manual_seed(10)).cuda()
    tr = FusedTrainer(model, lr=1e-3, weights=O.SCRIPT_WEIGHTS)
    hist = torch.stack([tr.step(batch, eps=eps[s]).cpu().clone() for s in range(steps)])
    np.testing.assert_allclose(got["hist"].numpy(), hist.numpy(), rtol=1e-5, atol=1e-7)
    gref, ggot = tr.grad_buf.cpu()[: tr.n_params], got["grads"][: tr.n_params]
    assert (ggot - gref).abs().max().item() <= 2e-5 * gref.abs().max().item()
    ref = model.flat_parameters().cpu()
    assert (got["params"] - ref).abs().max().item() <= 2e-5 * ref.abs().max().item() + 2e-5
    # sharded generation written by the ranks == one rank's output (the job used the weights after training;
    # regenerate with the same weights here)
    model.flat_parameters().copy_(got["params"].cuda())
    _, _, whole = generate_shard(model, [[11.0, 0.0]], 50_001, seed=3, rank=0, world=1)
    assert np.array_equal(np.load(os.path.join(tmp_path, "gen.npy")), whole.cpu().numpy())


def _dp_graph_worker(rank, world, port, tmp, exchange):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import torch.distributed as dist
    from dmvae.parallel import DataParallelTrainer, init_distributed, shard_range
    from dmvae.train import FusedTrainer
    init_distributed("nccl")
    kind, owned_from = _exchange_args(exchange)
    B, steps = 4096, 4
    batches = [_batch(B, 20 + s) for s in range(steps)]
    lo, hi = shard_range(B, rank, world)
    # host-driven data-parallel steps (Philox noise, keyed by the global row index) ...
    model_a, _ = _model(seed=6)
    dpa = DataParallelTrainer(FusedTrainer(model_a, lr=1e-3, weights=O.SCRIPT_WEIGHTS, seed=77), exchange=kind,
                              owned_from=owned_from)
    assert dpa.exchange == kind, dpa.exchange_note
    ha = [dpa.step(b[lo:hi].cuda()).cpu().clone() for b in batches]
    pa4 = model_a.flat_parameters().cpu().clone()
    # ... against the same steps replayed from one CUDA graph (exchange / all-reduce captured inside)
    model_b, _ = _model(seed=6)
    dpb = DataParallelTrainer(FusedTrainer(model_b, lr=1e-3, weights=O.SCRIPT_WEIGHTS, seed=77), exchange=kind,
                              owned_from=owned_from)
    gs = dpb.capture(hi - lo)
    hb = []
    for rep in range(6):      # the same batches again and again: many back-to-back replays exercise the words' step tags
        for b in batches:
            gs.batch.copy_(b[lo:hi])
            out = gs.replay()
            if rep == 0:
                hb.append(out.cpu().clone())
    dpb.check_exchange()
    assert dpb.parameter_checksum(model_b.flat_parameters())
    np.testing.assert_allclose(torch.stack(hb).numpy(), torch.stack(ha).numpy(), rtol=1e-6)
    for rep in range(5):
        for b in batches:
            dpa.step(b[lo:hi].cuda())
    dpa.check_exchange()
    assert dpa.parameter_checksum(model_a.flat_parameters())
    pa, pb = model_a.flat_parameters().cpu(), model_b.flat_parameters().cpu()
    assert (pa - pb).abs().max().item() <= 1e-6 * pa.abs().max().item()
    if kind == "peer":
        # ... and against a resident data set: every rank keeps its slices of all batches in HBM, the kernel picks
        # the slice of update t from the device-side step counter (dmvae_train_step_resident with peers)
        model_c, _ = _model(seed=6)
        dpc = DataParallelTrainer(FusedTrainer(model_c, lr=1e-3, weights=O.SCRIPT_WEIGHTS, seed=77), exchange=kind,
                                  owned_from=owned_from)
        shard = torch.cat([b[lo:hi] for b in batches], 0).cuda()
        gc = dpc.capture(hi - lo, dataset=shard)
        hc = [gc.replay().cpu().clone() for _ in range(steps)]
        np.testing.assert_allclose(torch.stack(hc).numpy(), torch.stack(ha).numpy(), rtol=1e-6)
        for _ in range(5 * steps):
            gc.replay()
        dpc.check_exchange()
        assert dpc.parameter_checksum(model_c.flat_parameters())
        pc = model_c.flat_parameters().cpu()
        assert (pa - pc).abs().max().item() <= 1e-6 * pa.abs().max().item()
        gc.graph.reset()
    if rank == 0:
        torch.save({"params": pa, "params4": pa4, "hist": torch.stack(ha)}, os.path.join(tmp, "graph.pt"))
    # the graph holds captured NCCL work: release it before the communicator goes away
    gs.graph.reset()
    del gs
    torch.cuda.synchronize()
    dist.barrier()
    os._exit(0)


@pytest.mark.parametrize("exchange", ["peer", "peer-owned", "nccl"])
@pytest.mark.parametrize("world", [2, 4, 8])
def test_graph_step_matches_host_driven_step(tmp_path, world, exchange):
    """Host-driven data-parallel steps == the same steps replayed from one CUDA graph == (peer exchange) the
    steps over a device-resident shard; and all of them == ONE rank on the whole batches.
    "peer-owned": the owner scheme of the exchange (the default from 3 ranks on) forced onto two ranks."""
    _need_gpus(world)
    from dmvae.train import FusedTrainer
    port = _free_port()
    mp.spawn(_dp_graph_worker, args=(world, port, str(tmp_path), exchange), nprocs=world, join=True)
    got = torch.load(os.path.join(tmp_path, "graph.pt"), weights_only=False)
    # the same 4 + 20 steps on one GPU over the whole batches (Philox noise is keyed by the global row index)
    model, _ = _model(seed=6)
    tr = FusedTrainer(model, lr=1e-3, weights=O.SCRIPT_WEIGHTS, seed=77)
    B, steps = 4096, 4
    batches = [_batch(B, 20 + s).cuda() for s in range(steps)]
    hist = torch.stack([tr.step(b).cpu().clone() for b in batches])
    np.testing.assert_allclose(got["hist"].numpy(), hist.numpy(), rtol=1e-5, atol=1e-7)
    ref4 = model.flat_parameters().cpu().clone()
    assert (got["params4"] - ref4).abs().max().item() <= 2e-5 * ref4.abs().max().item() + 2e-5
    for rep in range(5):
        for b in batches:
            tr.step(b)
    # 24 updates: Adam turns a rounding-level difference of a near-zero gradient into a +-lr difference of that one
    # parameter (update = lr * g / (|g| + eps)), so single elements drift apart; the update as a whole must not
    ref = model.flat_parameters().cpu()
    p0 = _model(seed=6)[0].flat_parameters().cpu()
    drift = ((got["params"] - ref).norm() / (ref - p0).norm()).item()
    assert drift < 0.02, drift


def _dp_timeout_worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import torch.distributed as dist
    from dmvae import _lib
    from dmvae.parallel import DataParallelTrainer, init_distributed
    from dmvae.train import FusedTrainer
    init_distributed("nccl")
    model, _ = _model(seed=5)
    dp = DataParallelTrainer(FusedTrainer(model, lr=1e-3, weights=O.SCRIPT_WEIGHTS), exchange="peer", timeout_ms=300)
    dp.step(_batch(256, 1).cuda())
    dp.check_exchange()                      # a normal step: clean status
    if rank == 0:                            # rank 1 never takes the second step
        dp.step(_batch(256, 2).cuda())
        torch.cuda.synchronize()             # the kernel ends by itself (no hang) ...
        with pytest.raises(_lib.DmvaeError, match="gave up waiting"):
            dp.check_exchange()              # ... and the status word says why
        assert not torch.isfinite(model.flat_parameters()).all()
        open(os.path.join(tmp, "ok"), "w").write("ok")
    dist.barrier()
    dist.destroy_process_group()


def test_peer_exchange_times_out_instead_of_hanging(tmp_path):
    """A peer that never delivers (dead, a step behind) must not hang the GPU: the polling threads give up after
    timeout_ms, the rank's parameters turn NaN and dmvae_dp_status reports the step."""
    _need_gpus(2)
    port = _free_port()
    mp.spawn(_dp_timeout_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert os.path.isfile(os.path.join(tmp_path, "ok"))
