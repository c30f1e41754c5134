"""Multi-GPU host logic: one process per GPU, ``torch.distributed``.

* Generation shards trivially: rows are independent, rank r decodes the contiguous
  global index range ``shard_range(n, r, world)``; the in-kernel Philox counter is the
  GLOBAL row index, so the bytes written are identical for any world size.  No collective.
* Training is data parallel: every rank runs the fused forward+loss+backward on its slice
  of the global batch with the loss means scaled by 1/B_global, then the flat fp32 buffer
  [gradients (128 942 at T=10, L=8) | 5 loss terms] is summed over the ranks, then the
  identical replicated Adam update.  Two exchanges:
    "peer"  (GPUs of one node, the default where it can be set up) every thread of the update
            kernel writes its slab sum, tagged with the step, straight into the peers' inboxes
            (torch symmetric memory = CUDA peer access over NVLink), polls its own inbox for
            theirs and adds the ranks in rank order: sum, exchange and update are one kernel,
            no library collective, no fence (``dmvae_train_step_dp``);
    "nccl"  ONE ``all_reduce(SUM)`` between the fused pass and the Adam kernel (also the
            path of the CPU tests, with gloo).
  Either way every rank adds the same numbers in the same order, so the replicas stay
  bit-identical; ``parameter_checksum`` asserts it.

The reference has no distributed code at all (single process, Training_VAE.py:327); this
module is the capability the north star adds.  The compute engine is injected (``engine``)
so that the host logic - scaling, collective, update order - is exercised on CPU with the
``gloo`` backend in tests/ (engine = the oracle there) and with NCCL on the GPUs
(engine = ``FusedTrainer``).
"""
from __future__ import annotations

import os
from typing import Optional, Protocol, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous partition of range(n): rank r owns [r*n//world, (r+1)*n//world)."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    return (rank * n) // world, ((rank + 1) * n) // world


def world_info() -> Tuple[int, int, int]:
    """(rank, world_size, local_rank) from the launcher's environment (torchrun)."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def init_distributed(backend: Optional[str] = None) -> Tuple[int, int, int]:
    """Initialise the default process group from the environment when WORLD_SIZE > 1
    (``nccl`` on GPUs, ``gloo`` otherwise) and bind this process to its GPU."""
    rank, world, local = world_info()
    if torch.cuda.is_available():
        torch.cuda.set_device(local)
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        kwargs = {"device_id": torch.device("cuda", local)} if backend == "nccl" else {}
        dist.init_process_group(backend, rank=rank, world_size=world, **kwargs)
    return rank, world, local


class GradEngine(Protocol):
    """What DataParallelTrainer needs from the per-rank compute engine
    (``dmvae.train.FusedTrainer`` provides exactly this)."""
    grad_buf: torch.Tensor       # flat fp32 [grads | total, recon, kld, start, time]
    n_params: int

    def loss_and_grads(self, batch, eps=None, global_batch=None, sample_offset=0): ...
    def apply(self, grads=None) -> None: ...


class DataParallelTrainer:
    """Synchronous data-parallel training step over ``torch.distributed``.

    ``step(local_batch)``: the rank's slice of the global batch (all ranks pass slices of
    the same step; sizes may differ).  Equivalent to one single-process step on the
    concatenated batch: same loss (means over the global batch), same update."""

    def __init__(self, engine: GradEngine, group=None, exchange: str = "auto", owned_from: int = 0,
                 timeout_ms: int = 0):
        """``exchange``: "peer" (gradient exchange inside the update kernel over peer memory; raises where it
        cannot be set up), "nccl" (one all-reduce per step), or "auto" (peer when possible, else nccl - the
        reason is kept in ``exchange_note``).  ``owned_from`` / ``timeout_ms``: the fields of ``DmvaeDpPeers``
        (0 = the library defaults: owner scheme from 3 ranks on, 2 s poll limit).  All ranks must pass the
        same values; the constructor checks that they do."""
        if exchange not in ("auto", "peer", "nccl"):
            raise ValueError(f"exchange={exchange!r}")
        self.engine = engine
        self.group = group
        self.owned_from, self.timeout_ms = int(owned_from), int(timeout_ms)
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self._sizes = torch.zeros(self.world, dtype=torch.int64)
        self.peers = None
        self.exchange_note = ""
        if self.world > 1 and exchange != "nccl":
            try:
                self._setup_peers()
            except Exception as e:  # noqa: BLE001 - any failure of the optional path selects the collective
                self.peers = None
                self.exchange_note = f"peer exchange unavailable ({type(e).__name__}: {e}); using all_reduce"
                ok = torch.zeros(1, device=self._device())
            else:
                ok = torch.ones(1, device=self._device())
            # the choice must be unanimous: a rank that could not map its peers sends everyone to the collective
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)
            if ok.item() < 1:
                if exchange == "peer":
                    raise RuntimeError(self.exchange_note or "peer exchange unavailable on another rank")
                self.peers = None
            # every rank must describe the exchange identically: a rank on another scheme would wait for words that
            # nobody sends (the kernel's poll limit would end that wait, but the step would be lost)
            mine = [self.owned_from, -self.owned_from, self.timeout_ms, -self.timeout_ms]
            seen = torch.tensor(mine, dtype=torch.int64, device=self._device())
            dist.all_reduce(seen, op=dist.ReduceOp.MAX, group=self.group)
            if seen.tolist() != mine:
                raise RuntimeError("the ranks disagree about owned_from / timeout_ms of the peer exchange")
        self.exchange = "peer" if self.peers is not None else ("nccl" if self.world > 1 else "none")

    def _device(self):
        buf = getattr(self.engine, "grad_buf", None)
        return buf.device if buf is not None else torch.device("cpu")

    def _setup_peers(self) -> None:
        """Allocates this rank's inbox in symmetric memory and maps every peer's."""
        from ._lib import MAX_PEERS, DmvaeDpPeers, check
        eng = self.engine
        lib, cfg_ref = getattr(eng, "lib", None), getattr(eng, "_cfg_ref", None)
        if lib is None or cfg_ref is None or not self._device().type == "cuda":
            raise RuntimeError("the engine is not a CUDA FusedTrainer")
        if self.world > MAX_PEERS:
            raise RuntimeError(f"{self.world} ranks, at most {MAX_PEERS} peers")
        import torch.distributed._symmetric_memory as symm_mem
        dev = self._device()
        with torch.cuda.device(dev):
            nbytes = check(lib.dmvae_dp_inbox_bytes(cfg_ref, self.world), "dmvae_dp_inbox_bytes")
            group = self.group if self.group is not None else dist.group.WORLD
            buf = symm_mem.empty(nbytes // 4, dtype=torch.float32, device=dev)
            hdl = symm_mem.rendezvous(buf, group)
            ptrs = [int(p) for p in hdl.buffer_ptrs]
        if len(ptrs) != self.world or int(hdl.rank) != self.rank:
            raise RuntimeError("symmetric memory handle does not match the process group")
        peers = DmvaeDpPeers()
        peers.world, peers.rank = self.world, self.rank
        peers.owned_from, peers.timeout_ms = self.owned_from, self.timeout_ms
        for p in range(self.world):
            peers.inbox[p] = ptrs[p]
        self._symm = (buf, hdl)
        self.peers = peers
        self._reset_flags()

    def _reset_flags(self) -> None:
        """Collective: every rank zeroes its own inbox (the step index in every word) between two barriers:
        the start of a step sequence."""
        torch.cuda.synchronize(self._device())
        dist.barrier(group=self.group)
        self._symm[0].zero_()
        torch.cuda.synchronize(self._device())
        dist.barrier(group=self.group)

    def check_exchange(self) -> None:
        """Raises if a thread of an earlier step of this rank gave up waiting for a peer (``dmvae_dp_status``;
        synchronises the stream).  Cheap enough for once per epoch; the step itself never synchronises."""
        if self.peers is None:
            return
        from ._lib import byref, check, stream_ptr
        with torch.cuda.device(self._device()):
            check(self.engine.lib.dmvae_dp_status(self.engine._cfg_ref, byref(self.peers), stream_ptr()), "dmvae_dp_status")

    def global_batch_layout(self, local_rows: int, equal: bool = True) -> Tuple[int, int]:
        """(global batch size, this rank's row offset).  With ``equal`` every rank holds
        ``local_rows`` (the usual case: no collective); otherwise sizes are all-gathered."""
        if self.world == 1:
            return local_rows, 0
        if equal:
            return local_rows * self.world, local_rows * self.rank
        dev = self.engine.grad_buf.device
        mine = torch.tensor([local_rows], dtype=torch.int64, device=dev)
        allv = [torch.zeros_like(mine) for _ in range(self.world)]
        dist.all_gather(allv, mine, group=self.group)
        sizes = [int(v.item()) for v in allv]
        return sum(sizes), sum(sizes[: self.rank])

    def step(self, local_batch: torch.Tensor, eps: Optional[torch.Tensor] = None, equal_shards: bool = True):
        B, offset = self.global_batch_layout(int(local_batch.shape[0]), equal_shards)
        if self.peers is not None:
            return self.engine.step_dp(local_batch, self.peers, B, eps=eps, sample_offset=offset)
        self.engine.loss_and_grads(local_batch, eps=eps, global_batch=B, sample_offset=offset)
        if self.world > 1:
            dist.all_reduce(self.engine.grad_buf, op=dist.ReduceOp.SUM, group=self.group)
        self.engine.apply()
        return self.engine.grad_buf[self.engine.n_params:]

    def capture(self, local_rows: int, host_batch=None, host_losses=None, dataset=None, shuffle=False, shuffle_seed=0):
        """The whole data-parallel step for ``local_rows`` rows per rank as one CUDA graph on every rank:
        fused forward+loss+backward, the SUM all-reduce of [gradients | losses] (NCCL, captured), Adam,
        repack (engine = ``FusedTrainer``; equal shards).  Returns the engine's ``GraphStep``."""
        B, offset = self.global_batch_layout(int(local_rows), True)
        if self.peers is not None:
            return self.engine.capture(int(local_rows), host_batch=host_batch, host_losses=host_losses, sample_offset=offset,
                                       global_batch=B, peers=self.peers, peers_reset=self._reset_flags, dataset=dataset,
                                       shuffle=shuffle, shuffle_seed=shuffle_seed)
        if dataset is not None and self.world == 1:
            return self.engine.capture(int(local_rows), host_losses=host_losses, dataset=dataset, shuffle=shuffle,
                                       shuffle_seed=shuffle_seed)
        if dataset is not None:
            raise ValueError("a resident dataset needs the peer exchange (exchange='peer') when there is more than one rank")
        reduce_fn = None
        if self.world > 1:
            def reduce_fn(t):
                dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        else:
            def reduce_fn(t):
                return None
        return self.engine.capture(int(local_rows), host_batch=host_batch, host_losses=host_losses, sample_offset=offset,
                                   all_reduce=reduce_fn, global_batch=B)

    def parameter_checksum(self, flat_params: torch.Tensor) -> bool:
        """True when every rank holds bit-identical parameters (compares the exact
        int64 sum of the raw fp32 bit patterns, a checksum of the replicas)."""
        s = flat_params.view(torch.int32).to(torch.int64).sum().reshape(1)
        if self.world == 1:
            return True
        lo, hi = s.clone(), s.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN, group=self.group)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX, group=self.group)
        return bool((lo == hi).item())


# ------------------------------------------------------------------------------------------
# sharded generation
# ------------------------------------------------------------------------------------------
def generate_shard(model, start_points, n: int, seed: int, rank: int, world: int, z: Optional[torch.Tensor] = None,
                   chunk: int = 1 << 22):
    """Decode this rank's rows [lo, hi) of a global job of ``n`` trajectories.
    ``start_points``: one shared (x, y) or a global (n, 2) array (sliced here); ``z``: an
    optional global (n, L) latent tensor (sliced) - by default latents come from the
    in-kernel Philox stream keyed by (seed, GLOBAL row index).  Returns (lo, hi, tensor)."""
    lo, hi = shard_range(n, rank, world)
    sp = torch.as_tensor(np.asarray(start_points, dtype=np.float64)).reshape(-1, 2)
    shared = sp.shape[0] == 1
    if not shared and sp.shape[0] != n:
        raise ValueError(f"start_points has {sp.shape[0]} rows, job has {n}")
    outs = []
    for c0 in range(lo, hi, chunk):
        c1 = min(hi, c0 + chunk)
        s = sp if shared else sp[c0:c1]
        zz = None if z is None else z[c0:c1]
        outs.append(model.generate(s.float(), z=zz, n=c1 - c0, seed=seed, sample_offset=c0))
    if not outs:
        dev = model.flat_parameters().device
        return lo, hi, torch.empty(0, model.seq_len, 3, dtype=torch.float32, device=dev)
    return lo, hi, (outs[0] if len(outs) == 1 else torch.cat(outs, 0))


def track_shard(waypoints, initial_states, dt: float, rank: int, world: int, **kw):
    """The MPC tracker (``dmvae.tracker.track_batch``) on this rank's rows [lo, hi) of a global batch of waypoint sets:
    trajectories are independent, so tracking shards like generation - contiguous row ranges, no collective, and every
    row's result is what a single GPU computes for it.  Returns (lo, hi, TrackResult or None for an empty shard)."""
    from .tracker import track_batch
    n = int(waypoints.shape[0])
    lo, hi = shard_range(n, rank, world)
    if hi == lo:
        return lo, hi, None
    total = kw.pop("total_time", None)
    if total is not None and np.ndim(total) > 0:
        total = np.asarray(total)[lo:hi]
    return lo, hi, track_batch(waypoints[lo:hi], initial_states[lo:hi], dt, total_time=total, **kw)


def write_sharded_npy(path: str, rows: np.ndarray, lo: int, n: int, rank: int, world: int) -> None:
    """All ranks write their slab [lo, lo+len(rows)) of one ``(n, T, 3)`` float32 ``.npy``:
    rank 0 creates the pre-sized file, a barrier, then every rank writes through a memmap."""
    shape = (n,) + tuple(rows.shape[1:])
    if rank == 0:
        os.makedirs(os.path.dirname(os.path.abspath(path)) or ".", exist_ok=True)
        mm = np.lib.format.open_memmap(path, mode="w+", dtype=np.float32, shape=shape)
        del mm
    if world > 1:
        dist.barrier()
    mm = np.lib.format.open_memmap(path, mode="r+")
    if mm.shape != shape:
        raise RuntimeError(f"{path}: shape {mm.shape}, expected {shape}")
    mm[lo:lo + rows.shape[0]] = rows
    mm.flush()
    del mm
    if world > 1:
        dist.barrier()


def generate_scenarios(model_paths: Sequence[str], scenario_names: Sequence[str], starts: Sequence[Tuple[float, float]],
                       n_per_scenario: int, out_dir: str = "results/GeneratedData", seq_len: int = 10,
                       latent_dim: int = 8, seed: int = 0):
    """Bulk generation for the defensive scenarios (BASELINE configs[2]): for each scenario,
    ``n_per_scenario`` decoded waypoint trajectories for its start point, sharded over the
    ranks of the job, written to ``<out_dir>/decoded_waypoints_<scenario>.npy`` as
    ``(n, T, 3)`` float32 ``[t, x, y]`` (an additive file; the reference's
    ``tracked_trajectory_*`` MPC outputs, Distribution.py:157, are not touched)."""
    from .model import ConditionalTrajectoryVAE
    rank, world, _ = init_distributed()
    paths = []
    for path, name, start in zip(model_paths, scenario_names, starts):
        model = ConditionalTrajectoryVAE(seq_len, 3, latent_dim)
        model.load_state_dict(torch.load(path, map_location="cpu"))
        model.eval()
        lo, hi, out = generate_shard(model, [start], n_per_scenario, seed, rank, world)
        dst = os.path.join(out_dir, f"decoded_waypoints_{name}.npy")
        write_sharded_npy(dst, out.cpu().numpy(), lo, n_per_scenario, rank, world)
        paths.append(dst)
    return paths
