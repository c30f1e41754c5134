"""Fused training step: the throughput path behind ``Training_VAE.py``'s loop.

One call = relative-offset transform + forward + five-term loss + backward
(``train_kernel``), fixed-order slab reduction + Adam (``reduce_kernel``) and the
refresh of the kernel-layout weights (``pack_kernel``): Training_VAE.py:345-363
without a single host synchronisation.  The five loss terms of the step stay on
the device in the tail of the gradient buffer; ``LossMeter`` accumulates them
there and is read back once per epoch (the reference does five ``.item()`` per
step, Training_VAE.py:366-370).
"""
from __future__ import annotations

import ctypes
from typing import Optional, Sequence

import torch

from . import _lib
from ._lib import DmvaeAdam, DmvaeLossWeights, byref, check, ptr, stream_ptr

LOSS_KEYS = ("total_loss", "recon_loss", "kld_loss", "start_loss", "time_loss")  # Training_VAE.py:337


class FusedTrainer:
    """Owns the Adam state (flat m, v, step), the gradient buffer and the kernel
    workspace for a ``ConditionalTrajectoryVAE``.

    ``step(batch)``              single-GPU fused step (4 launches)
    ``loss_and_grads(batch)``    fused forward+loss+backward only -> (losses, grads)
    ``apply(grads)``             Adam + repack from an (all-reduced) gradient buffer
    ``capture(B, ...)``          the whole step as one CUDA graph (``GraphStep``): no per-step host work
    """

    def __init__(self, model, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 weights: Sequence[float] = (0.1, 0.1, 1.0, 1.0), seed: int = 0):
        self.model = model
        self.lib = _lib.lib()
        self.cfg = model._cfg
        self.lr, self.betas, self.eps = float(lr), (float(betas[0]), float(betas[1])), float(eps)
        self.weights = DmvaeLossWeights(*[float(w) for w in weights])
        self.seed = int(seed)
        self.t = 0                                   # Adam step count (torch: state['step'])
        arena = model.flat_parameters()
        self.device = arena.device
        self.n_params = arena.numel()
        with torch.cuda.device(self.device):
            n = check(self.lib.dmvae_grad_count(byref(self.cfg)), "dmvae_grad_count")
        self.m = torch.zeros_like(arena)
        self.v = torch.zeros_like(arena)
        self.grad_buf = torch.zeros(n, dtype=torch.float32, device=self.device)   # grads + 5 loss terms
        self.step_dev = torch.zeros(1, dtype=torch.int64, device=self.device)     # device copy of ``t`` (GraphStep)
        self._dev_t = 0                                                            # what step_dev is known to hold
        self._ws = {}
        self._cfg_ref = byref(self.cfg)
        self._w_ref = byref(self.weights)

    # ------------------------------------------------------------------ helpers
    @property
    def grads(self) -> torch.Tensor:
        return self.grad_buf[: self.n_params]

    @property
    def losses(self) -> torch.Tensor:
        """[total, recon, kld, start, time] of the last pass (device tensor, no sync)."""
        return self.grad_buf[self.n_params:]

    def _workspace(self, B: int) -> torch.Tensor:
        ws = self._ws.get(B)
        if ws is None:
            with torch.cuda.device(self.device):
                nbytes = check(self.lib.dmvae_train_workspace_bytes(self._cfg_ref, B), "dmvae_train_workspace_bytes")
            ws = torch.zeros(nbytes, dtype=torch.uint8, device=self.device)   # zero once: self-resetting counters in its header
            self._ws = {B: ws}       # keep only the latest size
        return ws

    def _check_batch(self, batch: torch.Tensor) -> torch.Tensor:
        if batch.device != self.device or batch.dtype != torch.float32 or not batch.is_contiguous():
            batch = batch.to(device=self.device, dtype=torch.float32).contiguous()
        if batch.dim() != 3 or batch.shape[1] != self.model.seq_len or batch.shape[2] != 3 or batch.shape[0] == 0:
            raise ValueError(f"batch must be a non-empty (B, {self.model.seq_len}, 3) tensor, got {tuple(batch.shape)}")
        return batch

    def _check_eps(self, eps, B):
        if eps is None:
            return None
        eps = eps.to(device=self.device, dtype=torch.float32).contiguous()
        if tuple(eps.shape) != (B, self.model.latent_dim):
            raise ValueError(f"eps must be (B, {self.model.latent_dim})")
        return eps

    def sync_step_counter(self) -> None:
        """Device copy of the Adam step count := the host's (after host-driven steps)."""
        if self._dev_t != self.t:
            self.step_dev.fill_(self.t)
            self._dev_t = self.t

    def capture(self, B: int, host_batch: Optional[torch.Tensor] = None, host_losses: Optional[torch.Tensor] = None,
                sample_offset: int = 0, all_reduce=None, global_batch: Optional[int] = None, peers=None,
                peers_reset=None, dataset: Optional[torch.Tensor] = None, shuffle: bool = False,
                shuffle_seed: int = 0) -> "GraphStep":
        """The whole step for batch size ``B`` as one CUDA graph (host-driven ``step()`` / ``apply()`` calls may
        be mixed in: ``replay()`` re-synchronises the device-side step counter when needed)."""
        return GraphStep(self, B, host_batch, host_losses, sample_offset, all_reduce, global_batch, peers, peers_reset,
                         dataset, shuffle, shuffle_seed)

    # ------------------------------------------------------------------ passes
    def loss_and_grads(self, batch: torch.Tensor, eps: Optional[torch.Tensor] = None,
                       global_batch: Optional[int] = None, sample_offset: int = 0):
        """Fused forward + loss + backward.  ``global_batch`` is the size the means
        are taken over (data-parallel: sum of the ranks' batches).  Returns
        (losses[5], grads[n_params]) as views of the internal buffer."""
        batch = self._check_batch(batch)
        B = batch.shape[0]
        eps = self._check_eps(eps, B)
        packed = self.model.packed_weights()
        ws = self._workspace(B)
        inv = 1.0 / float(global_batch if global_batch is not None else B)
        with torch.cuda.device(self.device):
            check(self.lib.dmvae_train_fwd_bwd(self._cfg_ref, ptr(packed), ptr(batch), ptr(eps),
                                               ctypes.c_uint64(self.seed), ctypes.c_uint64(sample_offset),
                                               ctypes.c_uint64(self.t + 1), self._w_ref, ctypes.c_float(inv), B,
                                               ptr(ws), ptr(self.grad_buf), stream_ptr()), "dmvae_train_fwd_bwd")
        return self.losses, self.grads

    def _adam(self) -> DmvaeAdam:
        return DmvaeAdam(self.lr, self.betas[0], self.betas[1], self.eps, self.t)

    def apply(self, grads: Optional[torch.Tensor] = None) -> None:
        """optimizer.step() from ``grads`` (default: the internal buffer, e.g. after an
        in-place all-reduce) + refresh of the packed weights."""
        g = self.grad_buf if grads is None else grads
        arena = self.model.flat_parameters()
        packed = self.model.packed_weights()
        self.t += 1
        h = self._adam()
        with torch.cuda.device(self.device):
            check(self.lib.dmvae_adam_step(self._cfg_ref, ptr(arena), ptr(g), ptr(self.m), ptr(self.v), byref(h),
                                           ptr(packed), stream_ptr()), "dmvae_adam_step")
        self.model.mark_packed_current()

    def step(self, batch: torch.Tensor, eps: Optional[torch.Tensor] = None, sample_offset: int = 0) -> torch.Tensor:
        """One full single-GPU training step; returns the 5 loss terms (device view)."""
        batch = self._check_batch(batch)
        B = batch.shape[0]
        eps = self._check_eps(eps, B)
        arena = self.model.flat_parameters()
        packed = self.model.packed_weights()
        ws = self._workspace(B)
        self.t += 1
        h = self._adam()
        with torch.cuda.device(self.device):
            check(self.lib.dmvae_train_step(self._cfg_ref, ptr(arena), ptr(packed), ptr(self.m), ptr(self.v),
                                            ptr(batch), ptr(eps), ctypes.c_uint64(self.seed),
                                            ctypes.c_uint64(sample_offset), self._w_ref, ctypes.c_float(1.0 / B), B,
                                            byref(h), ptr(ws), ptr(self.grad_buf), stream_ptr()), "dmvae_train_step")
        self.model.mark_packed_current()
        return self.losses


    def step_dp(self, batch: torch.Tensor, peers, global_batch: int, eps: Optional[torch.Tensor] = None,
                sample_offset: int = 0) -> torch.Tensor:
        """One data-parallel step with the gradient exchange inside the update kernel (``dmvae_train_step_dp``:
        peer-memory reads over NVLink, no library collective).  ``peers``: a ``DmvaeDpPeers`` naming every rank's
        exchange / flag buffer; all ranks call with the same step.  Returns the loss terms of the GLOBAL batch."""
        batch = self._check_batch(batch)
        B = batch.shape[0]
        eps = self._check_eps(eps, B)
        arena = self.model.flat_parameters()
        packed = self.model.packed_weights()
        ws = self._workspace(B)
        self.t += 1
        h = self._adam()
        with torch.cuda.device(self.device):
            check(self.lib.dmvae_train_step_dp(self._cfg_ref, ptr(arena), ptr(packed), ptr(self.m), ptr(self.v),
                                               ptr(batch), ptr(eps), ctypes.c_uint64(self.seed),
                                               ctypes.c_uint64(sample_offset), self._w_ref,
                                               ctypes.c_float(1.0 / float(global_batch)), B, byref(h), None, ptr(ws),
                                               ptr(self.grad_buf), byref(peers), stream_ptr()), "dmvae_train_step_dp")
        self.model.mark_packed_current()
        return self.losses


class GraphStep:
    """One training step captured as a CUDA graph: optional H2D copy of the batch from a pinned host
    buffer, the fused step (``dmvae_train_step_dev``: chain, weight gradients, reduction + Adam, repack;
    the Adam step index lives in device memory so that nothing in the graph depends on the step), optional
    D2H copy of the five loss terms into a pinned host buffer.  ``replay()`` is one graph launch."""

    def __init__(self, trainer: "FusedTrainer", B: int, host_batch: Optional[torch.Tensor] = None,
                 host_losses: Optional[torch.Tensor] = None, sample_offset: int = 0,
                 all_reduce=None, global_batch: Optional[int] = None, peers=None, peers_reset=None,
                 dataset: Optional[torch.Tensor] = None, shuffle: bool = False, shuffle_seed: int = 0):
        """``all_reduce``: data-parallel ranks pass a callable that SUM-all-reduces a tensor in place (captured
        in the graph between the fused pass and the Adam update) and ``global_batch`` / ``sample_offset`` =
        the global batch size and this rank's row offset.  ``peers`` (a ``DmvaeDpPeers``) instead selects the
        step whose update kernel exchanges the gradients over peer memory itself (``dmvae_train_step_dp``).
        ``dataset``: a device tensor ``(n_batches * B, T, 3)`` that stays resident; update t then reads batch
        ``(t - 1) mod n_batches`` of it, selected in the kernel from the device-side step counter
        (``dmvae_train_step_resident``): replaying the graph walks the set with no per-step copy (``batch`` is unused).
        ``shuffle``: the rows of every epoch are picked through a new permutation keyed by ``(shuffle_seed, epoch)``, like
        the reference's ``DataLoader(shuffle=True)`` (``resident_order`` returns the same permutation on the host)."""
        model = trainer.model
        self.trainer = trainer
        self.B = int(B)
        dev = trainer.device
        if peers is not None and peers_reset is None:
            # the warm-up launch below publishes words tagged with step t + 1 in the peers' inboxes and the counter
            # is then rewound: without a collective reset the first replay could match those stale words
            raise ValueError("peers= needs peers_reset= (a collective that zeroes every rank's inbox between two "
                             "barriers; DataParallelTrainer.capture passes its own)")
        self.batch = torch.zeros(B, model.seq_len, 3, dtype=torch.float32, device=dev)   # static input buffer
        self.host_batch, self.host_losses = host_batch, host_losses
        if host_batch is not None and (not host_batch.is_pinned() or tuple(host_batch.shape) != tuple(self.batch.shape)):
            raise ValueError("host_batch must be a pinned (B, T, 3) fp32 tensor")
        arena = model.flat_parameters()
        packed = model.packed_weights()
        ws = trainer._workspace(B)
        self._keep = (arena, packed, ws)
        hyper = DmvaeAdam(trainer.lr, trainer.betas[0], trainer.betas[1], trainer.eps, 0)
        self._hyper = hyper
        lib = trainer.lib
        inv = 1.0 / float(global_batch if global_batch is not None else B)

        if dataset is not None:
            if all_reduce is not None or host_batch is not None:
                raise ValueError("dataset= excludes host_batch= and all_reduce= (use peers= for data parallel)")
            if (dataset.device != dev or dataset.dtype != torch.float32 or not dataset.is_contiguous() or dataset.dim() != 3
                    or dataset.shape[1] != model.seq_len or dataset.shape[2] != 3 or dataset.shape[0] < B):
                raise ValueError(f"dataset must be a contiguous fp32 ({'>='}{B}, {model.seq_len}, 3) tensor on {dev}")
            self.n_batches = int(dataset.shape[0]) // int(B)
            self._keep = self._keep + (dataset,)

        def launch():
            if host_batch is not None:
                self.batch.copy_(host_batch, non_blocking=True)
            if dataset is not None:
                check(lib.dmvae_train_step_resident(trainer._cfg_ref, ptr(arena), ptr(packed), ptr(trainer.m), ptr(trainer.v),
                                                    ptr(dataset), self.n_batches, int(bool(shuffle)),
                                                    ctypes.c_uint64(int(shuffle_seed)), ctypes.c_uint64(trainer.seed),
                                                    ctypes.c_uint64(sample_offset), trainer._w_ref, ctypes.c_float(inv), B,
                                                    byref(hyper), ptr(trainer.step_dev), ptr(ws), ptr(trainer.grad_buf),
                                                    byref(peers) if peers is not None else None, stream_ptr()),
                      "dmvae_train_step_resident")
            elif peers is not None:
                check(lib.dmvae_train_step_dp(trainer._cfg_ref, ptr(arena), ptr(packed), ptr(trainer.m), ptr(trainer.v),
                                              ptr(self.batch), None, ctypes.c_uint64(trainer.seed),
                                              ctypes.c_uint64(sample_offset), trainer._w_ref, ctypes.c_float(inv), B,
                                              byref(hyper), ptr(trainer.step_dev), ptr(ws), ptr(trainer.grad_buf),
                                              byref(peers), stream_ptr()), "dmvae_train_step_dp")
            elif all_reduce is None:
                check(lib.dmvae_train_step_dev(trainer._cfg_ref, ptr(arena), ptr(packed), ptr(trainer.m), ptr(trainer.v),
                                               ptr(self.batch), None, ctypes.c_uint64(trainer.seed),
                                               ctypes.c_uint64(sample_offset), trainer._w_ref, ctypes.c_float(inv), B,
                                               byref(hyper), ptr(trainer.step_dev), ptr(ws), ptr(trainer.grad_buf),
                                               stream_ptr()), "dmvae_train_step_dev")
            else:
                check(lib.dmvae_train_fwd_bwd_dev(trainer._cfg_ref, ptr(packed), ptr(self.batch), None,
                                                  ctypes.c_uint64(trainer.seed), ctypes.c_uint64(sample_offset),
                                                  ptr(trainer.step_dev), trainer._w_ref, ctypes.c_float(inv), B, ptr(ws),
                                                  ptr(trainer.grad_buf), stream_ptr()), "dmvae_train_fwd_bwd_dev")
                all_reduce(trainer.grad_buf)
                check(lib.dmvae_adam_step_dev(trainer._cfg_ref, ptr(arena), ptr(trainer.grad_buf), ptr(trainer.m),
                                              ptr(trainer.v), byref(hyper), ptr(trainer.step_dev), ptr(packed),
                                              stream_ptr()), "dmvae_adam_step_dev")
            if host_losses is not None:
                host_losses.copy_(trainer.losses, non_blocking=True)

        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.device(dev):
            trainer.sync_step_counter()
            torch.cuda.current_stream().synchronize()
            saved = (arena.clone(), trainer.m.clone(), trainer.v.clone(), trainer.step_dev.clone())
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):        # warm-up launch outside the capture (lazy module loading, attributes)
                launch()
            torch.cuda.current_stream().wait_stream(s)
            n0 = lib.dmvae_launch_count(-1)
            with torch.cuda.graph(self.graph):
                launch()
            self.kernels = int(lib.dmvae_launch_count(-1) - n0)   # libdmvae kernels launched by one replay
            # the warm-up applied one update: restore the state it changed
            arena.copy_(saved[0]); trainer.m.copy_(saved[1]); trainer.v.copy_(saved[2]); trainer.step_dev.copy_(saved[3])
            check(lib.dmvae_pack_weights(trainer._cfg_ref, ptr(arena), ptr(packed), stream_ptr()), "dmvae_pack_weights")
            model.mark_packed_current()
            if peers is not None:
                # the warm-up published its step index in the peers' inboxes and the counter was just rewound:
                # the inboxes start over (collective: every rank resets its own between two barriers).  The same
                # reset is needed whenever the step counter moves backwards (e.g. resuming an earlier checkpoint
                # with the same symmetric buffer): a word of the old run with the same step index would match.
                peers_reset()

    def replay(self) -> torch.Tensor:
        """Runs the captured step on the current contents of ``batch`` (or ``host_batch``); returns the five
        loss terms (device view, no sync)."""
        tr = self.trainer
        tr.sync_step_counter()
        self.graph.replay()
        tr.t += 1
        tr._dev_t = tr.t
        tr.model.mark_packed_current()
        return tr.losses


def resident_order(n_rows: int, epoch: int, shuffle_seed: int = 0):
    """The row order of one epoch of a shuffled resident set, as a list: position p reads row ``order[p]``
    (``dmvae_resident_row``, the function the kernel evaluates per row)."""
    lib = _lib.lib()
    return [check(lib.dmvae_resident_row(ctypes.c_uint64(int(shuffle_seed)), int(epoch), p, int(n_rows)), "dmvae_resident_row")
            for p in range(int(n_rows))]


class LossMeter:
    """Sample-weighted running sums of the five loss terms, kept on the device
    (Training_VAE.py:339, :366-380 without the per-step host syncs)."""

    def __init__(self, device):
        self.sums = torch.zeros(5, dtype=torch.float64, device=device)
        self.count = 0

    def update(self, losses: torch.Tensor, batch_size: int) -> None:
        self.sums += losses.double() * batch_size
        self.count += batch_size

    def means(self):
        """One device->host read: the epoch means in LOSS_KEYS order."""
        vals = (self.sums / max(self.count, 1)).cpu().tolist()
        self.sums.zero_()
        self.count = 0
        return vals
