"""ConditionalTrajectoryVAE with the reference's module surface on B200 kernels.

Mirrors ``Training_VAE.py:118-226`` of the reference: same constructor, same
attributes, same sub-module names (hence the same 24 ``state_dict`` keys in the
same order and the same seeded initialisation), same methods.  The arithmetic
is NOT torch's: every method hands device pointers to ``libdmvae.so``.

Storage: the parameters are ordinary ``nn.Parameter`` objects whose ``.data``
are views into one flat fp32 CUDA arena in ``state_dict`` order, so that
``state_dict()`` / ``load_state_dict()`` / ``torch.save`` round-trip natively
while the kernels see one contiguous buffer.  The kernel-layout ("packed") copy
of the weights is refreshed lazily whenever the arena's version counter moved.

Devices: there is no CPU path.  ``.to('cpu')`` and ``device='cpu'`` arguments
are accepted (the reference hard-codes them, ``Training_VAE.py:282``), the
parameters simply migrate to the current CUDA device at the first compute call;
results are returned on the device of the caller's input tensors.
"""
from __future__ import annotations

import ctypes
import warnings
from typing import Optional, Tuple

import torch
import torch.nn as nn

from . import _lib
from ._lib import byref, check, ptr, stream_ptr

HIDDEN = 128
MAX_LATENT = 64
MAX_SEQ = 400   # seq_len; beyond 3 * seq_len = 128 the first / last layers run in chunks on the FFMA kernels


def _envelope(seq_len: int, dim: int, latent_dim: int, hidden_dim: int) -> None:
    if dim != 3:
        raise NotImplementedError(
            f"dim={dim}: the loss, the offset transform and the start-offset add hard-code columns [t, x, y] "
            "(Training_VAE.py:250-261, :345-348); only dim=3 is supported")
    if hidden_dim != HIDDEN:
        raise NotImplementedError(f"hidden_dim={hidden_dim}: kernels are specialised for hidden_dim=128")
    if not (1 <= latent_dim <= MAX_LATENT):
        raise NotImplementedError(f"latent_dim={latent_dim}: supported range is 1..{MAX_LATENT}")
    if seq_len < 2 or seq_len > MAX_SEQ:
        raise NotImplementedError(f"seq_len={seq_len}: supported range is 2..{MAX_SEQ}")


def _dense(owner, linear: nn.Linear, x: torch.Tensor, relu: bool) -> torch.Tensor:
    """One ``nn.Linear`` (+ ReLU) of a sub-module called on its own: ``dmvae_dense`` on the layer's views of the
    parameter arena.  Forward only."""
    owner.flat_parameters()                      # parameters on the device, contiguous views of the arena
    w, b = linear.weight.data, (linear.bias.data if linear.bias is not None else None)
    xx = x.detach().to(device=w.device, dtype=torch.float32).contiguous()
    if xx.dim() != 2 or xx.shape[1] != linear.in_features:
        raise ValueError(f"expected (B, {linear.in_features}), got {tuple(x.shape)}")
    y = torch.empty(xx.shape[0], linear.out_features, dtype=torch.float32, device=w.device)
    if xx.shape[0] > 0:
        with torch.cuda.device(w.device):
            check(_lib.lib().dmvae_dense(ptr(w), ptr(b), ptr(xx), ptr(y), xx.shape[0], linear.in_features, linear.out_features,
                                         int(relu), stream_ptr()), "dmvae_dense")
    return y


class _KernelBackedSequential(nn.Sequential):
    """Keeps the reference's sub-module tree (state_dict keys, initialisation order).  Called on its own
    (``model.encoder(x)``, ``model.decoder(zc)`` - Training_VAE.py:141-151, :158-167; no caller in the reference does)
    it walks its layers through ``dmvae_dense``, one launch per Linear with the ReLU that follows it fused; forward
    only, no autograd graph.  ``model.encode`` / ``decode`` / ``forward`` never come here: they run in the fused kernels."""

    def forward(self, x):  # type: ignore[override]
        owner = self._owner()
        src = x.device
        layers = list(self.children())
        h = x
        i = 0
        while i < len(layers):
            layer = layers[i]
            if isinstance(layer, nn.Flatten):
                h = h.reshape(h.shape[0], -1)
            elif isinstance(layer, nn.Unflatten):
                h = h.reshape(h.shape[0], *layer.unflattened_size)
            elif isinstance(layer, nn.Linear):
                relu = i + 1 < len(layers) and isinstance(layers[i + 1], nn.ReLU)
                h = _dense(owner, layer, h, relu)
                i += int(relu)
            elif isinstance(layer, nn.ReLU):
                h = torch.clamp_min(h, 0.0)
            i += 1
        return h.to(src)


class _ConditionEncoder(_KernelBackedSequential):
    """``model.condition_encoder(c)`` is called directly by the reference's
    generate helpers (Tools.py:55, :898): the fused condition-encoder kernel."""

    def forward(self, condition: torch.Tensor) -> torch.Tensor:  # type: ignore[override]
        return self._owner()._cond_encode(condition)


class _KernelBackedLinear(nn.Linear):
    """``model.fc_mu(h)`` / ``model.fc_logvar(h)`` on their own (Training_VAE.py:154-155): ``dmvae_dense``."""

    def forward(self, x):  # type: ignore[override]
        return _dense(self._owner(), self, x, False).to(x.device)


class ConditionalTrajectoryVAE(nn.Module):
    """Drop-in for ``Training_VAE.ConditionalTrajectoryVAE`` (Training_VAE.py:118)."""

    def __init__(self, seq_len, dim, latent_dim, hidden_dim=128):
        super().__init__()
        _envelope(seq_len, dim, latent_dim, hidden_dim)
        self.seq_len = seq_len
        self.dim = dim
        self.latent_dim = latent_dim
        self.hidden_dim = hidden_dim

        # construction order == the reference's (Training_VAE.py:132-167) so that a
        # given torch.manual_seed yields bit-identical initial weights
        self.condition_encoder = _ConditionEncoder(
            nn.Linear(2, hidden_dim), nn.ReLU(), nn.Linear(hidden_dim, hidden_dim), nn.ReLU())
        self.encoder = _KernelBackedSequential(
            nn.Flatten(), nn.Linear(seq_len * dim, hidden_dim), nn.ReLU(), nn.Linear(hidden_dim, hidden_dim),
            nn.ReLU(), nn.Linear(hidden_dim, hidden_dim), nn.ReLU(), nn.Linear(hidden_dim, hidden_dim), nn.ReLU())
        self.fc_mu = _KernelBackedLinear(hidden_dim + hidden_dim, latent_dim)
        self.fc_logvar = _KernelBackedLinear(hidden_dim + hidden_dim, latent_dim)
        self.decoder = _KernelBackedSequential(
            nn.Linear(latent_dim + hidden_dim, hidden_dim), nn.ReLU(), nn.Linear(hidden_dim, hidden_dim), nn.ReLU(),
            nn.Linear(hidden_dim, hidden_dim), nn.ReLU(), nn.Linear(hidden_dim, seq_len * dim),
            nn.Unflatten(1, (seq_len, dim)))
        import weakref
        ref = weakref.ref(self)
        for sub in (self.condition_encoder, self.encoder, self.fc_mu, self.fc_logvar, self.decoder):
            object.__setattr__(sub, "_owner", ref)

        self._cfg = _lib.cfg(seq_len, latent_dim, dim, hidden_dim)
        self._arena: Optional[torch.Tensor] = None       # flat parameters, state_dict order
        self._packed: Optional[torch.Tensor] = None      # kernel layout
        self._packed_version = -1
        self._n_params = sum(p.numel() for p in self.parameters())

    # ------------------------------------------------------------------ storage
    @property
    def n_params(self) -> int:
        return self._n_params

    def _target_device(self) -> torch.device:
        if not torch.cuda.is_available():
            raise _lib.DmvaeError("no CUDA device is visible and dmvae has no CPU path")
        return torch.device("cuda", torch.cuda.current_device())

    def flat_parameters(self) -> torch.Tensor:
        """The flat fp32 CUDA arena (state_dict order) the parameters are views of;
        (re)built if ``.to()`` / ``load_state_dict(assign=True)`` broke the views."""
        params = list(self.parameters())
        a = self._arena
        if a is not None:
            base, off, ok = a.data_ptr(), 0, True
            for p in params:
                if p.data_ptr() != base + 4 * off or p.dtype != torch.float32 or not p.is_contiguous():
                    ok = False
                    break
                off += p.numel()
            if ok:
                return a
        dev = params[0].device if params[0].is_cuda else self._target_device()
        arena = torch.empty(self._n_params, dtype=torch.float32, device=dev)
        off = 0
        with torch.no_grad():
            for p in params:
                n = p.numel()
                view = arena[off:off + n].view(p.shape)
                view.copy_(p.detach().to(dtype=torch.float32))
                p.data = view
                off += n
        self._arena = arena
        self._packed_version = -1
        return arena

    def packed_weights(self) -> torch.Tensor:
        """Kernel-layout weights, repacked when the arena has been written to."""
        arena = self.flat_parameters()
        if self._packed is None or self._packed.device != arena.device:
            n = check(_lib.lib().dmvae_packed_count(byref(self._cfg)), "dmvae_packed_count")
            self._packed = torch.zeros(n, dtype=torch.float32, device=arena.device)
            self._packed_version = -1
        version = self._weights_version()
        if version != self._packed_version:
            with torch.cuda.device(arena.device):
                check(_lib.lib().dmvae_pack_weights(byref(self._cfg), ptr(arena), ptr(self._packed), stream_ptr()),
                      "dmvae_pack_weights")
            self._packed_version = version
        return self._packed

    def _weights_version(self) -> int:
        """Changes whenever torch writes a parameter in place (optimizer.step(),
        load_state_dict, p.data.copy_ ...): the parameters keep their own version
        counters after ``p.data = view``, so the arena's alone is not enough."""
        v = self._arena._version if self._arena is not None else 0
        for p in self.parameters():
            v += p._version
        return v

    def mark_packed_current(self) -> None:
        """Called by the fused optimiser, which refreshes the packed copy itself."""
        self._packed_version = self._weights_version()

    def _in(self, t, shape_tail) -> Tuple[torch.Tensor, torch.device]:
        """Bring a caller tensor to the arena's device as contiguous fp32."""
        if not torch.is_tensor(t):
            t = torch.as_tensor(t)
        src = t.device
        dev = self.flat_parameters().device
        t = t.detach().to(device=dev, dtype=torch.float32).contiguous()
        if tuple(t.shape[1:]) != tuple(shape_tail):
            raise ValueError(f"expected shape (B, {', '.join(map(str, shape_tail))}), got {tuple(t.shape)}")
        return t, src

    # ------------------------------------------------------------------ reference API
    def get_start_points(self, x):
        """Training_VAE.py:169-178."""
        return x[:, 0, 1:3]

    def _cond_encode(self, condition: torch.Tensor) -> torch.Tensor:
        c, src = self._in(condition, (2,))
        packed = self.packed_weights()
        h = torch.empty(c.shape[0], HIDDEN, dtype=torch.float32, device=c.device)
        with torch.cuda.device(c.device):
            check(_lib.lib().dmvae_cond_encode(byref(self._cfg), ptr(packed), ptr(c), ptr(h), c.shape[0], stream_ptr()),
                  "dmvae_cond_encode")
        return h.to(src)

    def decode(self, z, condition):
        """Training_VAE.py:208-215: ``condition`` is h_c = condition_encoder(start).
        Inference only (no autograd graph); training goes through forward()."""
        zz, src = self._in(z, (self.latent_dim,))
        hc, _ = self._in(condition, (HIDDEN,))
        if hc.shape[0] != zz.shape[0]:
            raise ValueError("z and condition batch sizes differ")
        packed = self.packed_weights()
        out = torch.empty(zz.shape[0], self.seq_len, 3, dtype=torch.float32, device=zz.device)
        with torch.cuda.device(zz.device):
            check(_lib.lib().dmvae_decode_from_condition(byref(self._cfg), ptr(packed), ptr(zz), ptr(hc), ptr(out),
                                                         zz.shape[0], stream_ptr()), "dmvae_decode_from_condition")
        return out.to(src)

    def reparameterize(self, mu, logvar):
        """Training_VAE.py:199-206 (elementwise; the fused train kernel does this
        in registers - this method exists for API compatibility)."""
        std = torch.exp(0.5 * logvar)
        eps = torch.randn_like(std)
        return mu + eps * std

    def encode(self, x, start_points):
        """Training_VAE.py:180-197 -> (mu, logvar, h_condition)."""
        from .autograd import vae_forward
        _, mu, logvar, h_c = vae_forward(self, x, start_points, eps=None, need_recon=False)
        return mu, logvar, h_c

    def forward(self, x, start_points, eps=None):
        """Training_VAE.py:217-226 -> (recon_x, mu, logvar, condition).  ``eps``
        (B, L) may be injected; by default it is drawn exactly where the reference
        draws it (``torch.randn_like`` on the caller's device/generator,
        Training_VAE.py:205), so that a seeded run consumes the same RNG stream."""
        from .autograd import vae_forward
        return vae_forward(self, x, start_points, eps=eps, need_recon=True)

    # ------------------------------------------------------------------ fused generation
    @torch.no_grad()
    def generate(self, start_points, z: Optional[torch.Tensor] = None, n: Optional[int] = None, seed: int = 0,
                 sample_offset: int = 0, add_start: bool = True, return_z: bool = False, out: Optional[torch.Tensor] = None):
        """Batched sample-and-decode (the loop of Tools.py:44-63 / :898-912 as one
        kernel).  ``start_points``: (B,2) per-row or (1,2)/(2,) shared.  ``z``: (B,L)
        fixed latents, or None to draw them in-kernel (Philox4x32-10, key ``seed``,
        counter ``sample_offset + row`` - independent of how rows are sharded).
        Returns (B, T, 3) fp32 [t, x, y] on the arena's device."""
        dev = self.flat_parameters().device
        sp = torch.as_tensor(start_points, dtype=torch.float64 if not torch.is_tensor(start_points) else None)
        sp = sp.detach().to(device=dev, dtype=torch.float32).reshape(-1, 2).contiguous()
        if z is not None:
            zz, _ = self._in(z, (self.latent_dim,))
            B = zz.shape[0]
        else:
            zz = None
            B = int(n) if n is not None else sp.shape[0]
        shared = sp.shape[0] == 1 and B > 1
        if not shared and sp.shape[0] != B:
            raise ValueError(f"start_points has {sp.shape[0]} rows, batch is {B}")
        packed = self.packed_weights()
        if out is None:
            out = torch.empty(B, self.seq_len, 3, dtype=torch.float32, device=dev)
        elif tuple(out.shape) != (B, self.seq_len, 3) or out.dtype != torch.float32 or not out.is_contiguous():
            raise ValueError("out must be a contiguous fp32 (B, T, 3) tensor")
        z_out = torch.empty(B, self.latent_dim, dtype=torch.float32, device=dev) if (return_z and zz is None) else None
        with torch.cuda.device(dev):
            check(_lib.lib().dmvae_decode(byref(self._cfg), ptr(packed), ptr(zz), ctypes.c_uint64(seed),
                                          ctypes.c_uint64(sample_offset), ptr(sp), int(shared), ptr(out), ptr(z_out),
                                          B, int(add_start), stream_ptr()), "dmvae_decode")
        if return_z:
            return out, (zz if zz is not None else z_out)
        return out
