"""autograd bridges: ``model(x, start)`` and ``conditional_vae_loss(...)`` keep
working with ``loss.backward()`` / ``optimizer.step()`` exactly as the reference's
training loop writes them (Training_VAE.py:352-363), while forward, loss and
backward each run as one fused kernel behind the C ABI
(``dmvae_forward`` / ``dmvae_loss`` / ``dmvae_loss_backward`` / ``dmvae_backward``).

The throughput path is ``dmvae.train.FusedTrainer`` (one fused kernel for all of
it); this module is the drop-in surface.
"""
from __future__ import annotations

import torch

from . import _lib
from ._lib import DmvaeLossWeights, byref, check, ptr, stream_ptr

HIDDEN = 128


def _param_views(flat: torch.Tensor, params):
    out, off = [], 0
    for p in params:
        n = p.numel()
        out.append(flat[off:off + n].view(p.shape))
        off += n
    return out


class _VaeForward(torch.autograd.Function):
    """(x_rel, start, eps, *params) -> (recon, mu, logvar, h_c)."""

    @staticmethod
    def forward(ctx, model, x_rel, start, eps, *params):
        lib = _lib.lib()
        cfg = model._cfg
        B = x_rel.shape[0]
        dev = x_rel.device
        packed = model.packed_weights()
        T, L = model.seq_len, model.latent_dim
        recon = torch.empty(B, T, 3, dtype=torch.float32, device=dev)
        mu = torch.empty(B, L, dtype=torch.float32, device=dev)
        logvar = torch.empty(B, L, dtype=torch.float32, device=dev)
        h_c = torch.empty(B, HIDDEN, dtype=torch.float32, device=dev)
        ctx.model, ctx.B = model, B
        if B == 0:
            ctx.stash = None
            return recon, mu, logvar, h_c
        with torch.cuda.device(dev):
            nbytes = check(lib.dmvae_stash_bytes(byref(cfg), B), "dmvae_stash_bytes")
            stash = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            check(lib.dmvae_forward(byref(cfg), ptr(packed), ptr(x_rel), ptr(start), ptr(eps), ptr(recon), ptr(mu),
                                    ptr(logvar), ptr(h_c), ptr(stash), B, stream_ptr()), "dmvae_forward")
        ctx.stash = stash
        # the backward kernel reads the weights as they were at forward time
        ctx.packed = packed.clone() if any(p.requires_grad for p in params) else None
        return recon, mu, logvar, h_c

    @staticmethod
    def backward(ctx, g_recon, g_mu, g_logvar, g_hc):
        model, B = ctx.model, ctx.B
        params = list(model.parameters())
        if B == 0 or ctx.stash is None:
            return (None, None, None, None) + tuple(torch.zeros_like(p) for p in params)
        lib = _lib.lib()
        cfg = model._cfg
        dev = ctx.stash.device

        def prep(g):
            return None if g is None else g.detach().to(device=dev, dtype=torch.float32).contiguous()

        g_recon, g_mu, g_logvar, g_hc = prep(g_recon), prep(g_mu), prep(g_logvar), prep(g_hc)
        with torch.cuda.device(dev):
            n = check(lib.dmvae_grad_count(byref(cfg)), "dmvae_grad_count")
            grads = torch.empty(n, dtype=torch.float32, device=dev)
            wbytes = check(lib.dmvae_train_workspace_bytes(byref(cfg), B), "dmvae_train_workspace_bytes")
            ws = torch.empty(wbytes, dtype=torch.uint8, device=dev)
            check(lib.dmvae_backward(byref(cfg), ptr(ctx.packed), ptr(g_recon), ptr(g_mu), ptr(g_logvar), ptr(g_hc),
                                     ptr(ctx.stash), ptr(ws), ptr(grads), B, stream_ptr()), "dmvae_backward")
        return (None, None, None, None) + tuple(_param_views(grads, params))


def vae_forward(model, x, start_points, eps=None, need_recon=True):
    """model.forward / model.encode.  Inputs may live on any device; results come
    back on the device of ``x``."""
    arena = model.flat_parameters()
    dev = arena.device
    src = x.device if torch.is_tensor(x) else torch.device("cpu")
    x_rel = torch.as_tensor(x).detach().to(device=dev, dtype=torch.float32).contiguous()
    if x_rel.dim() != 3 or tuple(x_rel.shape[1:]) != (model.seq_len, 3):
        raise ValueError(f"x must be (B, {model.seq_len}, 3), got {tuple(x_rel.shape)}")
    B = x_rel.shape[0]
    start = torch.as_tensor(start_points).detach().to(device=dev, dtype=torch.float32).contiguous()
    if tuple(start.shape) != (B, 2):
        raise ValueError(f"start_points must be (B, 2), got {tuple(start.shape)}")
    if eps is None:
        if need_recon:
            # drawn where the reference draws it: randn_like(std) on the caller's device
            # (Training_VAE.py:205) -> same generator, same stream position
            eps = torch.randn(B, model.latent_dim, dtype=torch.float32, device=src)
        else:
            eps = torch.zeros(B, model.latent_dim, dtype=torch.float32, device=dev)
    eps = torch.as_tensor(eps).detach().to(device=dev, dtype=torch.float32).contiguous()
    if tuple(eps.shape) != (B, model.latent_dim):
        raise ValueError(f"eps must be (B, {model.latent_dim}), got {tuple(eps.shape)}")
    outs = _VaeForward.apply(model, x_rel, start, eps, *model.parameters())
    if src != dev:
        outs = tuple(o.to(src) for o in outs)
    return outs


class _Loss(torch.autograd.Function):
    """(recon, x, mu, logvar) -> (5,) tensor [total, recon, kld, start, time]."""

    @staticmethod
    def forward(ctx, recon, x, mu, logvar, weights):
        lib = _lib.lib()
        B, T = recon.shape[0], recon.shape[1]
        L = mu.shape[1]
        cfg = _lib.cfg(T, L)
        w = DmvaeLossWeights(*[float(v) for v in weights])
        losses = torch.empty(5, dtype=torch.float32, device=recon.device)
        with torch.cuda.device(recon.device):
            check(lib.dmvae_loss(byref(cfg), ptr(recon), ptr(x), ptr(mu), ptr(logvar), byref(w), B, ptr(losses),
                                 stream_ptr()), "dmvae_loss")
        ctx.save_for_backward(recon, x, mu, logvar)
        ctx.weights, ctx.cfg = w, cfg
        return losses

    @staticmethod
    def backward(ctx, g_out):
        recon, x, mu, logvar = ctx.saved_tensors
        lib = _lib.lib()
        g_out = g_out.detach().to(dtype=torch.float32).contiguous()
        g_recon, g_mu, g_logvar = torch.empty_like(recon), torch.empty_like(mu), torch.empty_like(logvar)
        with torch.cuda.device(recon.device):
            check(lib.dmvae_loss_backward(byref(ctx.cfg), ptr(recon), ptr(x), ptr(mu), ptr(logvar), byref(ctx.weights),
                                          recon.shape[0], ptr(g_out), ptr(g_recon), ptr(g_mu), ptr(g_logvar),
                                          stream_ptr()), "dmvae_loss_backward")
        return g_recon, None, g_mu, g_logvar, None


def conditional_vae_loss(recon_x, x, mu, logvar, condition, recon_weight=0.1, kld_weight=0.1, start_weight=1.0,
                         time_weight=0.5):
    """Drop-in for Training_VAE.conditional_vae_loss (Training_VAE.py:229-268):
    returns (total, recon, kld, start, time); as in the reference a term whose
    weight is <= 0 is the python int 0 and ``condition`` is unused."""
    if not torch.cuda.is_available():
        raise _lib.DmvaeError("no CUDA device is visible and dmvae has no CPU path")
    src = recon_x.device
    dev = src if src.type == "cuda" else torch.device("cuda", torch.cuda.current_device())

    def prep(t):
        return t.to(device=dev, dtype=torch.float32).contiguous()

    r, xx, m, lv = prep(recon_x), prep(x).detach(), prep(mu), prep(logvar)
    if r.dim() != 3 or r.shape[2] != 3 or r.shape != xx.shape or m.shape != lv.shape or m.shape[0] != r.shape[0]:
        raise ValueError("conditional_vae_loss: expected recon_x/x (B,T,3) and mu/logvar (B,L)")
    if r.shape[0] == 0:
        raise ValueError("conditional_vae_loss: empty batch (the reference's means would be NaN)")
    out = _Loss.apply(r, xx, m, lv, (recon_weight, kld_weight, start_weight, time_weight))
    if src != dev:
        out = out.to(src)
    total, recon_loss, kld, start_loss, time_loss = out.unbind(0)
    if not start_weight > 0:
        start_loss = 0
    if not time_weight > 0:
        time_loss = 0
    return total, recon_loss, kld, start_loss, time_loss
