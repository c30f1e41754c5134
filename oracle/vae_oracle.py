"""CPU oracle: restatement of the reference's trajectory-VAE hot path.

TEST INFRASTRUCTURE - not part of the product path.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import
this module.  The product (``dmvae``) never imports it and has no CPU path.

Every FLOP of the reference's hot path executes inside PyTorch CPU kernels
(SURVEY.md section 8c), so the restatement is written against the same
library in fp32, functionally, over a plain ``dict`` of tensors keyed by the
reference's ``state_dict`` names.  What it follows (citations into
``/root/reference``):

* architecture / parameter creation order .... ``Training_VAE.py:124-167``
* dataflow and concat orders .................. ``Training_VAE.py:180-226``
* five-term loss .............................. ``Training_VAE.py:229-268``
* relative-offset transform ................... ``Training_VAE.py:345-348``
* train step (zero_grad/forward/backward/step)  ``Training_VAE.py:351-363``
* generate + start-offset add ................. ``Tools.py:44-63``, ``Tools.py:898-912``
* Adam update ................................. torch ``optim/adam.py::_single_tensor_adam``
  (third-party, un-pinned by the reference; container version torch 2.11.0)

Pinning: the reference has no tests for this path ("parity unpinned" by the
reference itself).  This oracle is pinned instead against (a) the reference's
own classes imported in the build container (``tests/test_oracle_vs_reference.py``,
runs only where ``/root/reference`` is mounted) and (b) golden vectors those
classes produced, committed under ``tests/golden`` by ``oracle/make_golden.py``
(``tests/test_oracle_golden.py``, runs everywhere).
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

HIDDEN = 128
LOSS_KEYS = ("total_loss", "recon_loss", "kld_loss", "start_loss", "time_loss")
# weights the reference's __main__ passes (Training_VAE.py:300-306)
SCRIPT_WEIGHTS = (0.1, 0.1, 1.0, 1.0)
# defaults of conditional_vae_loss itself (Training_VAE.py:229)
DEFAULT_WEIGHTS = (0.1, 0.1, 1.0, 0.5)

Params = Dict[str, torch.Tensor]


def param_shapes(seq_len: int, latent_dim: int, dim: int = 3, hidden: int = HIDDEN) -> "OrderedDict[str, tuple]":
    """state_dict keys in creation order with (out, in) shapes
    (Training_VAE.py:132-167)."""
    I = seq_len * dim
    H = hidden
    L = latent_dim
    return OrderedDict(
        [
            ("condition_encoder.0.weight", (H, 2)), ("condition_encoder.0.bias", (H,)),
            ("condition_encoder.2.weight", (H, H)), ("condition_encoder.2.bias", (H,)),
            ("encoder.1.weight", (H, I)), ("encoder.1.bias", (H,)),
            ("encoder.3.weight", (H, H)), ("encoder.3.bias", (H,)),
            ("encoder.5.weight", (H, H)), ("encoder.5.bias", (H,)),
            ("encoder.7.weight", (H, H)), ("encoder.7.bias", (H,)),
            ("fc_mu.weight", (L, 2 * H)), ("fc_mu.bias", (L,)),
            ("fc_logvar.weight", (L, 2 * H)), ("fc_logvar.bias", (L,)),
            ("decoder.0.weight", (H, L + H)), ("decoder.0.bias", (H,)),
            ("decoder.2.weight", (H, H)), ("decoder.2.bias", (H,)),
            ("decoder.4.weight", (H, H)), ("decoder.4.bias", (H,)),
            ("decoder.6.weight", (I, H)), ("decoder.6.bias", (I,)),
        ]
    )


def init_params(seq_len: int, latent_dim: int, seed: Optional[int] = None, dim: int = 3,
                hidden: int = HIDDEN) -> Params:
    """Default ``nn.Linear`` initialisation in the reference's constructor order
    (``kaiming_uniform_(a=sqrt(5))`` for the weight, then the bias, both
    U(+-1/sqrt(fan_in)); torch ``nn/modules/linear.py::reset_parameters``), so
    that ``torch.manual_seed(s)`` followed by this call yields the tensors the
    reference's ``ConditionalTrajectoryVAE(seq_len, dim, latent_dim)`` would
    hold under the same seed."""
    if seed is not None:
        torch.manual_seed(seed)
    p: Params = OrderedDict()
    shapes = param_shapes(seq_len, latent_dim, dim, hidden)
    names = list(shapes)
    for wname, bname in zip(names[0::2], names[1::2]):
        out_f, in_f = shapes[wname]
        w = torch.empty(out_f, in_f, dtype=torch.float32)
        torch.nn.init.kaiming_uniform_(w, a=math.sqrt(5))
        bound = 1.0 / math.sqrt(in_f) if in_f > 0 else 0.0
        b = torch.empty(out_f, dtype=torch.float32)
        torch.nn.init.uniform_(b, -bound, bound)
        p[wname], p[bname] = w, b
    return p


def clone_params(p: Params, requires_grad: bool = False) -> Params:
    return OrderedDict((k, v.detach().clone().requires_grad_(requires_grad)) for k, v in p.items())


# --------------------------------------------------------------------------- model
def condition_encoder(p: Params, c: torch.Tensor) -> torch.Tensor:
    """Training_VAE.py:132-137."""
    h = F.relu(F.linear(c, p["condition_encoder.0.weight"], p["condition_encoder.0.bias"]))
    return F.relu(F.linear(h, p["condition_encoder.2.weight"], p["condition_encoder.2.bias"]))


def encoder(p: Params, x: torch.Tensor, keep: Optional[dict] = None) -> torch.Tensor:
    """Training_VAE.py:141-151 (Flatten then 4x Linear+ReLU)."""
    h = x.reshape(x.shape[0], -1)
    for i in (1, 3, 5, 7):
        h = F.relu(F.linear(h, p[f"encoder.{i}.weight"], p[f"encoder.{i}.bias"]))
        if keep is not None:
            keep[f"enc{i}"] = h
    return h


def encode(p: Params, x: torch.Tensor, start_points: torch.Tensor, keep: Optional[dict] = None):
    """Training_VAE.py:180-197: cat order is [h_traj, h_condition]."""
    h_traj = encoder(p, x, keep)
    h_c = condition_encoder(p, start_points)
    h = torch.cat([h_traj, h_c], dim=1)
    mu = F.linear(h, p["fc_mu.weight"], p["fc_mu.bias"])
    logvar = F.linear(h, p["fc_logvar.weight"], p["fc_logvar.bias"])
    return mu, logvar, h_c


def reparameterize(mu: torch.Tensor, logvar: torch.Tensor, eps: torch.Tensor) -> torch.Tensor:
    """Training_VAE.py:199-206 with the noise injected instead of drawn."""
    std = torch.exp(0.5 * logvar)
    return mu + eps * std


def decode(p: Params, z: torch.Tensor, h_c: torch.Tensor, keep: Optional[dict] = None) -> torch.Tensor:
    """Training_VAE.py:208-215: cat order is [z, condition]; no output activation."""
    h = torch.cat([z, h_c], dim=1)
    for i in (0, 2, 4):
        h = F.relu(F.linear(h, p[f"decoder.{i}.weight"], p[f"decoder.{i}.bias"]))
        if keep is not None:
            keep[f"dec{i}"] = h
    out = F.linear(h, p["decoder.6.weight"], p["decoder.6.bias"])
    return out.reshape(out.shape[0], -1, 3)


def forward(p: Params, x_rel: torch.Tensor, start_points: torch.Tensor, eps: torch.Tensor,
            keep: Optional[dict] = None):
    """Training_VAE.py:217-226 -> (recon_x, mu, logvar, condition)."""
    mu, logvar, h_c = encode(p, x_rel, start_points, keep)
    z = reparameterize(mu, logvar, eps)
    recon = decode(p, z, h_c, keep)
    if keep is not None:
        keep.update(mu=mu, logvar=logvar, h_c=h_c, z=z, recon=recon)
    return recon, mu, logvar, h_c


def offset_transform(batch: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """Training_VAE.py:345-348: subtract the start point from the x/y columns,
    leave the time column alone.  Returns (batch_rel, start_points)."""
    start_points = batch[:, 0, 1:3]
    batch_rel = batch.clone()
    batch_rel[:, :, 1:3] = batch_rel[:, :, 1:3] - start_points.unsqueeze(1)
    return batch_rel, start_points


def vae_loss(recon_x, x, mu, logvar, recon_weight=0.1, kld_weight=0.1, start_weight=1.0, time_weight=0.5):
    """Training_VAE.py:229-268.  Note the KLD is a *mean* over batch and latent
    (``:243``), the start target is the (zero) relative start (``:250-252``) and
    the time term is start-at-zero MSE plus mean relu(-dt) (``:258-264``)."""
    recon_loss = F.mse_loss(recon_x, x, reduction="mean")
    kld = -0.5 * torch.mean(1 + logvar - mu.pow(2) - logvar.exp())
    start_loss = 0
    if start_weight > 0:
        start_loss = F.mse_loss(recon_x[:, 0, 1:3], x[:, 0, 1:3], reduction="mean")
    time_loss = 0
    if time_weight > 0:
        t0 = recon_x[:, 0, 0]
        time_start = F.mse_loss(t0, torch.zeros_like(t0), reduction="mean")
        dt = recon_x[:, 1:, 0] - recon_x[:, :-1, 0]
        time_loss = time_start + torch.mean(torch.relu(-dt))
    total = recon_weight * recon_loss + kld_weight * kld + start_weight * start_loss + time_weight * time_loss
    return total, recon_loss, kld, start_loss, time_loss


# --------------------------------------------------------------------------- training
def loss_and_grads(p: Params, batch: torch.Tensor, eps: torch.Tensor, weights=SCRIPT_WEIGHTS):
    """One forward + backward of Training_VAE.py:345-362 on absolute
    trajectories ``batch`` (B,T,3) with injected noise ``eps`` (B,L).
    Returns (losses[5] as python floats, grads dict, intermediates dict)."""
    q = clone_params(p, requires_grad=True)
    batch_rel, start_points = offset_transform(batch)
    keep: dict = {}
    recon, mu, logvar, _ = forward(q, batch_rel, start_points, eps, keep)
    losses = vae_loss(recon, batch_rel, mu, logvar, *weights)
    losses[0].backward()
    grads = OrderedDict((k, v.grad.detach().clone()) for k, v in q.items())
    keep = {k: v.detach() for k, v in keep.items()}
    keep["x_rel"] = batch_rel
    keep["start"] = start_points
    return [float(l.detach()) if torch.is_tensor(l) else float(l) for l in losses], grads, keep


class AdamState:
    """Restatement of torch's single-tensor Adam (``optim/adam.py``; defaults
    betas (0.9, 0.999), eps 1e-8, weight_decay 0, amsgrad off), the optimiser
    the reference constructs at Training_VAE.py:332."""

    def __init__(self, p: Params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        self.lr, self.betas, self.eps = lr, betas, eps
        self.t = 0
        self.m = OrderedDict((k, torch.zeros_like(v)) for k, v in p.items())
        self.v = OrderedDict((k, torch.zeros_like(v)) for k, v in p.items())

    def step(self, p: Params, grads: Params) -> None:
        b1, b2 = self.betas
        self.t += 1
        bc1 = 1 - b1 ** self.t
        bc2 = 1 - b2 ** self.t
        step_size = self.lr / bc1
        bc2_sqrt = bc2 ** 0.5
        for k in p:
            g, m, v = grads[k], self.m[k], self.v[k]
            m.lerp_(g, 1 - b1)
            v.mul_(b2).addcmul_(g, g, value=1 - b2)
            denom = (v.sqrt() / bc2_sqrt).add_(self.eps)
            p[k].addcdiv_(m, denom, value=-step_size)


def train_steps(p: Params, batch: torch.Tensor, eps_per_step: torch.Tensor, weights=SCRIPT_WEIGHTS,
                lr=1e-3, adam: Optional[AdamState] = None):
    """Run ``eps_per_step.shape[0]`` full-batch steps in place on ``p``.
    Returns (loss history (steps,5) float64 ndarray, adam state)."""
    adam = adam or AdamState(p, lr=lr)
    hist = np.zeros((eps_per_step.shape[0], 5), dtype=np.float64)
    for s in range(eps_per_step.shape[0]):
        losses, grads, _ = loss_and_grads(p, batch, eps_per_step[s], weights)
        adam.step(p, grads)
        hist[s] = losses
    return hist, adam


# --------------------------------------------------------------------------- generation
@torch.no_grad()
def generate(p: Params, z: torch.Tensor, start_points: torch.Tensor, add_start: bool = True) -> torch.Tensor:
    """Tools.py:44-63 / Tools.py:898-912 batched: relative trajectory from
    (z, condition_encoder(start)), then global = fp32(start) + rel on the x, y
    columns as ONE fp32 add each (the fp32 definition of SURVEY.md section 8a
    row 13); the time column is untouched."""
    start_points = start_points.to(torch.float32)
    if start_points.shape[0] == 1 and z.shape[0] != 1:
        start_points = start_points.expand(z.shape[0], 2)
    h_c = condition_encoder(p, start_points)
    out = decode(p, z, h_c).clone()
    if add_start:
        out[:, :, 1] = start_points[:, 0:1] + out[:, :, 1]
        out[:, :, 2] = start_points[:, 1:2] + out[:, :, 2]
    return out


# --------------------------------------------------------------------------- algorithmic work
def macs(seq_len: int, latent_dim: int, hidden: int = HIDDEN) -> dict:
    I, L, H = 3 * seq_len, latent_dim, hidden
    cond = 2 * H + H * H
    enc = I * H + 3 * H * H
    heads = 4 * H * L
    dec = (L + H) * H + 2 * H * H + I * H
    fwd = cond + enc + heads + dec
    return dict(cond=cond, enc=enc, heads=heads, dec=dec, fwd=fwd,
                decode=cond + dec, decode_shared=dec - H * H,  # hoisted condition: (L)H + 2H^2 + IH
                train=3 * fwd - (2 * H + I * H))


def flops(seq_len: int, latent_dim: int) -> dict:
    return {k: 2 * v for k, v in macs(seq_len, latent_dim).items()}


def n_params(seq_len: int, latent_dim: int) -> int:
    return sum(int(np.prod(s)) for s in param_shapes(seq_len, latent_dim).values())
