"""The oracle restatement against the UNMODIFIED reference classes, imported
from /root/reference.  Runs only in the build container (the reference does not
exist on the GPU box); the committed goldens cover the same ground elsewhere."""
import os

import numpy as np
import pytest
import torch

from oracle import vae_oracle as O
from oracle.ref_loader import REFERENCE_ROOT, load_reference, reference_available

pytestmark = pytest.mark.skipif(not reference_available(), reason="reference not mounted")
torch.set_num_threads(1)


@pytest.fixture(scope="module")
def ref():
    return load_reference()


@pytest.mark.parametrize("T,L", [(10, 8), (12, 8), (50, 16), (400, 64)])
def test_forward_loss_backward_match(ref, T, L):
    torch.manual_seed(11)
    model = ref.Training_VAE.ConditionalTrajectoryVAE(T, 3, L)
    p = {k: v.detach().clone() for k, v in model.state_dict().items()}
    assert list(p) == list(O.param_shapes(T, L))
    g = torch.Generator().manual_seed(2)
    batch = torch.randn(33, T, 3, generator=g) * 3 + torch.tensor([0.0, 150.0, -40.0])
    eps = torch.randn(33, L, generator=g)
    rel, start = O.offset_transform(batch)
    orig = torch.randn_like
    torch.randn_like = lambda t, *a, **k: eps
    try:
        recon, mu, lv, cond = model(rel, start)
    finally:
        torch.randn_like = orig
    out = ref.Training_VAE.conditional_vae_loss(recon, rel, mu, lv, cond, 0.1, 0.1, 1.0, 1.0)
    out[0].backward()
    losses, grads, keep = O.loss_and_grads(p, batch, eps, O.SCRIPT_WEIGHTS)
    np.testing.assert_array_equal(keep["recon"].numpy(), recon.detach().numpy())
    np.testing.assert_array_equal(np.array(losses), np.array([float(o.detach()) for o in out]))
    for k, v in model.named_parameters():
        np.testing.assert_array_equal(grads[k].numpy(), v.grad.numpy())


def test_every_shipped_offset_checkpoint_loads_into_the_layout(ref):
    mdir = os.path.join(REFERENCE_ROOT, "training", "models")
    names = sorted(f for f in os.listdir(mdir) if f.startswith("vae_offset_"))
    assert len(names) == 14
    for f in names:
        sd = torch.load(os.path.join(mdir, f), map_location="cpu")
        T = sd["decoder.6.bias"].numel() // 3
        L = sd["fc_mu.bias"].numel()
        shapes = O.param_shapes(T, L)
        assert list(sd) == list(shapes), f
        assert all(tuple(sd[k].shape) == s and sd[k].dtype == torch.float32 for k, s in shapes.items()), f


def test_adam_restatement_matches_torch_adam(ref):
    p = O.init_params(10, 8, seed=4)
    q = {k: torch.nn.Parameter(v.clone()) for k, v in p.items()}
    opt = torch.optim.Adam(q.values(), lr=1e-3)
    adam = O.AdamState(p)
    g = torch.Generator().manual_seed(8)
    for _ in range(7):
        grads = {k: torch.randn(v.shape, generator=g) * 10 ** float(torch.randint(-6, 3, (1,), generator=g)) for k, v in p.items()}
        for k in q:
            q[k].grad = grads[k].clone()
        opt.step()
        adam.step(p, grads)
    for k in p:
        np.testing.assert_array_equal(p[k].numpy(), q[k].detach().numpy())


def test_reg157_reference_behaviour(ref):
    f = ref.Driver_Models.Reg157
    assert f(0.0, 20.0, 100.0, 10.0) == -6          # ttc 10 > 10/12+0.35
    assert f(0.0, 20.0, 5.0, 10.0) is None          # ttc 0.5 < 1.18
    with pytest.raises(ZeroDivisionError):
        f(0.0, 10.0, 5.0, 10.0)
