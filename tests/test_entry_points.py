"""The path through the entry points a user of the reference calls - not through the kernels' own Python
surface: ``Tools.load_model_and_generate_trajectory`` (reference Tools.py:18-65), ``Training_VAE.train`` (the
reference's training mode, Training_VAE.py:316-394), ``Tools.generate_for_visualization`` (the decode block of
``visualize_trajectories``, Tools.py:862-912), ``LossMeter`` (:366-380), ``Driver_Models.Reg157``
(Driver_Models.py:2-9) and the generate -> track hand-off (Distribution.py:51-111).

Goldens: tests/golden/{generate_api,train_entry,visualize_entry,reg157}.npz were produced by running the
reference's own code (oracle/make_golden.py, oracle/make_golden_entry.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import vae_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ckpt(tmp_path, golden_dir, sce):
    ck = np.load(os.path.join(golden_dir, f"ckpt_{sce}_cond.npz"))
    path = str(tmp_path / f"vae_offset_{sce}_cond_ld8_epoch3000.pth")
    torch.save({k: torch.from_numpy(ck[k]) for k in ck.files}, path)
    return path, {k: torch.from_numpy(ck[k]) for k in ck.files}


# ------------------------------------------------------------------------------------------ CPU
def test_reg157_drop_in_equals_the_reference_table(golden_dir):
    """The repo's Driver_Models.Reg157 against outputs of the reference's function (None recorded as NaN)."""
    import Driver_Models
    assert os.path.dirname(os.path.abspath(Driver_Models.__file__)) == ROOT
    g = np.load(os.path.join(golden_dir, "reg157.npz"))
    rows = [a for a in g["args"] if a[1] != a[3]]
    assert len(rows) == len(g["out"]) >= 200
    for (x_e, v_e, x_f, v_f), want in zip(rows, g["out"]):
        got = Driver_Models.Reg157(x_e, v_e, x_f, v_f)
        assert (got is None and np.isnan(want)) or got == want
    assert {-6.0} <= set(g["out"][~np.isnan(g["out"])]) and np.isnan(g["out"]).any()      # both branches covered
    with pytest.raises(ZeroDivisionError):                                                # reference behaviour, kept
        Driver_Models.Reg157(0.0, 5.0, 10.0, 5.0)


def test_loss_meter_is_the_reference_bookkeeping():
    """LossMeter = the sample-weighted sums of Training_VAE.py:339,:366-380 (five ``.item()`` per step, divided by
    the dataset size per epoch), kept in one tensor.  Runs wherever a tensor can live (CPU here)."""
    from dmvae.train import LOSS_KEYS, LossMeter
    assert LOSS_KEYS == ("total_loss", "recon_loss", "kld_loss", "start_loss", "time_loss")        # Training_VAE.py:337
    rng = np.random.default_rng(3)
    steps = [(16, rng.normal(size=5)), (16, rng.normal(size=5)), (6, rng.normal(size=5))]           # ragged last batch
    meter = LossMeter("cpu")
    total = np.zeros(5)
    for n, vals in steps:
        meter.update(torch.tensor(vals, dtype=torch.float32), n)
        total += np.float32(vals).astype(np.float64) * n           # loss.item() * batch.size(0)
    np.testing.assert_allclose(meter.means(), total / 38, rtol=1e-12)
    assert meter.count == 0 and float(meter.sums.abs().sum()) == 0.0     # reset for the next epoch
    meter.update(torch.ones(5), 4)
    assert meter.means() == [1.0] * 5


# ------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
def test_load_model_and_generate_trajectory_reproduces_the_reference_call(tmp_path, golden_dir):
    """Seeded call of the drop-in == the trajectory the reference's function returned under the same seed
    (generate_api.npz): same checkpoint, same RNG stream (module construction burns the initialisation draws,
    then z = torch.randn(1, L)), fp32 offset add."""
    import Tools
    path, _ = _ckpt(tmp_path, golden_dir, "sce1")
    api = np.load(os.path.join(golden_dir, "generate_api.npz"))
    for seed, (sx, sy) in ((123, (-194.25, 19.0)), (7, (np.float32(-193.77), np.float32(18.82)))):
        torch.manual_seed(seed)
        got = Tools.load_model_and_generate_trajectory(path, sx, sy, seq_len=10, dim=3, latent_dim=8, device="cpu")
        want = api[f"seed{seed}"]
        assert got.shape == (10, 3) and got.dtype == np.float32
        assert np.abs(got - want).max() <= 1e-5 * np.abs(want).max()
        assert np.abs(got[:, 0] - want[:, 0]).max() <= 1e-5 * max(np.abs(want[:, 0]).max(), 1.0)   # time column: no offset
    # the generator was advanced exactly as by the reference (its next draw is the same)
    torch.manual_seed(123)
    O.init_params(10, 8)
    torch.randn(1, 8)
    want_next = torch.randn(3)
    torch.manual_seed(123)
    Tools.load_model_and_generate_trajectory(path, -194.25, 19.0, seq_len=10, dim=3, latent_dim=8)
    assert torch.equal(torch.randn(3), want_next)


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["b38", "b16"])
def test_training_entry_point_reproduces_the_reference_run(tmp_path, golden_dir, tag):
    """``Training_VAE.train`` under the seed of the golden run == the reference's training mode executed from its
    own source (train_entry.npz): DataLoader shuffle, default initialisation and reparameterisation noise come
    from the same host generator; loss history (weight-scaled component terms), the CSV plot_losses writes, and the
    saved state_dict (keys, order, shapes, values).  b38: the reference configuration (one full-batch step per
    epoch); b16: three steps per epoch, the last one ragged."""
    import Training_VAE
    g = np.load(os.path.join(golden_dir, "train_entry.npz"))
    seed, epochs, bs = int(g[f"{tag}/seed"]), int(g[f"{tag}/epochs"]), int(g[f"{tag}/batch_size"])
    save, loss_png = str(tmp_path / "m" / "model.pth"), str(tmp_path / "loss" / "curve.png")
    torch.manual_seed(seed)
    model, hist = Training_VAE.train(os.path.join(golden_dir, "data_sce1_cond.npy"), seq_len=10, dim=3, latent_dim=8,
                                     batch_size=bs, lr=1e-3, epochs=epochs, device="cpu", recon_weight=0.1, kld_weight=0.1,
                                     start_weight=1.0, time_weight=1.0, model_save_path=save, loss_save_path=loss_png,
                                     verbose=False)
    assert list(hist.keys()) == list(g["keys"])
    got = np.array([hist[k] for k in g["keys"]])
    want = g[f"{tag}/hist"]
    assert got.shape == want.shape == (5, epochs)
    # the first step is a plain forward (no drift yet); later epochs compound fp32 summation-order differences
    # through Adam (chaotic): 1e-4 on the first epochs, 2 % on all (tests/test_train_gpu.py)
    rel = np.abs(got - want) / np.maximum(np.abs(want), 1e-6)
    assert rel[:, 0].max() < 2e-5, rel[:, 0]
    assert rel[:, : min(epochs, 4)].max() < 1e-3, rel[:, :4]
    assert rel.max() < 2e-2, rel
    # the CSV next to the PNG: header = the five keys (Tools.py:747-771)
    import csv
    rows = list(csv.reader(open(os.path.splitext(loss_png)[0] + ".csv")))
    assert rows[0] == list(g["keys"]) and len(rows) == epochs + 1
    np.testing.assert_allclose(np.array(rows[1:], dtype=np.float64).T, got, rtol=1e-12)
    # the checkpoint: a plain state_dict with the reference's keys in the reference's order
    sd = torch.load(save, map_location="cpu")
    assert list(sd.keys()) == list(g[f"{tag}/state_keys"])
    p0 = O.init_params(10, 8, seed=seed)            # the seeded default initialisation both runs started from
    num = den = 0.0
    for k, v in sd.items():
        assert v.dtype == torch.float32 and v.device.type == "cpu"
        want_d = g[f"{tag}/final_digest/{k}"]
        assert abs(float(v.double().sum()) - want_d[0]) <= 2e-3 * want_d[1] + 1e-9, k
        want_r = g[f"{tag}/final/{k}"]
        cut = (lambda a: a[:4] if a.size >= 128 * 128 else a)
        got_r, init_r = cut(v.numpy()), cut(p0[k].numpy())
        assert got_r.shape == want_r.shape, k
        num += float(((got_r - want_r).astype(np.float64) ** 2).sum())
        den += float(((want_r - init_r).astype(np.float64) ** 2).sum())
    # Adam turns a rounding-level difference of a near-zero gradient into a +-lr difference of that one parameter, so
    # single elements may drift; the learned update as a whole must be the reference's
    assert (num / den) ** 0.5 < 0.05, (num / den) ** 0.5
    # and it loads back into the drop-in and the oracle alike
    m2 = Training_VAE.ConditionalTrajectoryVAE(10, 3, 8)
    m2.load_state_dict(sd)
    assert torch.equal(m2.flat_parameters().cpu(), model.flat_parameters().cpu())


@pytest.mark.gpu
@pytest.mark.parametrize("tag,kwargs", [
    ("train_starts", dict(use_training_start_end=True, train_traj_start=3, train_traj_end=12)),
    ("custom_start", dict(use_training_start_end=False, custom_start_end=[(-194.0, 19.1), (0.0, 0.0)],
                          train_traj_start=0, train_traj_end=9))])
def test_generate_for_visualization_reproduces_the_reference_block(tmp_path, golden_dir, tag, kwargs):
    """The decode block of Tools.visualize_trajectories on the shipped sce1 checkpoint under the golden's seed
    (visualize_entry.npz: arrays taken from the reference function's own frame)."""
    import Tools
    import Training_VAE
    path, sd = _ckpt(tmp_path, golden_dir, "sce1")
    g = np.load(os.path.join(golden_dir, "visualize_entry.npz"))
    model = Training_VAE.ConditionalTrajectoryVAE(10, 3, 8)
    model.load_state_dict(sd)
    dataset = Training_VAE.TrajectoryDataset(os.path.join(golden_dir, "data_sce1_cond.npy"))
    torch.manual_seed(int(g[f"{tag}/seed"]))
    train_data, generated = Tools.generate_for_visualization(model, dataset, **kwargs)
    np.testing.assert_array_equal(train_data, g[f"{tag}/train_data"])
    want = g[f"{tag}/generated"]
    assert generated.shape == want.shape and generated.dtype == np.float32
    assert np.abs(generated - want).max() <= 1e-5 * np.abs(want).max()
    # the public wrapper returns the same arrays (no matplotlib here: it reports and returns)
    torch.manual_seed(int(g[f"{tag}/seed"]))
    _, again = Tools.visualize_trajectories(model, dataset, str(tmp_path / "vis.pth"), axis_flip="y", **kwargs)
    np.testing.assert_array_equal(again, generated)


@pytest.mark.gpu
def test_handoff_waypoints_equal_the_oracle_then_the_reference_lines(tmp_path, golden_dir):
    """dmvae.handoff.generate_tracker_jobs: the batched decode behind it against the ORACLE's generate for the same
    latents and start points, followed by the reference's own two lines (Distribution.py:77-78)."""
    from test_handoff import _write_csv
    from dmvae import handoff
    path, sd = _ckpt(tmp_path, golden_dir, "sce4")
    csvs, starts = [], []
    for i, (x, y) in enumerate([(14.2, 80.1), (15.9, -20.5), (16.1, 60.0), (13.25, -44.0), (16.7, 107.0)]):
        p = str(tmp_path / f"exp_{i + 1}_unpred_{i}.csv")
        _write_csv(p, x, y)
        csvs.append(p)
        starts.append([x, y])
    torch.manual_seed(77)
    jobs = handoff.generate_tracker_jobs(path, csvs, seq_len=10, dim=3, latent_dim=8)      # latents: the loop's own draws
    torch.manual_seed(77)
    z = torch.cat([torch.randn(1, 8) for _ in csvs], 0)                                      # Tools.py:46, once per CSV
    ref = O.generate(sd, z, torch.tensor(starts, dtype=torch.float64).float(), add_start=True).numpy()
    assert len(jobs) == len(csvs)
    for i, j in enumerate(jobs):
        waypoints = ref[i].copy()
        waypoints = waypoints[:, [1, 2, 0]]          # Distribution.py:77
        waypoints[0, 2] = 0.0                        # Distribution.py:78
        assert j.waypoints.shape == (10, 3) and j.waypoints[0, 2] == 0.0
        assert np.abs(j.waypoints - waypoints).max() <= 1e-5 * np.abs(waypoints).max()
        np.testing.assert_array_equal(j.initial_state[:2], np.array(starts[i]))
        assert j.trackable == bool(np.all(np.diff(j.waypoints[:, 2]) > 0))


@pytest.mark.gpu
def test_generate_scenarios_writes_the_bulk_files(tmp_path, golden_dir):
    """dmvae.parallel.generate_scenarios (configs[2] driver, one rank here): one (n, T, 3) float32 file per scenario,
    rows = the oracle's decode of the Philox latents the kernel reports."""
    from dmvae.parallel import generate_scenarios
    import Tools
    p1, sd1 = _ckpt(tmp_path, golden_dir, "sce1")
    p4, sd4 = _ckpt(tmp_path, golden_dir, "sce4")
    out = generate_scenarios([p1, p4], ["sce1", "sce4"], [(-193.3, 50.0), (11.0, 0.0)], 3001, out_dir=str(tmp_path / "gen"), seed=5)
    assert [os.path.basename(o) for o in out] == ["decoded_waypoints_sce1.npy", "decoded_waypoints_sce4.npy"]
    for o, path, sd, start in zip(out, (p1, p4), (sd1, sd4), ((-193.3, 50.0), (11.0, 0.0))):
        arr = np.load(o)
        assert arr.shape == (3001, 10, 3) and arr.dtype == np.float32
        model = Tools._cached_model(path, 10, 3, 8)
        same, z = model.generate(torch.tensor([start]), n=3001, seed=5, return_z=True)
        np.testing.assert_array_equal(arr, same.cpu().numpy())
        ref = O.generate(sd, z.cpu(), torch.tensor([start], dtype=torch.float32)).numpy()
        assert np.abs(arr - ref).max() <= 1e-5 * np.abs(ref).max()
