"""CPU-side checks of the C ABI (no compute calls): libdmvae.so loads, exports every function that
include/dmvae.h declares, the ctypes table of dmvae/_lib.py mirrors the header one to one, the pure host
queries answer, and a compute call without a B200 fails loudly instead of falling back."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "dmvae.h")
DEBUG_HEADER = os.path.join(ROOT, "include", "dmvae_debug.h")


def declared_functions(header=HEADER):
    text = open(header).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)          # comments
    return sorted(set(re.findall(r"\b(dmvae_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    from dmvae import _lib
    if not os.path.isfile(_lib.LIB_PATH):
        pytest.skip("libdmvae.so not built (python -c 'import __graft_entry__ as g; g.build()')")
    return _lib.lib()


def test_header_declares_the_expected_surface():
    names = declared_functions()
    for must in ("dmvae_decode", "dmvae_train_step", "dmvae_train_step_dev", "dmvae_train_step_dp", "dmvae_adam_step",
                 "dmvae_forward", "dmvae_backward", "dmvae_loss", "dmvae_pack_weights", "dmvae_last_error"):
        assert must in names
    assert len(names) >= 30


def test_library_exports_every_declared_symbol(lib):
    missing = [n for n in declared_functions() if not hasattr(lib, n)]
    assert not missing, f"declared in include/dmvae.h but not exported: {missing}"


def test_ctypes_table_mirrors_the_header():
    from dmvae import _lib
    header, table = set(declared_functions()), set(_lib.SIGNATURES)
    assert table <= header, f"bound but not declared: {sorted(table - header)}"
    # everything the Python layer does not bind is listed here on purpose (debug / unused-by-Python entry points)
    assert header == table, sorted(header ^ table)
    # the development aids live in their own header, outside the drop-in boundary
    assert not [n for n in header if n.startswith("dmvae_debug_")]
    assert set(declared_functions(DEBUG_HEADER)) == set(_lib.DEBUG_SIGNATURES)


def test_struct_layouts_match_the_header():
    from dmvae import _lib
    assert ctypes.sizeof(_lib.DmvaeCfg) == 16
    assert ctypes.sizeof(_lib.DmvaeLossWeights) == 16
    assert ctypes.sizeof(_lib.DmvaeAdam) == 40
    assert _lib.MAX_PEERS == 8 and ctypes.sizeof(_lib.DmvaeDpPeers) == 16 + 8 * 8
    assert ctypes.sizeof(_lib.DmvaeMpcCfg) == 6 * 4 + 8 * 8
    text = open(HEADER).read()
    assert "#define DMVAE_MAX_PEERS 8" in text
    assert f"#define DMVAE_KERNEL_COUNT {_lib.KERNEL_COUNT}" in text


def test_host_side_queries(lib):
    from dmvae import _lib
    assert lib.dmvae_abi_version() == _lib.ABI_VERSION == 3
    assert f"#define DMVAE_ABI_VERSION {_lib.ABI_VERSION}" in open(HEADER).read()
    cfg = _lib.cfg(10, 8)
    assert lib.dmvae_param_count(ctypes.byref(cfg)) == 128942          # T = 10, L = 8
    assert lib.dmvae_grad_count(ctypes.byref(cfg)) == 128942 + 5
    assert lib.dmvae_packed_count(ctypes.byref(cfg)) > 128942
    assert lib.dmvae_dp_inbox_bytes(ctypes.byref(cfg), 8) == (8 + 1) * 2 * 128948 * 8 + 16
    cfg400 = _lib.cfg(400, 64)
    assert lib.dmvae_grad_count(ctypes.byref(cfg400)) == 465589        # T = 400, L = 64 (+ 5 loss terms)
    bad = _lib.cfg(401, 8)
    assert lib.dmvae_param_count(ctypes.byref(bad)) < 0 and b"seq_len" in lib.dmvae_last_error()
    mpc = _lib.DmvaeMpcCfg(n_way=10, way_f32=1, horizon=30, blocks=20, max_iter=50, reserved=0, wheelbase=2.8, max_steer=0.5,
                           max_accel=7.0, q_theta=20.0, q_v=5.0, r_accel=1.0, r_steer=50.0, tol=1e-11)
    assert lib.dmvae_mpc_workspace_bytes(ctypes.byref(mpc), 1000) == (10 + 8 * 9 + 8 + 40) * 1000 * 8
    mpc.n_way = 1
    assert lib.dmvae_mpc_workspace_bytes(ctypes.byref(mpc), 1000) < 0 and b"n_way" in lib.dmvae_last_error()
    names = {lib.dmvae_kernel_name(i).decode() for i in range(_lib.KERNEL_COUNT)}
    assert {"mpc_prepare_kernel", "mpc_track_kernel"} <= names
    assert {"decode_tc_kernel", "chain_kernel", "wgrad_kernel", "reduce_tc_kernel", "train_tc_fused_kernel"} <= names


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the behaviour WITHOUT a CUDA device")
def test_compute_calls_fail_loudly_without_a_device(lib):
    """No CPU path exists: the C entry point reports DMVAE_ERR_DEVICE, the Python surface raises."""
    from dmvae import ConditionalTrajectoryVAE, _lib
    cfg = _lib.cfg(10, 8)
    buf = (ctypes.c_float * 64)()
    rc = lib.dmvae_decode(ctypes.byref(cfg), buf, None, 0, 0, buf, 1, buf, None, 1, 1, None)
    assert rc < 0 and lib.dmvae_last_error()
    model = ConditionalTrajectoryVAE(10, 3, 8)
    with pytest.raises(Exception) as err:
        model.generate(torch.zeros(1, 2), n=4)
    assert "oracle" not in str(err.value).lower()


def test_resident_row_is_a_permutation_per_epoch(lib):
    """dmvae_resident_row (host evaluation of the function the kernel uses to pick the rows of a shuffled resident
    set): a bijection of [0, n) for every epoch, different between epochs and seeds, and loud on bad arguments."""
    for n in (1, 2, 3, 255, 256, 257, 5000):
        for epoch in (0, 1, 1000):
            order = [lib.dmvae_resident_row(7, epoch, p, n) for p in range(n)]
            assert sorted(order) == list(range(n)), (n, epoch)
    a = [lib.dmvae_resident_row(7, 0, p, 5000) for p in range(5000)]
    b = [lib.dmvae_resident_row(7, 1, p, 5000) for p in range(5000)]
    c = [lib.dmvae_resident_row(8, 0, p, 5000) for p in range(5000)]
    assert a != b and a != c and sum(x == y for x, y in zip(a, b)) < 50
    assert sum(x == i for i, x in enumerate(a)) < 50              # not the identity
    assert lib.dmvae_resident_row(7, 0, 5, 5) < 0 and lib.dmvae_resident_row(7, -1, 0, 5) < 0
