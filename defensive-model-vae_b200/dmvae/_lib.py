"""ctypes binding of libdmvae.so (the C ABI in include/dmvae.h).

The library is built in-tree by ``defensive-model-vae_b200/csrc/build.sh``
(``__graft_entry__.build()``).  There is no CPU implementation behind this
module: if the shared object is missing, or the device is not a B200, every
compute call raises.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, byref, c_char_p, c_float, c_int, c_int32, c_int64, c_uint64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "csrc", "libdmvae.so")


class DmvaeError(RuntimeError):
    pass


class DmvaeCfg(ctypes.Structure):
    _fields_ = [("seq_len", c_int32), ("dim", c_int32), ("latent_dim", c_int32), ("hidden_dim", c_int32)]


class DmvaeLossWeights(ctypes.Structure):
    _fields_ = [("recon", c_float), ("kld", c_float), ("start", c_float), ("time", c_float)]


class DmvaeAdam(ctypes.Structure):
    _fields_ = [("lr", ctypes.c_double), ("beta1", ctypes.c_double), ("beta2", ctypes.c_double),
                ("eps", ctypes.c_double), ("step", c_int64)]


MAX_PEERS = 8


class DmvaeDpPeers(ctypes.Structure):
    _fields_ = [("world", c_int32), ("rank", c_int32), ("owned_from", c_int32), ("timeout_ms", c_int32),
                ("inbox", c_void_p * MAX_PEERS)]


class DmvaeMpcCfg(ctypes.Structure):
    _fields_ = [("n_way", c_int32), ("way_f32", c_int32), ("horizon", c_int32), ("blocks", c_int32), ("max_iter", c_int32),
                ("reserved", c_int32), ("wheelbase", ctypes.c_double), ("max_steer", ctypes.c_double), ("max_accel", ctypes.c_double),
                ("q_theta", ctypes.c_double), ("q_v", ctypes.c_double), ("r_accel", ctypes.c_double), ("r_steer", ctypes.c_double),
                ("tol", ctypes.c_double)]


# name -> (restype, argtypes); mirrors include/dmvae.h one to one
_P = c_void_p
_CFG = POINTER(DmvaeCfg)
SIGNATURES = {
    "dmvae_abi_version": (c_int, []),
    "dmvae_last_error": (c_char_p, []),
    "dmvae_device_sm_count": (c_int, []),
    "dmvae_param_count": (c_int64, [_CFG]),
    "dmvae_param_offset": (c_int64, [_CFG, c_int]),
    "dmvae_packed_count": (c_int64, [_CFG]),
    "dmvae_pack_weights": (c_int, [_CFG, _P, _P, _P]),
    "dmvae_decode": (c_int, [_CFG, _P, _P, c_uint64, c_uint64, _P, c_int, _P, _P, c_int64, c_int, _P]),
    "dmvae_grad_count": (c_int64, [_CFG]),
    "dmvae_train_workspace_bytes": (c_int64, [_CFG, c_int64]),
    "dmvae_train_fwd_bwd": (c_int, [_CFG, _P, _P, _P, c_uint64, c_uint64, c_uint64, POINTER(DmvaeLossWeights),
                                    c_float, c_int64, _P, _P, _P]),
    "dmvae_train_step": (c_int, [_CFG, _P, _P, _P, _P, _P, _P, c_uint64, c_uint64, POINTER(DmvaeLossWeights),
                                 c_float, c_int64, POINTER(DmvaeAdam), _P, _P, _P]),
    "dmvae_train_step_dev": (c_int, [_CFG, _P, _P, _P, _P, _P, _P, c_uint64, c_uint64, POINTER(DmvaeLossWeights),
                                     c_float, c_int64, POINTER(DmvaeAdam), _P, _P, _P, _P]),
    "dmvae_resident_row": (c_int64, [c_uint64, c_int64, c_int64, c_int64]),
    "dmvae_train_step_resident": (c_int, [_CFG, _P, _P, _P, _P, _P, c_int64, c_int, c_uint64, c_uint64, c_uint64, POINTER(DmvaeLossWeights),
                                          c_float, c_int64, POINTER(DmvaeAdam), _P, _P, _P, POINTER(DmvaeDpPeers), _P]),
    "dmvae_train_fwd_bwd_dev": (c_int, [_CFG, _P, _P, _P, c_uint64, c_uint64, _P, POINTER(DmvaeLossWeights),
                                        c_float, c_int64, _P, _P, _P]),
    "dmvae_adam_step_dev": (c_int, [_CFG, _P, _P, _P, _P, POINTER(DmvaeAdam), _P, _P, _P]),
    "dmvae_dp_inbox_bytes": (c_int64, [_CFG, c_int]),
    "dmvae_dp_status": (c_int, [_CFG, POINTER(DmvaeDpPeers), _P]),
    "dmvae_train_step_dp": (c_int, [_CFG, _P, _P, _P, _P, _P, _P, c_uint64, c_uint64, POINTER(DmvaeLossWeights),
                                    c_float, c_int64, POINTER(DmvaeAdam), _P, _P, _P, POINTER(DmvaeDpPeers), _P]),
    "dmvae_adam_step": (c_int, [_CFG, _P, _P, _P, _P, POINTER(DmvaeAdam), _P, _P]),
    "dmvae_stash_bytes": (c_int64, [_CFG, c_int64]),
    "dmvae_forward": (c_int, [_CFG, _P, _P, _P, _P, _P, _P, _P, _P, _P, c_int64, _P]),
    "dmvae_backward": (c_int, [_CFG, _P, _P, _P, _P, _P, _P, _P, _P, c_int64, _P]),
    "dmvae_loss": (c_int, [_CFG, _P, _P, _P, _P, POINTER(DmvaeLossWeights), c_int64, _P, _P]),
    "dmvae_loss_backward": (c_int, [_CFG, _P, _P, _P, _P, POINTER(DmvaeLossWeights), c_int64, _P, _P, _P, _P, _P]),
    "dmvae_cond_encode": (c_int, [_CFG, _P, _P, _P, c_int64, _P]),
    "dmvae_decode_from_condition": (c_int, [_CFG, _P, _P, _P, _P, c_int64, _P]),
    "dmvae_set_decode_impl": (c_int, [c_int]),
    "dmvae_set_train_impl": (c_int, [c_int]),
    "dmvae_waypoint_speeds": (c_int, [_P, c_int64, c_int32, c_int32, _P, _P, _P]),
    "dmvae_histogram": (c_int, [_P, c_int64, POINTER(ctypes.c_double), c_int32, _P, _P]),
    "dmvae_trajectories_per_cell": (c_int, [_P, c_int64, c_int32, c_int32, ctypes.c_double, ctypes.c_double, c_int32,
                                            ctypes.c_double, ctypes.c_double, c_int32, _P, _P]),
    "dmvae_dense": (c_int, [_P, _P, _P, _P, c_int64, c_int32, c_int32, c_int32, _P]),
    "dmvae_mpc_workspace_bytes": (c_int64, [POINTER(DmvaeMpcCfg), c_int64]),
    "dmvae_mpc_prepare": (c_int, [POINTER(DmvaeMpcCfg), _P, _P, c_int64, _P, _P, _P, _P, _P]),
    "dmvae_mpc_track": (c_int, [POINTER(DmvaeMpcCfg), _P, c_int64, ctypes.c_double, _P, _P, c_int32, c_int32, _P, _P, _P, c_int64, _P, _P]),
    "dmvae_mpc_windows": (c_int, [POINTER(DmvaeMpcCfg), _P, c_int64, ctypes.c_double, POINTER(ctypes.c_double), c_int32, _P, _P, _P]),
    "dmvae_kernel_name": (c_char_p, [c_int]),
    "dmvae_launch_count": (c_int64, [c_int]),
    "dmvae_profile_begin": (c_int, []),
    "dmvae_profile_end": (c_int, [POINTER(ctypes.c_double), POINTER(c_int64), c_int]),
    "dmvae_ffma_probe": (c_int, [c_int64, _P, POINTER(ctypes.c_double), _P]),
    "dmvae_tf32_probe": (c_int, [c_int64, c_int, _P, POINTER(ctypes.c_double), _P]),
}
KERNEL_COUNT = 22
ABI_VERSION = 3
# include/dmvae_debug.h (development aids, outside the drop-in boundary)
DEBUG_SIGNATURES = {
    "dmvae_debug_decode_trace": (c_int, [_P]),
    "dmvae_debug_train_trace": (c_int, [_P]),
    "dmvae_debug_train_trace_tile": (c_int, [_P, c_int]),
}

_lib = None

def lib() -> ctypes.CDLL:
    """Load libdmvae.so (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise DmvaeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or defensive-model-vae_b200/csrc/build.sh).  dmvae has no CPU or PyTorch fallback.")
    handle = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in list(SIGNATURES.items()) + list(DEBUG_SIGNATURES.items()):
        fn = getattr(handle, name)  # AttributeError here = header/library mismatch
        fn.restype = res
        fn.argtypes = args
    got = handle.dmvae_abi_version()
    if got != ABI_VERSION:
        raise DmvaeError(f"libdmvae ABI version {got}, expected {ABI_VERSION}: rebuild (csrc/build.sh)")
    _lib = handle
    return _lib


def last_error() -> str:
    return (lib().dmvae_last_error() or b"").decode("utf-8", "replace")


def check(rc: int, what: str) -> int:
    if rc < 0:
        raise DmvaeError(f"{what} failed ({rc}): {last_error()}")
    return rc


def cfg(seq_len: int, latent_dim: int, dim: int = 3, hidden_dim: int = 128) -> DmvaeCfg:
    return DmvaeCfg(int(seq_len), int(dim), int(latent_dim), int(hidden_dim))


def ptr(t) -> c_void_p:
    """Device pointer of a torch tensor (None -> NULL)."""
    if t is None:
        return c_void_p(0)
    return c_void_p(t.data_ptr())


def stream_ptr() -> c_void_p:
    import torch
    return c_void_p(torch.cuda.current_stream().cuda_stream)


__all__ = ["lib", "check", "cfg", "ptr", "stream_ptr", "byref", "DmvaeCfg", "DmvaeLossWeights", "DmvaeAdam",
           "DmvaeDpPeers", "MAX_PEERS", "DmvaeError", "SIGNATURES", "LIB_PATH", "last_error"]
