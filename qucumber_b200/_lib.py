"""ctypes binding of libqucumber_b200.so (C ABI declared in include/qcb.h).

There is no CPU fallback: if the shared library is missing or a call fails, a RuntimeError is
raised. PyTorch is used only to own device memory and streams.
"""
import ctypes
import os
from ctypes import c_char_p, c_double, c_int, c_int64, c_size_t, c_uint64, c_void_p

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("QCB_LIB_PATH") or os.path.join(_HERE, "libqucumber_b200.so")  # override: kernel experiments only

QCB_OK = 0
OP_GIBBS, OP_ENERGY, OP_GRAD_REDUCE, OP_LOGZ, OP_EXACT_STATS, OP_CWF_ROTATED, OP_DM_ROTATED, OP_CDK_NEG, OP_CD_STEP, OP_KL_TERMS, OP_SAMPLE = range(1, 12)

# name -> (restype, argtypes); must list every symbol declared in include/qcb.h
_P = c_void_p
SIGNATURES = {
    "qcb_version": (c_int, []),
    "qcb_last_error": (c_char_p, []),
    "qcb_launch_count": (c_uint64, []),
    "qcb_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int, c_int64]),
    "qcb_rbm_gibbs": (c_int, [_P, _P, _P, _P, _P, c_int, c_int, c_int, _P, c_int64, c_int, c_uint64, c_uint64, c_uint64,
                              _P, _P, c_size_t, _P]),
    "qcb_set_gibbs_mode": (c_int, [c_int]),
    "qcb_set_grad_reduce_mode": (c_int, [c_int]),
    "qcb_set_cd_epoch_mode": (c_int, [c_int]),
    "qcb_debug_gibbs_filter": (c_int, [_P, _P, _P, _P, _P, c_int, c_int, c_int, _P, c_int64, c_int, c_uint64, _P, _P, c_size_t, _P]),
    "qcb_linear_sigmoid": (c_int, [_P, c_int, c_int, c_int, _P, _P, _P, c_int, _P, c_int64, _P, _P]),
    "qcb_linear": (c_int, [_P, c_int, c_int, c_int, _P, _P, c_int64, _P, _P]),
    "qcb_bernoulli": (c_int, [_P, c_int64, _P, c_uint64, c_uint64, c_uint64, _P, _P]),
    "qcb_rbm_energy": (c_int, [_P, _P, _P, _P, _P, c_int, c_int, c_int, _P, c_int64, _P, _P, c_size_t, _P]),
    "qcb_rbm_grad_reduce": (c_int, [_P, _P, _P, _P, _P, c_int, c_int, c_int, _P, c_int64, c_double, _P, c_int, _P,
                                    c_size_t, _P]),
    "qcb_rbm_grad_batch": (c_int, [_P, _P, _P, _P, _P, c_int, c_int, c_int, _P, c_int64, _P, _P]),
    "qcb_rbm_flip_logratio": (c_int, [_P, _P, _P, _P, _P, c_int, c_int, c_int, _P, c_int64, _P, _P]),
    "qcb_rbm_cdk_neg": (c_int, [_P, _P, _P, _P, _P, c_int, c_int, c_int, _P, c_int64, c_int, c_uint64, c_uint64, c_uint64,
                                _P, c_double, _P, _P, c_size_t, _P]),
    "qcb_rbm_cd_step": (c_int, [_P, _P, _P, _P, _P, c_int, c_int, c_int, _P, c_int64, _P, c_int64, c_int, c_uint64, c_uint64,
                                c_uint64, c_double, c_double, _P, _P, c_size_t, _P]),
    "qcb_rbm_cd_step_packed": (c_int, [_P, _P, _P, _P, _P, c_int, c_int, c_int, _P, c_int64, _P, c_int64, c_int, c_uint64, c_uint64,
                                c_uint64, c_double, c_double, _P, _P, c_size_t, _P]),
    "qcb_rbm_cd_epoch": (c_int, [_P, _P, _P, c_int, c_int, _P, c_int64, c_int, _P, c_int64, c_int, c_int, c_int, c_double, c_uint64,
                                 c_uint64, _P, _P]),
    "qcb_rbm_logz_range": (c_int, [_P, _P, _P, _P, _P, c_int, c_int, c_int, c_uint64, c_uint64, _P, _P, c_size_t, _P]),
    "qcb_neg_logsumexp": (c_int, [_P, c_int64, _P, _P, c_size_t, _P]),
    "qcb_kron_rotate": (c_int, [_P, _P, c_int, c_int64, _P, _P, c_int, _P]),
    "qcb_kl_terms": (c_int, [_P, _P, _P, _P, c_int64, c_double, _P, _P, c_size_t, _P]),
    "qcb_rbm_exact_stats_range": (c_int, [_P, _P, _P, _P, _P, c_int, c_int, c_int, c_uint64, c_uint64, c_double, _P, _P,
                                          _P, _P, _P, c_size_t, _P]),
    "qcb_cwf_rotated_groups": (c_int, [_P, _P, _P, _P, _P, _P, c_int, c_int, _P, _P, _P, c_int, _P, c_int, _P, _P, _P, _P, c_size_t, _P]),
    "qcb_dm_rotated_groups": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, c_int, c_int, c_int, _P, _P, _P, c_int, _P, c_int, _P, _P,
                                      _P, _P, c_size_t, _P]),
    "qcb_fit_epoch_bases": (c_int, [_P] * 10 + [c_int, c_int, c_int, _P, _P, _P, c_int, c_int, _P, c_int, _P, c_int64, c_int, c_double,
                                               c_uint64, c_uint64, _P, _P, _P, _P, c_size_t, _P, _P, c_size_t, _P]),
    "qcb_fit_epoch_desc_bytes": (c_size_t, [c_int, c_int]),
    "qcb_cwf_rotated": (c_int, [_P, _P, _P, _P, _P, _P, c_int, c_int, _P, c_int64, _P, _P, c_int, _P, _P, _P, _P,
                                c_size_t, _P]),
    "qcb_dm_rotated": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, c_int, c_int, c_int, _P, c_int64, _P, _P, c_int,
                               _P, _P, _P, _P, c_size_t, _P]),
    "qcb_pack_states": (c_int, [_P, c_int64, c_int, _P, _P]),
    "qcb_unpack_states": (c_int, [_P, c_int64, c_int, _P, _P]),
    "qcb_rbm_sample": (c_int, [_P, _P, _P, _P, _P, c_int, c_int, c_int, c_int64, c_int, c_uint64, c_uint64, c_uint64, _P, _P, _P,
                               c_size_t, _P]),
    "qcb_rbm_gibbs_packed": (c_int, [_P, _P, _P, _P, _P, c_int, c_int, c_int, _P, c_int64, c_int, c_uint64, c_uint64, c_uint64, _P,
                                     c_size_t, _P]),
    "qcb_gibbs_host_packed_scratch_bytes": (c_size_t, [c_int, c_int, c_int, c_int64]),
    "qcb_rbm_gibbs_host_packed": (c_int, [_P, _P, _P, _P, _P, c_int, c_int, c_int, _P, c_int64, c_int, c_uint64, c_uint64,
                                          c_uint64, c_int, _P, c_size_t, _P]),
    "qcb_gibbs_host_scratch_bytes": (c_size_t, [c_int, c_int, c_int, c_int64]),
    "qcb_rbm_gibbs_host": (c_int, [_P, _P, _P, _P, _P, c_int, c_int, c_int, _P, c_int64, c_int, c_uint64, c_uint64,
                                   c_uint64, _P, c_size_t, _P]),
}

_lib = None


def load():
    """Load the shared library (once). Raises RuntimeError when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m qucumber_b200.build` "
                "(qucumber_b200 has no CPU or PyTorch fallback)"
            )
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc, what):
    if rc != QCB_OK:
        msg = load().qcb_last_error()
        raise RuntimeError(f"{what} failed (code {rc}): {msg.decode() if msg else ''}")


def ptr(t):
    """Device/host pointer of a contiguous float64/uint8 tensor, or None."""
    if t is None:
        return None
    if not t.is_contiguous():
        raise ValueError("qucumber_b200 kernels need contiguous tensors")
    return c_void_p(t.data_ptr())


def _raw_stream(device):
    """Raw cudaStream_t (int) of torch's current stream on ``device`` — the cheap C accessor, since the
    training loop asks several times per step."""
    idx = device.index if isinstance(device, torch.device) else torch.device(device).index
    if idx is None:
        idx = torch.cuda.current_device()
    return torch._C._cuda_getCurrentRawStream(idx)


def stream_ptr(device):
    return c_void_p(_raw_stream(device))


_ws_bytes_cache = {}


def workspace_bytes(op, n, nh, na, rows):
    """qcb_workspace_bytes with a host-side cache (the training loop asks the same question every step)."""
    key = (op, n, nh, na, rows)
    val = _ws_bytes_cache.get(key)
    if val is None:
        val = _ws_bytes_cache[key] = load().qcb_workspace_bytes(op, n, nh, na, rows)
    return val


_workspaces = {}


def workspace(device, nbytes):
    """A per-(device, stream) scratch tensor of at least nbytes (grown geometrically, owned by torch)."""
    key = (device.index if isinstance(device, torch.device) else torch.device(device).index, _raw_stream(device))
    buf = _workspaces.get(key)
    nbytes = max(int(nbytes), 256)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(nbytes, 2 * (buf.numel() if buf is not None else 0)), dtype=torch.uint8, device=device)
        _workspaces[key] = buf
    return buf


def require_cuda(t, what):
    if not t.is_cuda:
        raise RuntimeError(f"{what}: tensor is on {t.device}; qucumber_b200 runs on CUDA (sm_100a) only, there is no CPU path")
