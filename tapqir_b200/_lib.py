"""
ctypes binding of ``lib/libtapqir_b200.so`` (C ABI declared in ``include/tapqir_b200.h``).

There is NO CPU fallback: if the shared library is missing or cannot be loaded, every compute
entry point raises :class:`NativeLibraryError`.  Tensors are passed as raw device pointers, work is
enqueued on torch's current CUDA stream.
"""

import ctypes
from ctypes import POINTER, Structure, c_char_p, c_double, c_int, c_int32, c_int64, c_uint8, c_uint64, c_void_p
from pathlib import Path

import torch

from tapqir_b200.exceptions import CudaOutOfMemoryError, NativeLibraryError

import os

# TQ_LIB: an experimental build of the same C ABI (csrc/build.py --variant=...) for A/B timings
LIB_PATH = Path(os.environ["TQ_LIB"]) if os.environ.get("TQ_LIB") else Path(__file__).resolve().parent / "lib" / "libtapqir_b200.so"

TQ_F32, TQ_F64 = 0, 1
TQ_PIX_U16, TQ_PIX_F32, TQ_PIX_F64 = 0, 1, 2
TQ_OK, TQ_ERR_ARG, TQ_ERR_CUDA, TQ_ERR_OOM, TQ_ERR_UNSUPPORTED = 0, 1, 2, 3, 4
K, M = 2, 4


class PatchView(Structure):
    """``tq_patch_view`` of include/tapqir_b200.h."""

    _fields_ = [
        ("nb", c_int32), ("fb", c_int32), ("C", c_int32), ("F", c_int32), ("P", c_int32), ("O", c_int32),
        ("pixtype", c_int32),
        ("ndx", c_void_p), ("fdx", c_void_p), ("pixels", c_void_p), ("xy", c_void_p),
        ("is_ontarget", c_void_p), ("mask", c_void_p), ("offset_samples", c_void_p), ("offset_logits", c_void_p),
    ]


# name -> (restype, argtypes); kept in one table so tests can check every header symbol is exported
_VP = c_void_p
SIGNATURES = {
    "tq_version": (c_int, []),
    "tq_last_error": (c_char_p, []),
    "tq_gaussian_spots": (c_int, [c_int, c_int64, c_int, c_int, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP]),
    "tq_ksmogn_fwd": (c_int, [c_int, POINTER(PatchView), _VP, _VP, _VP, _VP, _VP, _VP, _VP, c_int, _VP, _VP]),
    "tq_ksmogn_fwd_bwd": (c_int, [c_int, POINTER(PatchView), _VP, _VP, _VP, _VP, _VP, _VP, _VP, c_int, _VP,
                                   _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP]),
    # ---- SVI step (csrc/cosmos_step.cu) ----
    "tq_sizeof_tables": (c_int, []),
    "tq_sizeof_gstate": (c_int, []),
    "tq_sizeof_model_const": (c_int, []),
    "tq_local_post_scratch": (c_int64, [c_int, c_int, c_int]),
    "tq_local_post_tickets": (c_int64, [c_int, c_int, c_int]),
    "tq_cosmos_globals_sample": (c_int, [c_int, c_int, _VP, _VP, _VP, c_uint64, _VP, _VP, _VP, _VP, _VP]),
    "tq_site_record_rows": (c_int, []),
    "tq_cosmos_sites": (c_int, [c_int, POINTER(PatchView), c_int64, _VP, _VP, c_int64, c_uint64, _VP, _VP, _VP, _VP,
                                 _VP, _VP]),
    "tq_cosmos_sites_ws": (c_int, [c_int, POINTER(PatchView), c_int64, _VP, _VP, c_int64, c_uint64, _VP, _VP, _VP, _VP,
                                 _VP, _VP, _VP, _VP]),
    "tq_cosmos_sites_adam": (c_int, [c_int, POINTER(PatchView), c_int64, _VP, _VP, c_int64, c_uint64, _VP, _VP, _VP, _VP,
                                      _VP, _VP, _VP, _VP, _VP, _VP, c_double, c_double, c_double, _VP]),
    "tq_local_deferred_range": (c_int, [c_int64, c_int64, c_int64, POINTER(c_int64), POINTER(c_int64)]),
    "tq_cosmos_local_post": (c_int, [c_int, POINTER(PatchView), c_int64, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP,
                                      c_double, c_double, _VP, _VP, _VP, _VP, _VP]),
    "tq_cosmos_fused_supported": (c_int, [c_int, POINTER(PatchView)]),
    "tq_cosmos_fused_scratch": (c_int64, [c_int, c_int, c_int]),
    "tq_cosmos_fused_tickets": (c_int64, [c_int, c_int, c_int]),
    "tq_cosmos_fused_step": (c_int, [c_int, POINTER(PatchView), c_int64, _VP, _VP, _VP, _VP, c_int64, c_uint64, _VP, _VP,
                                      c_double, c_double, _VP, _VP, _VP, _VP, _VP, _VP, _VP]),
    "tq_cosmos_zprobs": (c_int, [c_int, POINTER(PatchView), c_int64, _VP, _VP, _VP, _VP, c_double, _VP, _VP, _VP]),
    "tq_cosmos_globals_grad": (c_int, [c_int, c_int, _VP, _VP, _VP, _VP, c_double, c_double, _VP, _VP, _VP, _VP]),
    "tq_sizeof_gprep": (c_int, []),
    "tq_cosmos_globals_prepare": (c_int, [c_int, c_int, _VP, _VP, _VP, _VP, _VP]),
    "tq_cosmos_globals_finish": (c_int, [c_int, c_int, _VP, _VP, _VP, _VP, c_double, c_double, _VP, _VP, _VP]),
    "tq_hmm_local_numel": (c_int64, [c_int64, c_int64, c_int64]),
    "tq_hmm_chain_sums": (c_int, []),
    "tq_hmm_globals_sample": (c_int, [c_int, c_int, _VP, _VP, _VP, c_uint64, _VP, _VP, _VP, _VP, _VP]),
    "tq_hmm_globals_prepare": (c_int, [c_int, c_int, _VP, _VP, _VP, _VP, _VP]),
    "tq_hmm_globals_finish": (c_int, [c_int, c_int, _VP, _VP, _VP, _VP, _VP, c_double, _VP, _VP, _VP]),
    "tq_hmm_chain_rows": (c_int, []),
    "tq_hmm_forward": (c_int, [c_int, POINTER(PatchView), c_int64, _VP, _VP, _VP, _VP, _VP, _VP, _VP]),
    "tq_hmm_local_post": (c_int, [c_int, POINTER(PatchView), c_int64, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP, c_double,
                                  _VP, _VP, _VP, _VP, _VP, _VP]),
    "tq_hmm_theta_probs": (c_int, [c_int, POINTER(PatchView), c_int64, _VP, _VP, _VP, _VP, _VP, c_double, _VP, _VP]),
    "tq_hmm_backward": (c_int, [c_int, POINTER(PatchView), c_int64, _VP, _VP, _VP, _VP, _VP, _VP, c_double, _VP, _VP, _VP, _VP]),
    "tq_p2p_bytes": (c_int64, []),
    "tq_p2p_max_values": (c_int, []),
    "tq_p2p_max_ranks": (c_int, []),
    "tq_p2p_alloc": (c_int, [POINTER(c_void_p), _VP]),
    "tq_p2p_open": (c_int, [_VP, POINTER(c_void_p)]),
    "tq_p2p_close": (c_int, [_VP]),
    "tq_p2p_free": (c_int, [_VP]),
    "tq_p2p_push": (c_int, [_VP, c_int, c_int, c_int, _VP, _VP]),
    "tq_p2p_wait_sum": (c_int, [_VP, c_int, c_int, _VP, _VP]),
    "tq_p2p_timed_out": (c_int, [_VP, POINTER(c_uint64)]),
    "tq_crop_aois": (c_int, [_VP, c_int, c_int, c_int, c_int, _VP, _VP, c_int, c_int, c_int, _VP, _VP, _VP, _VP]),
    "tq_offset_hist": (c_int, [_VP, c_int, c_int, c_int, c_int, c_int, c_int, _VP, _VP]),
    "tq_gamma_interval": (c_int, [c_int64, _VP, _VP, c_double, _VP, _VP, _VP]),
    "tq_beta_interval": (c_int, [c_int64, _VP, _VP, c_double, _VP, _VP, _VP]),
    "tq_snr_chi2": (c_int, [c_int64, c_int, c_int, _VP, _VP, _VP, _VP, _VP, _VP, _VP, c_double, c_double, c_double, _VP, _VP, _VP]),
    "tq_adam_dense": (c_int, [c_int, c_int64, _VP, _VP, _VP, _VP, c_double, c_double, c_double, c_double, _VP, _VP]),
    "tq_step_advance": (c_int, [_VP, _VP]),
    "tq_sizeof_step_state": (c_int, []),
    "tq_step_advance_deferred": (c_int, [_VP, c_double, c_double, c_double, _VP]),
    "tq_adam_deferred_flush": (c_int, [c_int, c_int64, _VP, _VP, _VP, _VP, c_double, c_double, c_double, _VP, _VP]),
    "tq_peak_fma": (c_int, [c_int, c_int, _VP, POINTER(c_double), _VP]),
    "tq_peak_mufu": (c_int, [c_int, c_int, _VP, POINTER(c_double), _VP]),
    "tq_subsample": (c_int, [c_int, c_int, c_uint64, _VP, c_uint64, _VP, _VP, _VP]),
    "tq_subsample_pair_supported": (c_int, [c_int, c_int]),
    "tq_subsample_pair": (c_int, [c_int, c_int, c_uint64, _VP, c_int, c_int, c_uint64, _VP, c_uint64, _VP, _VP]),
}

_lib = None


def load(required=True):
    """Load (once) and return the ctypes handle; raise NativeLibraryError when unavailable."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        if required:
            raise NativeLibraryError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). tapqir_b200 has no CPU fallback."
            )
        return None
    try:
        lib = ctypes.CDLL(str(LIB_PATH))
    except OSError as err:
        raise NativeLibraryError(f"cannot load {LIB_PATH}: {err}")
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def check(status, what=""):
    if status == TQ_OK:
        return
    msg = load().tq_last_error().decode("utf-8", "replace")
    if status == TQ_ERR_OOM:
        raise CudaOutOfMemoryError()
    if status == TQ_ERR_ARG:
        raise ValueError(f"{what}: {msg}")
    raise RuntimeError(f"{what}: {msg}")


def dtype_code(dtype):
    if dtype == torch.float32:
        return TQ_F32
    if dtype == torch.float64:
        return TQ_F64
    raise ValueError(f"unsupported dtype {dtype}: the kernels are instantiated for float32 and float64")


def pix_code(dtype):
    try:
        return {torch.uint16: TQ_PIX_U16, torch.float32: TQ_PIX_F32, torch.float64: TQ_PIX_F64}[dtype]
    except KeyError:
        raise ValueError(f"unsupported pixel dtype {dtype}")


def ptr(t):
    """Device pointer of a contiguous CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise ValueError("tapqir_b200 kernels take CUDA tensors only (no CPU path)")
    if not t.is_contiguous():
        raise ValueError("tensor must be contiguous")
    return t.data_ptr()


def stream_ptr(device=None):
    return torch.cuda.current_stream(device).cuda_stream


def make_view(pixels, xy, offset_samples, offset_logits, nb, fb, C, F, P, ndx=None, fdx=None,
              is_ontarget=None, mask=None):
    """Fill a PatchView; the caller must keep the tensors alive while kernels use them."""
    v = PatchView()
    v.nb, v.fb, v.C, v.F, v.P, v.O = int(nb), int(fb), int(C), int(F), int(P), int(offset_samples.numel())
    v.pixtype = pix_code(pixels.dtype)
    for idx in (ndx, fdx):
        if idx is not None and idx.dtype != torch.int32:
            raise ValueError("minibatch indices must be int32")
    v.ndx, v.fdx = ptr(ndx), ptr(fdx)
    v.pixels, v.xy = ptr(pixels), ptr(xy)
    v.is_ontarget, v.mask = ptr(is_ontarget), ptr(mask)
    v.offset_samples, v.offset_logits = ptr(offset_samples), ptr(offset_logits)
    return v
