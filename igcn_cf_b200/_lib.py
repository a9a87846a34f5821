"""ctypes binding of libigcn_b200.so (the C ABI declared in include/igcn_b200.h).

There is NO fallback: if the library is missing or a call fails, a RuntimeError is raised.
Tensors are passed as raw device pointers (`tensor.data_ptr()`), the stream is torch's
current CUDA stream, and PyTorch owns every buffer.
"""
import ctypes as C
import os

import torch

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, 'libigcn_b200.so')
ABI_VERSION = 2
MAX_ADD = 8
MAX_PEERS = 8

c_void_p, c_int32, c_int64, c_uint64, c_float = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.c_float


class CsrStruct(C.Structure):
    _fields_ = [('n_rows', c_int64), ('n_cols', c_int64), ('nnz', c_int64),
                ('rowptr', c_void_p), ('col', c_void_p), ('val', c_void_p),
                ('long_threshold', c_int32), ('n_chunks', c_int32),
                ('chunk_row', c_void_p), ('chunk_begin', c_void_p), ('chunk_len', c_void_p),
                ('chunk_first', c_void_p), ('chunk_count', c_void_p),
                ('partial', c_void_p), ('counters', c_void_p), ('row_order', c_void_p),
                ('n_long_rows', c_int32), ('n_medium_rows', c_int32)]


class DropoutStruct(C.Structure):
    _fields_ = [('mode', c_int32), ('p', c_float), ('seed', c_uint64), ('seed_dev', c_void_p),
                ('edge_keep', c_void_p), ('self_keep', c_void_p), ('tperm', c_void_p)]


# name -> argtypes; every function returns int (0 == ok) except the two accessors.
_SIGNATURES = {
    'igcn_spmm': [C.POINTER(CsrStruct), c_void_p, c_void_p, c_int32, C.POINTER(c_void_p), c_int32,
                  c_void_p, c_float, C.POINTER(c_void_p), c_int32, c_void_p],
    'igcn_inmo_fwd': [C.POINTER(CsrStruct), c_void_p, c_void_p, C.POINTER(DropoutStruct), c_void_p,
                      c_void_p, c_int32, c_int64, c_int64, c_int64, c_int64, C.POINTER(c_void_p), c_int32, c_void_p],
    'igcn_inmo_bwd': [C.POINTER(CsrStruct), c_void_p, C.POINTER(DropoutStruct), c_void_p, c_void_p,
                      c_int32, c_int64, C.POINTER(c_void_p), c_int32, c_void_p],
    'igcn_colsum_masked': [c_void_p, c_int64, c_int64, c_int32, C.POINTER(DropoutStruct), c_void_p,
                           c_void_p, c_void_p],
    'igcn_sample_triples': [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int64, c_uint64, c_uint64,
                            c_void_p, c_void_p, c_void_p],
    'igcn_bpr_fwd': [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int32, c_void_p,
                     c_void_p, c_void_p, c_void_p],
    'igcn_bpr_partial': [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int32, C.POINTER(c_void_p), c_int32, c_int32,
                         c_int32, c_int64, c_void_p, c_void_p],
    'igcn_bpr_combine': [c_void_p, c_int64, c_int64, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p],
    'igcn_loss_finalize': [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_float, c_float, c_void_p,
                           c_void_p, c_void_p],
    'igcn_bpr_plan': [c_void_p, c_int64, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p],
    'igcn_spmm_rows': [C.POINTER(CsrStruct), c_void_p, c_void_p, c_int32, C.POINTER(c_void_p), c_int32,
                       c_void_p, c_float, c_void_p, c_void_p, c_int64, c_int64, C.POINTER(c_void_p), c_int32, c_void_p],
    'igcn_spmm_cols': [C.POINTER(CsrStruct), c_void_p, c_void_p, c_int32, C.POINTER(c_void_p), c_int32,
                       c_void_p, c_float, c_void_p, C.POINTER(c_void_p), c_int32, c_void_p],
    'igcn_bpr_bwd': [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int32, c_void_p, c_float, c_float,
                     c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_void_p,
                     c_void_p, c_void_p],
    'igcn_bpr_dw': [c_void_p, c_void_p, c_int64, c_int64, c_int32, c_void_p, c_float, c_void_p, c_void_p, c_void_p],
    'igcn_l2_rows_bwd': [c_void_p, c_void_p, c_int32, c_float, c_void_p, c_void_p, c_void_p, c_int64,
                         c_void_p],
    'igcn_adam': [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_float, c_float, c_float, c_float,
                  c_int64, c_void_p, c_void_p],
    'igcn_step_tick': [c_void_p, c_float, c_float, c_float, c_void_p],
    'igcn_score_topk_exact': [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int32, c_void_p, c_void_p,
                              c_int64, c_int64, c_void_p, c_int32, c_void_p, c_void_p, c_void_p, c_void_p,
                              c_int32, c_void_p, c_int64, c_void_p],
    'igcn_tc_workspace': [c_int64, c_int64, c_int32, c_int32, C.POINTER(c_int64), C.POINTER(c_int64),
                          C.POINTER(c_int64)],
    'igcn_tc_pack': [c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int64, c_int32, c_void_p, c_void_p, c_void_p, c_void_p,
                     c_void_p, c_void_p, c_void_p],
    'igcn_tc_candidates': [c_void_p, c_void_p, c_int64, c_int64, c_int32, c_int32, c_int32, c_int64, c_int64, c_void_p,
                           c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p],
    'igcn_tc_finalize': [c_void_p, c_void_p, c_int64, c_int64, c_int32, c_int32, c_void_p, c_void_p, c_void_p,
                         c_void_p, c_void_p, c_int64, c_void_p, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p],
    'igcn_predict_scores': [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int32, c_void_p, c_void_p],
    'igcn_hits': [c_void_p, c_int64, c_int32, c_void_p, c_void_p, c_void_p, c_void_p],
    'igcn_user_metrics': [c_void_p, c_int64, c_int32, c_void_p, c_void_p, c_void_p, C.POINTER(c_int32), c_int32, c_void_p, c_void_p,
                          c_void_p],
    'igcn_peer_alloc': [c_int64, C.POINTER(c_void_p), c_void_p],
    'igcn_peer_open': [c_void_p, C.POINTER(c_void_p)],
    'igcn_peer_close': [c_void_p],
    'igcn_peer_free': [c_void_p],
    'igcn_peer_barrier': [C.POINTER(c_void_p), c_int32, c_int32, c_void_p, c_void_p, c_void_p],
    'igcn_peer_push': [C.POINTER(c_void_p), c_int32, c_int32, c_int64, c_int64, c_void_p],
    'igcn_peer_push_cols': [C.POINTER(c_void_p), c_int32, c_void_p, c_int64, c_int32, c_int32, c_int32, c_void_p],
}
EXPORTS = ['igcn_abi_version', 'igcn_last_error'] + sorted(_SIGNATURES)

_lib = None


def load():
    """Load the shared library once; raise if it is missing or has the wrong ABI."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError('igcn_cf_b200: %s not found -- build it with `python -m igcn_cf_b200.build` '
                           '(there is no CPU or PyTorch fallback)' % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    lib.igcn_abi_version.restype = C.c_int
    lib.igcn_last_error.restype = C.c_char_p
    if lib.igcn_abi_version() != ABI_VERSION:
        raise RuntimeError('igcn_cf_b200: ABI version mismatch, rebuild the library')
    for name, argtypes in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = C.c_int
    _lib = lib
    return lib


def stream_ptr():
    return torch.cuda.current_stream().cuda_stream


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    if t is None:
        return None
    return t.data_ptr()


# kernels launched per entry point (igcn_bpr_bwd launches 2 more when dw is requested)
KERNELS_PER_CALL = {'igcn_colsum_masked': 2, 'igcn_bpr_dw': 2, 'igcn_score_topk_exact': 3, 'igcn_tc_pack': 5, 'igcn_tc_workspace': 0, 'igcn_peer_alloc': 0,
                    'igcn_peer_open': 0, 'igcn_peer_close': 0, 'igcn_peer_free': 0}
launch_count = 0          # running total of kernel launches issued through this binding
profile_hook = None       # optional callable(name, phase, args) used by bench.py to time launches


def call(name, *args):
    global launch_count
    lib = load()
    launch_count += KERNELS_PER_CALL.get(name, 1) + (2 if name == 'igcn_bpr_bwd' and args[16] else 0)
    if profile_hook is not None:
        profile_hook(name, 0, args)
        rc = getattr(lib, name)(*args)
        profile_hook(name, 1, args)
    else:
        rc = getattr(lib, name)(*args)
    if rc != 0:
        raise RuntimeError('%s failed (%d): %s' % (name, rc, lib.igcn_last_error().decode()))


def require_cuda(t, dtype=None, name='tensor'):
    if not t.is_cuda:
        raise RuntimeError('igcn_cf_b200: %s must live on a CUDA device (no CPU path exists)' % name)
    if dtype is not None and t.dtype != dtype:
        raise RuntimeError('igcn_cf_b200: %s must be %s, got %s' % (name, dtype, t.dtype))
    if not t.is_contiguous():
        raise RuntimeError('igcn_cf_b200: %s must be contiguous' % name)
    return t
