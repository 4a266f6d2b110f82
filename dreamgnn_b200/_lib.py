"""ctypes binding of libdreamgnn.so (the C ABI declared in include/dreamgnn.h).

There is no CPU fallback: if the shared library has not been built, or a tensor is not on a CUDA
device, the call raises. Build with `python -m dreamgnn_b200.build`.
"""
import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_double, c_float, c_int, c_int64, c_size_t, c_uint64, c_ulonglong, c_void_p

import torch as th

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, 'lib', 'libdreamgnn.so')
ABI_VERSION = 5

_P = c_void_p


class AdamTensor(Structure):
    """dg_adam_tensor_t (include/dreamgnn.h): one parameter with its gradient and Adam moments."""
    _fields_ = [('param', c_void_p), ('grad', c_void_p), ('exp_avg', c_void_p), ('exp_avg_sq', c_void_p), ('numel', c_int64)]


class CopyItem(Structure):
    """dg_copy_t (include/dreamgnn.h)."""
    _fields_ = [('dst', c_void_p), ('src', c_void_p), ('bytes', c_int64)]


_SIGNATURES = {
    'dg_abi_version': (c_int, []),
    'dg_last_error': (c_char_p, []),
    'dg_launch_count': (c_ulonglong, []),
    'dg_reset_launch_count': (None, []),
    'dg_scan_workspace_bytes': (c_size_t, [c_int64]),
    'dg_exclusive_scan_i32': (c_int, [_P, _P, c_int64, _P, c_size_t, _P]),
    'dg_sort_workspace_bytes': (c_size_t, [c_int64]),
    'dg_sort_pairs_u64': (c_int, [_P, _P, _P, _P, c_int64, c_int, _P, c_size_t, _P]),
    'dg_csr_build_workspace_bytes': (c_size_t, [c_int64, c_int64]),
    'dg_csr_build': (c_int, [_P, _P, c_int64, c_int64, c_int64, _P, _P, _P, _P, c_size_t, _P]),
    'dg_degree_norm': (c_int, [_P, c_int64, _P, _P]),
    'dg_keep_flags_from_perm': (c_int, [_P, c_int64, c_int64, _P, _P]),
    'dg_random_subset_workspace_bytes': (c_size_t, []),
    'dg_random_subset_flags': (c_int, [_P, c_int64, c_int64, _P, _P, c_size_t, _P]),
    'dg_csr_compact_workspace_bytes': (c_size_t, [c_int64]),
    'dg_csr_compact': (c_int, [_P, _P, _P, _P, c_int64, _P, _P, _P, _P, _P, _P, c_size_t, _P]),
    'dg_csr_expand_rows': (c_int, [_P, c_int64, _P, _P]),
    'dg_spmm_csr_f32': (c_int, [_P, _P, _P, _P, _P, _P, _P, c_int64, _P, c_int64, c_int64, c_int64, c_int, _P]),
    'dg_spmm_csr_bf16': (c_int, [_P, _P, _P, _P, _P, _P, _P, c_int64, _P, c_int64, c_int64, c_int64, c_int, _P]),
    'dg_decoder_fwd_f32': (c_int, [_P, _P, _P, c_int64, _P, _P, _P, _P, _P, _P, c_float, c_uint64, _P, _P, _P, _P]),
    'dg_decoder_bwd_workspace_bytes': (c_size_t, [c_int64]),
    'dg_decoder_bwd_f32': (c_int, [_P, _P, _P, c_int64, _P, _P, _P, _P, c_float, c_uint64, _P, _P, _P, _P, _P, _P, _P, _P,
                                   _P, _P, _P, c_size_t, _P]),
    'dg_gemm_nt_workspace_bytes': (c_size_t, [c_int64, c_int64, c_int64, c_int64, c_int, c_int]),
    'dg_gemm_nt_f32': (c_int, [_P, c_int64, c_int64, _P, c_int64, c_int64, _P, c_int64, c_int64, c_int64, c_int64, c_int64,
                               c_int64, _P, c_int, _P, c_size_t, _P]),
    'dg_gemm_f32': (c_int, [_P, c_int64, c_int64, c_int, _P, c_int64, c_int64, c_int, _P, c_int64, c_int64, c_int64, c_int64,
                            c_int64, c_int64, _P, c_int, _P, c_size_t, _P]),
    'dg_small_gemm_workspace_bytes': (c_size_t, [c_int64, c_int64, c_int64, c_int64]),
    'dg_small_gemm_tickets': (c_int64, [c_int64, c_int64, c_int64, c_int64]),
    'dg_small_gemm_f32': (c_int, [_P, c_int64, c_int64, c_int, _P, c_int64, c_int64, c_int, _P, _P, c_int64, c_int64, c_int64,
                                  c_int64, c_int64, c_int64, c_int, _P, c_size_t, _P, _P]),
    'dg_colsum_workspace_bytes': (c_size_t, [c_int64, c_int64]),
    'dg_colsum_f32': (c_int, [_P, c_int64, _P, c_int64, _P, c_int64, c_int64, c_int64, _P, _P, c_size_t, _P, _P]),
    'dg_center_normalize_f64': (c_int, [_P, c_int64, _P, c_int64, c_int64, c_double, _P, c_int64, _P, _P]),
    'dg_center_normalize_bwd_f64': (c_int, [_P, c_int64, _P, c_int64, _P, c_int64, c_int64, _P, c_int64, _P, c_double, _P]),
    'dg_topk_rows_f64': (c_int, [_P, c_int64, c_int64, c_int64, c_int, _P, _P]),
    'dg_knn_graph_workspace_bytes': (c_size_t, [c_int64, c_int]),
    'dg_knn_graph_from_neighbors': (c_int, [_P, c_int64, c_int, _P, _P, _P, _P, _P, _P, c_size_t, _P]),
    'dg_act_dropout_f32': (c_int, [_P, c_int64, _P, c_int64, _P, c_int64, c_int64, c_int64, c_int, c_float, c_float, c_uint64, _P, _P]),
    'dg_attention_fwd_f32': (c_int, [_P, c_int64, _P, c_int64, c_int64, c_int64, _P, _P, _P, c_int, c_float, c_uint64, _P, _P,
                                     c_int64, _P, _P]),
    'dg_attention_bwd_workspace_bytes': (c_size_t, [c_int64, c_int64]),
    'dg_attention_bwd_f32': (c_int, [_P, c_int64, _P, c_int64, c_int64, c_int64, _P, _P, _P, c_int, c_float, c_uint64, _P, _P,
                                     c_int64, _P, _P, c_int64, _P, c_int64, _P, _P, c_size_t, _P]),
    'dg_bce_logits_workspace_bytes': (c_size_t, [c_int64]),
    'dg_bce_logits_fwd_f32': (c_int, [_P, _P, c_int64, c_float, _P, _P, c_size_t, _P]),
    'dg_bce_logits_bwd_f32': (c_int, [_P, _P, c_int64, c_float, _P, _P, _P]),
    'dg_gram_common_loss_f64': (c_int, [_P, c_int64, c_int64, c_double, _P, _P, _P]),
    'dg_basis_combine_fwd_f32': (c_int, [_P, _P, c_int, c_int, c_int64, c_int64, c_int64, _P, _P]),
    'dg_basis_combine_bwd_workspace_bytes': (c_size_t, [c_int64, c_int64]),
    'dg_basis_combine_bwd_f32': (c_int, [_P, _P, _P, c_int, c_int, c_int64, c_int64, c_int64, _P, _P, _P, c_size_t, _P]),
    'dg_adam_workspace_bytes': (c_size_t, [POINTER(AdamTensor), c_int]),
    'dg_adam_clip_step_f32': (c_int, [POINTER(AdamTensor), c_int, _P, _P, c_double, c_double, c_double, c_double, c_double, c_double,
                                      _P, _P, c_size_t, _P]),
    'dg_multi_copy': (c_int, [POINTER(CopyItem), c_int, _P]),
    'dg_bench_read_rows': (c_int, [_P, c_int64, c_int64, c_int64, c_int, c_int, c_int, _P, _P]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)
_lib = None


def load():
    """Load libdreamgnn.so (once). Raises RuntimeError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise RuntimeError('libdreamgnn.so not found at %s -- build it with `python -m dreamgnn_b200.build` '
                           '(there is no CPU fallback)' % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)            # AttributeError if the .so does not export a declared symbol
        fn.restype, fn.argtypes = res, args
    if lib.dg_abi_version() != ABI_VERSION:
        raise RuntimeError('libdreamgnn.so ABI %d != expected %d: rebuild' % (lib.dg_abi_version(), ABI_VERSION))
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        msg = load().dg_last_error().decode('utf-8', 'replace')
        raise RuntimeError('%s failed (rc=%d): %s' % (what, rc, msg))


def ptr(t, dtype=None, name='tensor'):
    """Device pointer of a contiguous CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not isinstance(t, th.Tensor) or not t.is_cuda:
        raise RuntimeError('%s must be a CUDA tensor: dreamgnn_b200 has no CPU path' % name)
    if dtype is not None and t.dtype != dtype:
        raise TypeError('%s must be %s, got %s' % (name, dtype, t.dtype))
    if not t.is_contiguous():
        raise ValueError('%s must be contiguous' % name)
    return t.data_ptr()


def stream():
    return th.cuda.current_stream().cuda_stream


def workspace(nbytes, device):
    return th.empty(max(int(nbytes), 1), dtype=th.uint8, device=device)


def launch_count():
    return int(load().dg_launch_count())


def reset_launch_count():
    load().dg_reset_launch_count()
