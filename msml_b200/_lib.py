"""ctypes binding of libmsml_b200.so (the C ABI in include/msml_b200.h).

There is exactly one compute path: the sm_100a CUDA library.  If the library is missing, fails to
load, or a call returns non-zero, a RuntimeError is raised — nothing here ever falls back to
PyTorch ops or to the CPU.
"""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libmsml_b200.so")

F32, BF16, F16 = 0, 1, 2
ACT = {"tanh": 0, "sigmoid": 1}
ARITH = {"add": 0, "sub": 1, "div": 2, "mul": 3}
MARGIN = {"arc": 0, "cos": 1}

c_p = ctypes.c_void_p
c_i64 = ctypes.c_int64
c_int = ctypes.c_int
c_size = ctypes.c_size_t


class MarginParams(ctypes.Structure):
    _fields_ = [("kind", c_int), ("s", ctypes.c_float), ("m", ctypes.c_float),
                ("a", ctypes.c_float), ("k", ctypes.c_float)]


# name -> (restype, argtypes); every symbol include/msml_b200.h declares
SIGNATURES = {
    "msml_abi_version": (c_int, []),
    "msml_last_error": (ctypes.c_char_p, []),
    "msml_launch_count": (c_i64, []),
    "msml_launch_count_reset": (None, []),
    "msml_profile_enable": (None, [c_int]),
    "msml_profile_collect": (c_i64, [ctypes.c_char_p, c_i64]),
    "msml_fm_gate_fwd": (c_int, [c_p, c_p, c_p, c_p, c_i64, c_int, c_int, c_int, c_p]),
    "msml_fm_gate_bwd": (c_int, [c_p, c_p, c_p, c_p, c_p, c_i64, c_int, c_int, c_int, c_p]),
    "msml_fm_gate_fwd_multi": (c_int, [c_int, c_p, c_p, c_p, c_p, c_int, c_int, c_int, c_p]),
    "msml_fm_gate_bwd_multi": (c_int, [c_int, c_p, c_p, c_p, c_p, c_p, c_p, c_int, c_int, c_int, c_p]),
    "msml_fm_mask_fwd": (c_int, [c_p, c_p, c_p] + [c_i64] * 7 + [c_int, c_int, c_int, c_p]),
    "msml_fm_mask_bwd": (c_int, [c_p, c_p, c_p, c_p, c_p] + [c_i64] * 7 + [c_int, c_int, c_int, c_p]),
    "msml_fm_cat_fwd": (c_int, [c_p, c_p, c_p, c_i64, c_i64, c_i64, c_i64, c_int, c_p]),
    "msml_fm_cat_bwd": (c_int, [c_p, c_p, c_p, c_p, c_i64, c_i64, c_i64, c_i64, c_int, c_p]),
    "msml_consensus_workspace": (c_size, [c_i64, c_i64, c_i64, c_i64]),
    "msml_consensus_fwd": (c_int, [c_p, c_p, c_p, c_i64, c_i64, c_i64, c_i64, c_int, c_int, ctypes.c_float, ctypes.c_float,
                                   c_int, c_int, c_p, c_p, c_p, c_size, c_p]),
    "msml_consensus_bwd": (c_int, [c_p, c_p, c_p, c_p, c_p, c_i64, c_i64, c_i64, c_i64, c_int, c_int, c_p]),
    "msml_bn_workspace": (c_size, [c_i64, c_i64]),
    "msml_fm_peer_mul_fwd": (c_int, [c_p] * 5 + [c_i64, c_int, c_int, c_int, c_p]),
    "msml_fm_peer_mul_bwd": (c_int, [c_p] * 7 + [c_i64, c_int, c_int, c_int, c_p]),
    "msml_mse_workspace": (c_size, []),
    "msml_mse_fwd": (c_int, [c_p, c_p, c_i64, c_int, c_p, c_p, c_size, c_p]),
    "msml_mse_bwd": (c_int, [c_p] * 5 + [c_i64, c_int, c_p]),
    "msml_bn_fwd": (c_int, [c_p] * 11 + [c_i64, c_i64, c_int, c_int, ctypes.c_float, ctypes.c_float, c_p, c_size, c_p]),
    "msml_bn_fwd_ex": (c_int, [c_p] * 11 + [c_i64, c_i64, c_int, c_int, ctypes.c_float, ctypes.c_float, c_p, c_size, c_p, c_size,
                                c_int, c_p]),
    "msml_bn_bwd": (c_int, [c_p] * 14 + [c_i64, c_i64, c_int, c_int, c_int, c_p, c_size, c_p]),
    "msml_pfc_sgd_update": (c_int, [c_p, c_p, c_p, c_p, c_i64, c_i64, c_i64, c_p, ctypes.c_float, ctypes.c_float, ctypes.c_float,
                                    ctypes.c_float, c_int, c_p, c_p, c_p]),
    "msml_pfc_sgd_update_raw": (c_int, [c_p, c_p, c_p, c_p, c_i64, c_i64, c_i64, c_p, ctypes.c_float, ctypes.c_float, ctypes.c_float,
                                        ctypes.c_float, c_int, c_p, c_p, c_p]),
    "msml_sgd_flat": (c_int, [c_p, c_p, c_p, c_p, c_i64, c_p, c_p, ctypes.c_float, ctypes.c_float, c_int, c_p]),
    "msml_accum_bf16_multi": (c_int, [c_int, c_p, c_p, c_p, c_p]),
    "msml_dap_fwd": (c_int, [c_p, c_p, c_p, c_i64, c_i64, c_i64, c_i64, c_i64, c_int, c_int, c_p]),
    "msml_dap_bwd": (c_int, [c_p, c_p, c_i64, c_i64, c_i64, c_i64, c_i64, c_int, c_int, c_p]),
    "msml_pfc_remap": (c_int, [c_p, c_i64, c_i64, c_i64, c_p]),
    "msml_pfc_mark_positive": (c_int, [c_p, c_p, c_i64, c_i64, c_p]),
    "msml_pfc_select_workspace": (c_size, [c_i64]),
    "msml_pfc_select": (c_int, [c_p, c_i64, c_i64, c_p, c_p, c_p, c_size, c_p]),
    "msml_pfc_searchsorted": (c_int, [c_p, c_i64, c_p, c_p, c_p]),
    "msml_gather_rows_f32": (c_int, [c_p, c_p, c_p, c_i64, c_i64, c_p]),
    "msml_scatter_rows_f32": (c_int, [c_p, c_p, c_p, c_i64, c_i64, c_p]),
    "msml_wnorm_cast": (c_int, [c_p, c_p, c_p, c_i64, c_p, c_i64, c_i64, c_p]),
    "msml_cast_bf16": (c_int, [c_p, c_p, c_p, c_i64, c_i64, c_i64, c_p]),
    "msml_transpose_bf16": (c_int, [c_p, c_p, c_i64, c_i64, c_i64, c_p]),
    "msml_head_workspace": (c_size, [c_i64, c_i64, c_i64]),
    "msml_head_fwd": (c_int, [c_p, c_p, c_p, c_i64, c_i64, c_i64, ctypes.POINTER(MarginParams), c_p, c_p, c_size, c_p]),
    "msml_head_merge_stats": (c_int, [c_p, c_i64, c_i64, c_p, c_p, c_p]),
    "msml_head_bwd": (c_int, [c_p, c_p, c_p, c_p, c_i64, c_i64, c_i64,
                              ctypes.POINTER(MarginParams), c_p, c_p, c_p, c_p, c_size, c_p]),
    "msml_nccl_unique_id": (c_int, [c_p]),
    "msml_nccl_init": (c_int, [c_p, c_int, c_int, ctypes.POINTER(c_p)]),
    "msml_nccl_destroy": (c_int, [c_p]),
    "msml_comm_world": (c_int, [c_p]),
    "msml_comm_rank": (c_int, [c_p]),
    "msml_head_gather_workspace": (c_size, [c_i64, c_i64, c_i64]),
    "msml_head_gather": (c_int, [c_p, c_p, c_p, c_i64, c_i64, c_i64, c_i64, c_p, c_p, c_p, c_size, c_p]),
    "msml_head_step_workspace": (c_size, [c_i64, c_i64, c_i64, c_i64]),
    "msml_head_step": (c_int, [c_p, c_p, c_p, c_p, c_p, c_i64, c_i64, c_i64, ctypes.POINTER(MarginParams), c_p, c_p, c_p, c_p, c_size, c_p]),
    "msml_head_bwd_raw": (c_int, [c_p, c_p, c_p, c_i64, c_i64, c_i64, ctypes.POINTER(MarginParams), c_p, c_p, c_p, c_p, c_size, c_p]),
    "msml_head_step_raw": (c_int, [c_p, c_p, c_p, c_p, c_i64, c_i64, c_i64, ctypes.POINTER(MarginParams), c_p, c_p, c_p, c_p, c_size, c_p]),
    "msml_margin_fwd": (c_int, [c_p, c_p, c_i64, c_i64, c_i64, ctypes.POINTER(MarginParams), c_p]),
    "msml_margin_bwd": (c_int, [c_p, c_p, c_p, c_i64, c_i64, c_i64, ctypes.POINTER(MarginParams), c_p]),
    "msml_gemm_bf16": (c_int, [c_p, c_i64, c_int, c_p, c_i64, c_int, c_p, c_i64, c_i64, c_i64, c_i64, c_int, c_p]),
    "msml_gemm_bf16_pair": (c_int, [c_p, c_i64, c_int, c_p, c_i64, c_int, c_p, c_i64, c_i64, c_i64, c_i64, c_int, c_p]),
    "msml_gemm_bf16_tn": (c_int, [c_p, c_i64, c_p, c_i64, c_p, c_i64, c_i64, c_i64, c_i64, c_p]),
}

_lib = None


def load():
    """Load (once) and return the ctypes handle; raises if the library is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "msml_b200: %s is missing — build it with `python -m msml_b200._build` "
            "(there is no fallback path)" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)   # AttributeError if the library does not export the symbol
        fn.restype = res
        fn.argtypes = args
    if lib.msml_abi_version() != 1:
        raise RuntimeError("msml_b200: ABI version mismatch")
    _lib = lib
    return lib


def check(code):
    if code != 0:
        raise RuntimeError("msml_b200 error %d: %s" % (code, load().msml_last_error().decode()))


def dtype_code(t):
    import torch
    try:
        return {torch.float32: F32, torch.bfloat16: BF16, torch.float16: F16}[t]
    except KeyError:
        raise RuntimeError("msml_b200: unsupported dtype %s" % t)


def stream_ptr():
    import torch
    return torch.cuda.current_stream().cuda_stream


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("msml_b200: this operator has only a CUDA (sm_100a) implementation; "
                               "got a %s tensor" % t.device)


def margin_params(kind, s, m, a, k):
    return MarginParams(MARGIN[kind], float(s), float(m), float(a), float(k))
