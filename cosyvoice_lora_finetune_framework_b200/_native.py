"""ctypes binding of libcvflow.so (the C ABI declared in include/cvflow.h).

There is no CPU fallback: if the library is missing or a call fails this raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# CVFLOW_LIB_PATH: load another build of the same ABI (A/B timing of kernel changes on one box)
LIB_PATH = os.environ.get("CVFLOW_LIB_PATH") or os.path.join(_HERE, "libcvflow.so")

DTYPE_F16 = 0
DTYPE_BF16 = 1

ACT_NONE, ACT_GELU_TANH, ACT_GELU_ERF, ACT_MUL_GELU_TANH_GRAD, ACT_MUL_GELU_ERF_GRAD = range(5)


class GemmSeg(C.Structure):
    _fields_ = [("a_map", C.c_int32), ("row_shift", C.c_int32), ("a_col0", C.c_int32),
                ("nkb", C.c_int32)]


class GemmDesc(C.Structure):
    _fields_ = [
        ("A", C.c_void_p * 2), ("a_rows", C.c_int32 * 2), ("a_cols", C.c_int32 * 2),
        ("a_ld", C.c_int64 * 2), ("a_bstride", C.c_int64 * 2), ("nbatch", C.c_int32),
        ("dtype", C.c_int32), ("W", C.c_void_p), ("N", C.c_int32), ("Ktot", C.c_int32),
        ("seg", GemmSeg * 8), ("nseg", C.c_int32), ("R", C.c_int32), ("rmul", C.c_int32),
        ("roff", C.c_int32), ("out_rows", C.c_int32), ("out", C.c_void_p), ("out_f32", C.c_int32),
        ("transposed_out", C.c_int32), ("ldc", C.c_int64), ("col_off", C.c_int32),
        ("n_valid", C.c_int32), ("alpha", C.c_float), ("act", C.c_int32), ("bias", C.c_void_p),
        ("aux_out", C.c_void_p), ("mul_src", C.c_void_p), ("ld_aux", C.c_int64),
        ("rowmask", C.c_void_p), ("resid", C.c_void_p), ("ldr", C.c_int64), ("dbg", C.c_void_p),
        ("ln_gamma", C.c_void_p), ("ln_beta", C.c_void_p), ("gn_part", C.c_void_p),
    ]


_lib = None


def lib():
    """Load libcvflow.so once; fail loudly when it is absent (no fallback path exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "libcvflow.so not found at %s - build it with "
                "`python -m cosyvoice_lora_finetune_framework_b200.build` (there is no CPU or "
                "PyTorch fallback for the flow hot path)" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        L.cvflow_last_error.restype = C.c_char_p
        L.cvflow_abi_version.restype = C.c_int
        L.cvflow_gemm.argtypes = [C.POINTER(GemmDesc), C.c_void_p]
        L.cvflow_gemm.restype = C.c_int
        _lib = L
    return _lib


def check(rc, what="cvflow"):
    if rc != 0:
        raise RuntimeError("%s failed (%d): %s" % (what, rc, lib().cvflow_last_error().decode()))


def current_stream():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def dtype_code(torch_dtype):
    import torch
    if torch_dtype == torch.float16:
        return DTYPE_F16
    if torch_dtype == torch.bfloat16:
        return DTYPE_BF16
    raise ValueError("16-bit operand dtype must be float16 or bfloat16, got %s" % torch_dtype)
