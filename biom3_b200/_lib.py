"""ctypes binding of libbiom3_b200.so (include/biom3_b200.h).  No torch types cross this boundary:
only raw device pointers, sizes and a cudaStream_t."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# BIOM3_LIB: an alternative build of the same library (A/B of compile-time options, see build.py --variant); still CUDA only
LIB_PATH = os.environ.get('BIOM3_LIB') or os.path.join(HERE, 'libbiom3_b200.so')

# every symbol include/biom3_b200.h declares
SYMBOLS = [
    'biom3_create', 'biom3_destroy', 'biom3_last_error', 'biom3_set_weight', 'biom3_finalize_weights',
    'biom3_forward', 'biom3_decode', 'biom3_sample_all', 'biom3_unmask', 'biom3_gemm_test',
    'biom3_profile_step', 'biom3_launches_per_step', 'biom3_debug_copy', 'biom3_facilitator', 'biom3_attention_test',
    'biom3_set_precision', 'biom3_random_paths', 'biom3_debug_noise', 'biom3_debug_trace',
    'biom3_facilitator_create', 'biom3_facilitator_forward', 'biom3_facilitator_destroy', 'biom3_input_errors',
]

# BIOM3_DTYPE_* of include/biom3_b200.h
DTYPE_F32, DTYPE_BF16, DTYPE_F16, DTYPE_F64 = 0, 1, 2, 3


class Config(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        'seq_len', 'dim', 'heads', 'depth', 'n_blocks', 'local_heads', 'local_window', 'num_classes',
        'text_emb_dim', 'reversible')]


class StepProfile(C.Structure):
    _fields_ = [(n, C.c_float) for n in (
        'total_ms', 'gemm_qkv_ms', 'gemm_out_ms', 'gemm_ff1_ms', 'gemm_ff2_ms', 'local_attn_ms',
        'linear_attn_ms', 'layernorm_ms', 'embed_ms', 'head_ms', 'other_ms')] + [('launches', C.c_int32), ('compact_rows', C.c_int32)]


_lib = None


def load() -> C.CDLL:
    """Load the CUDA library.  There is no fallback: a missing library is an error."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f'{LIB_PATH} is missing: build it with `python -m biom3_b200.build` '
            '(nvcc, sm_100a). biom3_b200 has no CPU or PyTorch fallback.')
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64, u64 = C.c_void_p, C.c_int, C.c_int64, C.c_uint64
    lib.biom3_create.argtypes = [C.POINTER(Config), i32, i32, C.POINTER(vp)]
    lib.biom3_create.restype = i32
    lib.biom3_destroy.argtypes = [vp]
    lib.biom3_destroy.restype = None
    lib.biom3_last_error.argtypes = []
    lib.biom3_last_error.restype = C.c_char_p
    lib.biom3_set_weight.argtypes = [vp, C.c_char_p, vp, i32, C.POINTER(i64), i32]
    lib.biom3_set_weight.restype = i32
    lib.biom3_finalize_weights.argtypes = [vp]
    lib.biom3_finalize_weights.restype = i32
    lib.biom3_forward.argtypes = [vp, vp, vp, vp, i32, vp, vp]
    lib.biom3_forward.restype = i32
    lib.biom3_decode.argtypes = [vp, vp, vp, vp, i32, i32, i32, vp, u64, vp, vp, vp, i32, vp]
    lib.biom3_decode.restype = i32
    lib.biom3_sample_all.argtypes = [vp, vp, vp, i32, i32, i32, vp]
    lib.biom3_sample_all.restype = i32
    lib.biom3_unmask.argtypes = [vp, vp, vp, i32, i32, i32, i32, vp]
    lib.biom3_unmask.restype = i32
    lib.biom3_gemm_test.argtypes = [vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, vp]
    lib.biom3_gemm_test.restype = i32
    lib.biom3_profile_step.argtypes = [vp, i32, i32, C.POINTER(StepProfile)]
    lib.biom3_profile_step.restype = i32
    lib.biom3_debug_copy.argtypes = [vp, C.c_char_p, vp, i64]
    lib.biom3_debug_copy.restype = i32
    lib.biom3_facilitator.argtypes = [vp, i32, i32, i32, i32, vp, C.c_float, vp, vp, C.c_float, vp, vp, vp]
    lib.biom3_facilitator.restype = i32
    lib.biom3_attention_test.argtypes = [vp, vp, i32, i32, i32, i32, i32, vp]
    lib.biom3_attention_test.restype = i32
    lib.biom3_set_precision.argtypes = [vp, i32]
    lib.biom3_set_precision.restype = i32
    lib.biom3_random_paths.argtypes = [C.c_uint64, i32, i32, vp, vp]
    lib.biom3_random_paths.restype = i32
    lib.biom3_debug_noise.argtypes = [u64, i32, i32, i32, i32, vp, vp]
    lib.biom3_debug_noise.restype = i32
    lib.biom3_debug_trace.argtypes = [i32, vp, i64]
    lib.biom3_debug_trace.restype = i32
    lib.biom3_facilitator_create.argtypes = [i32, i32, i32, vp, C.c_float, vp, vp, C.c_float, vp, i32, C.POINTER(vp)]
    lib.biom3_facilitator_create.restype = i32
    lib.biom3_facilitator_forward.argtypes = [vp, vp, i32, vp, vp]
    lib.biom3_facilitator_forward.restype = i32
    lib.biom3_facilitator_destroy.argtypes = [vp]
    lib.biom3_facilitator_destroy.restype = None
    lib.biom3_input_errors.argtypes = [vp, C.POINTER(C.c_int32)]
    lib.biom3_input_errors.restype = i32
    lib.biom3_launches_per_step.argtypes = [vp]
    lib.biom3_launches_per_step.restype = i32
    _lib = lib
    return lib


def check(code: int) -> None:
    if code != 0:
        msg = load().biom3_last_error()
        raise RuntimeError(f'biom3_b200 error {code}: {msg.decode() if msg else "?"}')
