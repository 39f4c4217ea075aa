"""ctypes binding of libjacobi_b200.so (include/jacobi_b200.h).  No fallback: a missing library raises ImportError
with the build command; a missing GPU makes jf_create fail."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libjacobi_b200.so")

EXPORTS = [
    "jf_create", "jf_destroy", "jf_last_error", "jf_device_info", "jf_load_tensor", "jf_finalize_weights",
    "jf_set_rope_tables", "jf_kv_info", "jf_forward", "jf_commit_rows", "jf_jacobi_run", "jf_launch_count",
    "jf_profile_enable", "jf_profile_read", "jf_tp_export", "jf_tp_import", "jf_tp_attach_local",
]

JF_OK, JF_ERR_INVALID, JF_ERR_CUDA, JF_ERR_CAPACITY, JF_ERR_UNSUPPORTED, JF_ERR_POLICY = 0, -1, -2, -3, -4, -5
DTYPE_BF16, DTYPE_F32, DTYPE_F16 = 0, 1, 2


class ModelCfg(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "hidden_size", "num_hidden_layers", "num_attention_heads", "num_key_value_heads", "head_dim",
        "intermediate_size", "vocab_size", "tie_word_embeddings", "max_position_embeddings")] + [
        ("rms_norm_eps", C.c_float), ("rope_theta", C.c_float)]


class RuntimeCfg(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "device", "max_rows", "max_seqs", "num_kv_pages", "max_pages_per_seq", "attn_chunk", "attn_slots",
        "gen_capacity", "max_calls", "trace_capacity", "tp_rank", "tp_size", "tp_attn", "num_ctas")]


class PolicyKnobsC(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "mode", "n", "K", "spawn_threshold", "pool_size", "eos", "pad", "max_iteration_count", "max_new_tokens",
        "max_calls", "lookahead_min_acc", "vocab", "temp_bits")] + [("seed_lo", C.c_uint32), ("seed_hi", C.c_uint32),
                                                                     ("rng_offset", C.c_int32)]


_lib = None


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("JF_LIB_PATH", LIB_PATH)           # developer aid: A/B an older build of the library
    if not os.path.exists(path):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc -gencode arch=compute_100a,code=sm_100a). There is no CPU fallback.")
    lib = C.CDLL(path)
    vp, i32, i64p, i32p, fp = C.c_void_p, C.c_int32, C.POINTER(C.c_int64), C.POINTER(C.c_int32), C.POINTER(C.c_float)
    lib.jf_create.argtypes = [C.POINTER(ModelCfg), C.POINTER(RuntimeCfg), C.POINTER(vp)]
    lib.jf_destroy.argtypes = [vp]
    lib.jf_last_error.argtypes = [vp]
    lib.jf_last_error.restype = C.c_char_p
    lib.jf_device_info.argtypes = [vp, i32p, i64p, i64p, i64p]
    lib.jf_load_tensor.argtypes = [vp, C.c_char_p, vp, i64p, i32, i32]
    lib.jf_finalize_weights.argtypes = [vp]
    lib.jf_set_rope_tables.argtypes = [vp, vp, vp, i32]
    lib.jf_kv_info.argtypes = [vp, i32p, i32p, i64p, C.POINTER(vp)]
    lib.jf_forward.argtypes = [vp, i32, vp, i32, vp, i32, vp, i32, i32, i32, vp, vp, vp, vp, fp]
    lib.jf_commit_rows.argtypes = [vp, i32, vp, i32, vp]
    lib.jf_jacobi_run.argtypes = [vp, i32, C.POINTER(PolicyKnobsC), vp, vp, vp, vp, vp, vp, i32, vp, i32, vp, i32,
                                  vp, i32, vp, vp, vp, vp, i32p]
    lib.jf_profile_enable.argtypes = [vp, i32]
    lib.jf_profile_read.argtypes = [vp, i64p, i32, i32p, i32p]
    lib.jf_tp_export.argtypes = [vp, vp]
    lib.jf_tp_import.argtypes = [vp, vp]
    if hasattr(lib, "jf_tp_attach_local"):
        lib.jf_tp_attach_local.argtypes = [vp, C.POINTER(vp)]
    lib.jf_launch_count.argtypes = [vp]
    lib.jf_launch_count.restype = C.c_int64
    for name in EXPORTS:
        if name not in ("jf_last_error", "jf_launch_count") and hasattr(lib, name):
            getattr(lib, name).restype = C.c_int
    _lib = lib
    return lib


_EXC = {JF_ERR_INVALID: ValueError, JF_ERR_CUDA: RuntimeError, JF_ERR_CAPACITY: RuntimeError,
        JF_ERR_UNSUPPORTED: NotImplementedError, JF_ERR_POLICY: RuntimeError}


def check(rc: int, handle=None) -> None:
    """Map a jf_status to the Python exception type the reference raises in the same situation."""
    if rc == JF_OK:
        return
    msg = load().jf_last_error(handle)
    raise _EXC.get(rc, RuntimeError)((msg or b"").decode() or f"libjacobi_b200 error {rc}")
