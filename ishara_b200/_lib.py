"""ctypes binding of libishara_b200.so (the C ABI declared in include/ishara_b200.h).

The library is built in-tree by ``__graft_entry__.build()`` / ``make -C ishara_b200/csrc``. There is no
fallback: if the shared object is missing, importing the compute API raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# ISHARA_B200_LIB selects another build of the SAME library (the profiling build with timeline tracing, `make TRACE=1`)
LIB_PATH = os.environ.get("ISHARA_B200_LIB") or os.path.join(_HERE, "lib", "libishara_b200.so")

OK, ERR_INVALID, ERR_SHAPE, ERR_CUDA, ERR_STATE, ERR_COMM = 0, 1, 2, 3, 4, 5


class IsharaError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"ishara_b200 status {status}: {message}")
        self.status = status
        self.message = message


class Config(C.Structure):
    """ishara_config_t — get_model(...) kwargs + INPUT_SHAPE + len(char_to_num)."""

    _fields_ = [
        ("dim", C.c_int32),
        ("num_conv_squeeze_blocks", C.c_int32),
        ("num_conv_conform_blocks", C.c_int32),
        ("num_conv_per_block", C.c_int32),
        ("kernel_sizes", C.c_int32 * 8),
        ("num_kernel_sizes", C.c_int32),
        ("num_heads", C.c_int32),
        ("expansion_factor", C.c_int32),
        ("transformer_kernel_size", C.c_int32),
        ("frames", C.c_int32),
        ("features", C.c_int32),
        ("num_classes", C.c_int32),
    ]


class AdamW(C.Structure):
    """ishara_adamw_t — BASELINE optimiser (integration.py:675-679,750)."""

    _fields_ = [("lr", C.c_float), ("weight_decay", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float),
                ("eps", C.c_float), ("max_norm", C.c_float)]


class RAdamLookahead(C.Structure):
    """ishara_radam_lookahead_t — the reference's optimiser (c7:68-69), tensorflow_addons defaults."""

    _fields_ = [("lr", C.c_float), ("weight_decay", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float), ("eps", C.c_float),
                ("max_norm", C.c_float), ("sma_threshold", C.c_float), ("sync_period", C.c_int32), ("slow_step_size", C.c_float)]


class GemmArgs(C.Structure):
    """ishara_gemm_args_t"""

    _fields_ = [
        ("a", C.c_void_p),
        ("wt", C.c_void_p),
        ("out0", C.c_void_p),
        ("out1", C.c_void_p),
        ("bias", C.c_void_p),
        ("gate", C.c_void_p),
        ("rowtab", C.c_void_p),
        ("resid", C.c_void_p),
        ("ln0_g", C.c_void_p),
        ("ln0_b", C.c_void_p),
        ("ln0_eps", C.c_float),
        ("ln1_g", C.c_void_p),
        ("ln1_b", C.c_void_p),
        ("ln1_eps", C.c_float),
        ("M", C.c_int32),
        ("N", C.c_int32),
        ("K", C.c_int32),
        ("lda", C.c_int32),
        ("nout", C.c_int32),
        ("rows_per_seq", C.c_int32),
        ("act", C.c_int32),
        ("block_n", C.c_int32),
        ("out_f32", C.c_int32),
        ("row_mode", C.c_int32),
    ]


# name -> (restype, argtypes); must list every symbol include/ishara_b200.h declares
_vp, _i32, _i64, _f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float
SIGNATURES = {
    "ishara_version": (C.c_char_p, []),
    "ishara_last_error": (C.c_char_p, []),
    "ishara_device_count": (_i32, []),
    "ishara_model_create": (_i32, [C.POINTER(Config), _i32, C.POINTER(_vp)]),
    "ishara_model_destroy": (_i32, [_vp]),
    "ishara_model_num_params": (_i32, [_vp]),
    "ishara_model_param_info": (_i32, [_vp, _i32, C.POINTER(C.c_char_p), C.POINTER(_i64), C.POINTER(_i32), C.POINTER(_i64 * 4)]),
    "ishara_model_set_param": (_i32, [_vp, C.c_char_p, _vp, _i64]),
    "ishara_model_get_param": (_i32, [_vp, C.c_char_p, _vp, _i64]),
    "ishara_model_finalize": (_i32, [_vp]),
    "ishara_model_forward": (_i32, [_vp, _vp, _i32, _vp, _vp]),
    "ishara_model_forward_host": (_i32, [_vp, _vp, _i32, _vp]),
    "ishara_model_infer_host": (_i32, [_vp, _vp, _i32, _vp, _i32, _vp, _vp, _vp, _vp]),
    "ishara_ctc_loss": (_i32, [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp]),
    "ishara_greedy_decode": (_i32, [_vp, _i32, _i32, _i32, _i32, _vp, _vp, _vp]),
    "ishara_op_gemm": (_i32, [C.POINTER(GemmArgs), _vp]),
    "ishara_op_dwconv": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _vp]),
    "ishara_op_attention": (_i32, [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _f32, _vp]),
    "ishara_op_relpos_attention": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _f32, _vp]),
    "ishara_op_time_reduce": (_i32, [_vp, _vp, _vp, _f32, _i32, _i32, _i32, _i32, _vp]),
    "ishara_op_upsample_add": (_i32, [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp]),
    "ishara_op_conv2d_subsample": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp]),
    "ishara_op_layernorm": (_i32, [_vp, _vp, _vp, _vp, _f32, _i64, _i32, _vp]),
    "ishara_op_cast_pad": (_i32, [_vp, _vp, _i64, _i32, _i32, _vp]),
    "ishara_ids_to_text": (_i32, [_vp, _vp, _i32, _i32, C.c_char_p, _i32, _vp, _vp]),
    "ishara_device_malloc": (_i32, [_i32, _i64, C.POINTER(_vp)]),
    "ishara_device_free": (_i32, [_i32, _vp]),
    "ishara_host_malloc_pinned": (_i32, [_i64, C.POINTER(_vp)]),
    "ishara_host_free_pinned": (_i32, [_vp]),
    "ishara_memcpy_async": (_i32, [_vp, _vp, _i64, _i32, _vp]),
    "ishara_stream_synchronize": (_i32, [_i32, _vp]),
    "ishara_model_stream": (_vp, [_vp]),
    "ishara_model_set_profile": (_i32, [_vp, _i32]),
    "ishara_model_profile_count": (_i32, [_vp]),
    "ishara_model_profile_entry": (_i32, [_vp, _i32, C.POINTER(C.c_char_p), C.POINTER(C.c_char_p), C.POINTER(_f32),
                                   C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "ishara_launch_count": (C.c_uint64, []),
    "ishara_model_set_debug": (_i32, [_vp, _i32]),
    "ishara_model_debug_fetch": (_i32, [_vp, C.c_char_p, _vp, _i64]),
    "ishara_model_infer_submit": (_i32, [_vp, _vp, _i32, _vp, _i32, _vp, _vp, _vp, _vp]),
    "ishara_model_infer_collect": (_i32, [_vp]),
    "ishara_edit_distances": (_i32, [_vp, _vp, _vp, _vp, _i32, _vp]),
    "ishara_preprocess": (_i32, [_vp, _vp, _i32, _i32, _vp, _vp, _i32, _i32, _vp, _vp]),
    "ishara_model_train_configure": (_i32, [_vp, _f32, C.c_uint64, _i32]),
    "ishara_model_train_forward_backward": (_i32, [_vp, _vp, _vp, _i32, _i32, C.POINTER(_f32), _vp]),
    "ishara_model_train_grad_buffer": (_i32, [_vp, C.POINTER(_vp), C.POINTER(_i64)]),
    "ishara_model_train_apply": (_i32, [_vp, _vp, _f32, _vp]),
    "ishara_model_train_step": (_i32, [_vp, _vp, _vp, _i32, _i32, _vp, C.POINTER(_f32), _vp]),
    "ishara_model_train_step_host": (_i32, [_vp, _vp, _vp, _i32, _i32, _vp, C.POINTER(_f32)]),
    "ishara_model_train_sync": (_i32, [_vp]),
    "ishara_model_train_param_grad": (_i32, [_vp, C.c_char_p, _vp, _i64]),
    "ishara_model_train_fetch": (_i32, [_vp, C.c_char_p, _i32, _vp, _i64]),
    "ishara_model_train_apply_radam": (_i32, [_vp, _vp, _f32, _vp]),
    "ishara_model_train_state_info": (_i32, [_vp, C.POINTER(_i64), C.POINTER(_i64), C.POINTER(_i64), C.POINTER(_i32)]),
    "ishara_model_train_state_get": (_i32, [_vp, _i32, _vp, _i64]),
    "ishara_model_train_state_set": (_i32, [_vp, _i32, _vp, _i64]),
    "ishara_model_train_state_set_counters": (_i32, [_vp, _i64, _i64]),
    "ishara_model_set_mask_mode": (_i32, [_vp, _i32]),
    "ishara_model_forward_masked": (_i32, [_vp, _vp, _vp, _i32, _vp, _vp]),
    "ishara_model_train_loss": (_i32, [_vp, C.POINTER(_f32), _vp]),
    "ishara_model_train_counters": (_i32, [_vp, C.POINTER(_i64), C.POINTER(_i64), C.POINTER(_i64)]),
    "ishara_comm_unique_id": (_i32, [_vp]),
    "ishara_model_comm_init": (_i32, [_vp, _vp, _i32, _i32]),
    "ishara_model_comm_destroy": (_i32, [_vp]),
    "ishara_nccl_version": (_i32, []),
    "ishara_comm_bucket_plan": (_i32, [C.POINTER(_i64), _i32, _i64, _i64, C.POINTER(_i64), C.POINTER(_i64)]),
}

_lib = None


def load() -> C.CDLL:
    """Load the shared library (once) and attach prototypes. Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C ishara_b200/csrc`. ishara_b200 has no CPU or PyTorch fallback."
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status: int) -> None:
    if status != OK:
        msg = load().ishara_last_error()
        raise IsharaError(status, msg.decode() if msg else "")
