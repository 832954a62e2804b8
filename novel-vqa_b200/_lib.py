"""ctypes binding of libnvqa.so -- the same declarations a LuaJIT ``ffi.cdef(include/nvqa.h)`` makes
(INTEGRATION.md).  No torch import here: the library owns device memory and streams."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libnvqa.so")

c_i32p = C.POINTER(C.c_int32)
c_f32p = C.POINTER(C.c_float)
c_i64p = C.POINTER(C.c_int64)


class nvqa_config(C.Structure):
    _fields_ = [(n, C.c_int32) for n in
                ("arch", "V", "E", "H", "L", "I", "C", "O", "T", "B", "precision", "img_norm", "device")] + \
               [("dropout", C.c_float)]


PREC_FP32_SIMT, PREC_BF16X3, PREC_BF16, PREC_BF16X2 = 0, 1, 2, 3
BLOCK_ENCODER, BLOCK_EMBEDDING, BLOCK_MULTIMODAL = 0, 1, 2
MODE_EVAL, MODE_TRAIN = 0, 1
FUSION_AXB, FUSION_ASKIPB = 0, 1
PHASE_HEAD, PHASE_LSTM, PHASE_EMBED, PHASE_ALL = 0, 1, 2, 3

# name -> (restype, argtypes); must list every symbol include/nvqa.h declares (tests/test_abi.py checks)
SIGNATURES = {
    "nvqa_last_error": (C.c_char_p, []),
    "nvqa_version": (C.c_int, []),
    "nvqa_device_count": (C.c_int, []),
    "nvqa_model_create": (C.c_int, [C.POINTER(nvqa_config), C.POINTER(C.c_void_p)]),
    "nvqa_model_destroy": (C.c_int, [C.c_void_p]),
    "nvqa_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "nvqa_sync": (C.c_int, [C.c_void_p]),
    "nvqa_param_count": (C.c_int, [C.c_void_p, C.c_int, c_i64p]),
    "nvqa_params_set": (C.c_int, [C.c_void_p, C.c_int, c_f32p]),
    "nvqa_params_get": (C.c_int, [C.c_void_p, C.c_int, c_f32p]),
    "nvqa_grads_get": (C.c_int, [C.c_void_p, C.c_int, c_f32p]),
    "nvqa_rms_get": (C.c_int, [C.c_void_p, C.c_int, c_f32p]),
    "nvqa_rms_set": (C.c_int, [C.c_void_p, C.c_int, c_f32p]),
    "nvqa_device_views": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), c_i64p]),
    "nvqa_right_align": (C.c_int, [c_i32p, c_i32p, C.c_int32, C.c_int32, c_i32p]),
    "nvqa_pack_batch": (C.c_int, [c_i32p, c_i32p, C.c_int32, C.c_int32, c_i32p, c_i32p, c_i32p, c_i32p, c_i32p, c_i32p]),
    "nvqa_mc_select": (C.c_int, [c_f32p, c_i32p, C.c_int32, C.c_int32, C.c_int32, c_i32p]),
    "nvqa_set_batch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32]),
    "nvqa_set_batch_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32]),
    "nvqa_set_steps": (C.c_int, [C.c_void_p, C.c_int32]),
    "nvqa_set_masks": (C.c_int, [C.c_void_p] + [C.c_void_p] * 5),
    "nvqa_forward": (C.c_int, [C.c_void_p, C.c_int, C.c_uint64]),
    "nvqa_loss": (C.c_int, [C.c_void_p, c_f32p]),
    "nvqa_backward": (C.c_int, [C.c_void_p, C.c_int]),
    "nvqa_rmsprop_step": (C.c_int, [C.c_void_p] + [C.c_float] * 6),
    "nvqa_adam_step": (C.c_int, [C.c_void_p] + [C.c_float] * 7),
    "nvqa_logprobs_get": (C.c_int, [C.c_void_p, C.c_int32, c_f32p]),
    "nvqa_set_lookup_grad_literal": (C.c_int, [C.c_void_p, C.c_int32]),
    "nvqa_set_stale_h0_literal": (C.c_int, [C.c_void_p, C.c_int32]),
    "nvqa_set_variant": (C.c_int, [C.c_void_p, C.c_int32, C.c_float, C.c_int32]),
    "nvqa_scores_get": (C.c_int, [C.c_void_p, c_f32p]),
    "nvqa_argmax_get": (C.c_int, [C.c_void_p, c_i32p]),
    "nvqa_state_get": (C.c_int, [C.c_void_p, c_f32p]),
    "nvqa_train_step_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32,
                                       C.c_float, C.c_uint64, c_f32p]),
    "nvqa_train_step": (C.c_int, [C.c_void_p, C.c_float, C.c_uint64]),
    "nvqa_eval_step_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    "nvqa_lstm_cell_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    "nvqa_axb_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    "nvqa_grads_zero": (C.c_int, [C.c_void_p, C.c_int]),
    "nvqa_embedding_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    "nvqa_embedding_backward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32]),
    "nvqa_lstm_cell_backward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p,
                                          C.c_void_p]),
    "nvqa_axb_backward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32,
                                    C.c_void_p, C.c_void_p]),
    "nvqa_multimodal_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32,
                                          C.c_void_p]),
    "nvqa_multimodal_backward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                           C.c_int32, C.c_void_p, C.c_void_p]),
    "nvqa_rmsprop_vector": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64] + [C.c_float] * 6),
    "nvqa_cross_entropy": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, c_f32p, C.c_void_p]),
    "nvqa_dp_blob_size": (C.c_int, []),
    "nvqa_dp_export": (C.c_int, [C.c_void_p, C.c_void_p]),
    "nvqa_dp_connect": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]),
    "nvqa_dp_disconnect": (C.c_int, [C.c_void_p]),
    "nvqa_dp_rmsprop_step": (C.c_int, [C.c_void_p] + [C.c_float] * 5),
    "nvqa_dp_train_step": (C.c_int, [C.c_void_p, C.c_float, C.c_uint64] + [C.c_float] * 4),
    "nvqa_dp_status": (C.c_int, [C.c_void_p, C.POINTER(C.c_int32)]),
    "nvqa_dp_layout": (C.c_int, [C.c_void_p, C.POINTER(C.c_int32)]),
    "nvqa_host_alloc": (C.c_int, [C.POINTER(C.c_void_p), C.c_int64]),
    "nvqa_host_free": (C.c_int, [C.c_void_p]),
    "nvqa_device_alloc": (C.c_int, [C.POINTER(C.c_void_p), C.c_int64]),
    "nvqa_device_free": (C.c_int, [C.c_void_p]),
    "nvqa_memcpy_h2d": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]),
    "nvqa_memcpy_d2h": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]),
    "nvqa_profile": (C.c_int, [C.c_void_p, C.c_int]),
    "nvqa_profile_report": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int32]),
    "nvqa_launch_count": (C.c_int64, []),
    "nvqa_gemm_test": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                 C.c_void_p, C.c_void_p]),
    "nvqa_gemm_test_ex": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_int32, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
}

_lib = None


class NvqaError(RuntimeError):
    pass


def load():
    """dlopen libnvqa.so; fails loudly when the CUDA extension has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NvqaError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                            "(there is no CPU fallback)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def check(rc):
    if rc != 0:
        raise NvqaError(load().nvqa_last_error().decode())
