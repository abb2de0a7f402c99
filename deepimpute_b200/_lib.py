"""ctypes binding of ``include/deepimpute_b200.h`` (the same stub INTEGRATION.md gives to reference maintainers).

There is no fallback: if the shared library has not been built (``python -c "import __graft_entry__ as g;
g.build()"`` or ``make -C deepimpute_b200/csrc``) loading raises, and every non-zero return code raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libdeepimpute_b200.so")

DI_MATH = {"fp32": 0, "tf32": 1, "tf32x3": 2}
DI_DTYPE = {"float32": 0, "float64": 1}
DI_POLICY = {None: 0, "none": 0, "restore": 1, "max": 2}


class DiConfig(C.Structure):
    _fields_ = [("n_subnets", C.c_int32), ("hidden", C.c_int32), ("sub_outputdim", C.c_int32),
                ("batch_size", C.c_int32), ("learning_rate", C.c_float), ("beta1", C.c_float),
                ("beta2", C.c_float), ("epsilon", C.c_float), ("dropout_rate", C.c_float),
                ("seed", C.c_uint64), ("math_mode", C.c_int32), ("device", C.c_int32)]


_f32p = C.POINTER(C.c_float)
_i32p = C.POINTER(C.c_int32)
_i64p = C.POINTER(C.c_int64)
_H = C.c_void_p

# name -> (restype, argtypes); kept in the order of the header
SIGNATURES = {
    "di_create": (C.c_int, [C.POINTER(_H), C.POINTER(DiConfig), _i32p]),
    "di_destroy": (None, [_H]),
    "di_last_error": (C.c_char_p, [_H]),
    "di_set_subnet_ids": (C.c_int, [_H, _i32p]),
    "di_upload_matrix": (C.c_int, [_H, _f32p, C.c_int64, C.c_int64]),
    "di_upload_counts": (C.c_int, [_H, C.c_void_p, C.c_int32, C.c_int64, C.c_int64]),
    "di_set_partition": (C.c_int, [_H, _i32p, _i64p, _i32p]),
    "di_set_split": (C.c_int, [_H, _i32p, C.c_int64, _i32p, C.c_int64]),
    "di_set_weights": (C.c_int, [_H, C.c_int32, _f32p, _f32p, _f32p, _f32p]),
    "di_get_weights": (C.c_int, [_H, C.c_int32, _f32p, _f32p, _f32p, _f32p]),
    "di_get_adam_state": (C.c_int, [_H, C.c_int32] + [_f32p] * 8 + [_i64p]),
    "di_train_epoch": (C.c_int, [_H, _i32p, C.c_int64, _f32p, _f32p]),
    "di_train_step": (C.c_int, [_H, _i32p, C.c_int32, C.c_int64, _f32p]),
    "di_validation_loss": (C.c_int, [_H, _f32p]),
    "di_predict": (C.c_int, [_H, _i32p, C.c_int64, _f32p]),
    "di_predict_device": (C.c_int, [_H, _i32p, C.c_int64, C.c_void_p, C.c_int64]),
    "di_impute": (C.c_int, [_H, C.c_int32, _i32p, C.c_int64, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p]),
    "di_gene_stats": (C.c_int, [C.c_int32, C.c_void_p, C.c_int32, C.c_int64, C.c_int64, C.POINTER(C.c_double),
                                C.POINTER(C.c_double), _f32p]),
    "di_gene_stats_last_error": (C.c_char_p, []),
    "di_corr_topk": (C.c_int, [C.c_int32, _f32p, C.c_int64, C.c_int64, _i32p, C.c_int64, _i32p, C.c_int32, C.c_int32,
                               C.c_int32, _i32p, _f32p, _f32p]),
    "di_corr_last_error": (C.c_char_p, []),
    "di_device_sync": (C.c_int, [_H]),
    "di_timer_start": (C.c_int, [_H]),
    "di_timer_stop": (C.c_int, [_H, _f32p]),
    "di_launch_count": (C.c_int64, [_H]),
    "di_last_device_ms": (C.c_float, [_H]),
    "di_set_profiling": (C.c_int, [_H, C.c_int32]),
    "di_kernel_ms": (C.c_float, [_H, C.c_char_p]),
    "di_kernel_launches": (C.c_int64, [_H, C.c_char_p]),
    "di_debug_read": (C.c_int, [_H, C.c_char_p, _f32p, C.c_int64, _i64p]),
    "di_describe": (C.c_char_p, [_H]),
    "di_graph_fallbacks": (C.c_int64, [_H]),
    "di_version": (C.c_int, []),
    "di_math_mode_available": (C.c_int, [C.c_int32]),
}

_lib = None


def load():
    """dlopen the C-ABI library and attach signatures.  Raises if it is missing -- never falls back."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "deepimpute_b200: CUDA library not built ({} missing). Build it with "
                "`python -c 'import __graft_entry__ as g; g.build()'` -- there is no CPU fallback.".format(LIB_PATH))
        lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def f32(a):
    return a.ctypes.data_as(_f32p)


def i32(a):
    return a.ctypes.data_as(_i32p)


def i64(a):
    return a.ctypes.data_as(_i64p)
