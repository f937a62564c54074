"""ctypes loader for libat_b200.so (the C ABI declared in include/at_b200.h).

The library is the product: if it is missing this module raises -- there is no Python or CPU
implementation of the localization path to fall back to.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# AT_LIB_VARIANT=checked | prof selects a debug build of the same library (csrc/Makefile); default: the product
_VARIANT = os.environ.get("AT_LIB_VARIANT", "")
LIB_PATH = os.path.join(_HERE, "libat_b200%s.so" % ("_" + _VARIANT if _VARIANT else ""))

AT_OK, AT_EINVAL, AT_ECUDA, AT_ENOGPU, AT_ENOMEM = 0, -1, -2, -3, -4
AT_MAX_MICS = 8
KERNELS = {"auto": 0, "imad": 1, "imma": 2, "umma": 4}
AT_CORR_PACKED, AT_CORR_STRUCT = 0, 1
AT_LUT_PLANE, AT_LUT_POINTS = 0, 1
SYNTH_INTEGER_DELAYS, SYNTH_RANDOM_HEADS, SYNTH_KATS, SYNTH_MAX_NOISE, SYNTH_WHITE = 1, 2, 4, 8, 16
UBENCH = {"imad_wide": 0, "imad": 1, "dp2a": 2, "dp4a": 3, "imma_s8": 4, "lds": 5, "dfma": 6, "umma_i8": 7, "umma_frame": 8}


class AtConfig(C.Structure):
    _fields_ = [("device", C.c_int32), ("n_mics", C.c_int32), ("n_bits", C.c_int32),
                ("max_shift", C.c_int32), ("kernel", C.c_int32),
                ("sample_rate_hz", C.c_float), ("speed_of_sound", C.c_float),
                ("half_w", C.c_int32), ("half_h", C.c_int32),
                ("px_per_m", C.c_float), ("height_m", C.c_float),
                ("use_reference_triangle", C.c_int32),
                ("mic_xy", (C.c_float * 2) * AT_MAX_MICS),
                ("lut_mode", C.c_int32), ("n_points", C.c_int32), ("points_xyz", C.c_void_p)]


class AtOutputs(C.Structure):
    _fields_ = [("lags", C.c_void_p), ("corr", C.c_void_p), ("corr_layout", C.c_int32),
                ("raw", C.c_void_p), ("cell", C.c_void_p), ("highest", C.c_void_p),
                ("xy", C.c_void_p), ("gate", C.c_void_p), ("classes", C.c_void_p),
                ("windowed", C.c_void_p), ("power", C.c_void_p), ("stats", C.c_void_p)]


class AtError(RuntimeError):
    def __init__(self, code, text):
        super().__init__(f"libat_b200 error {code}: {text}")
        self.code = code


_lib = None


def load():
    """Load the shared library once and declare every prototype of include/at_b200.h."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `make -C audio_triangulation_b200/csrc` "
            "(or __graft_entry__.build()). There is no CPU fallback for the localization path.")
    lib = C.CDLL(LIB_PATH)
    vp, sz, i32, u64 = C.c_void_p, C.c_size_t, C.c_int32, C.c_uint64
    ctx = C.c_void_p
    protos = {
        "at_config_reference": (None, [C.POINTER(AtConfig)]),
        "at_hemisphere_points": (None, [C.c_int, C.c_int, C.c_float, vp]),
        "at_create": (C.c_int, [C.POINTER(AtConfig), C.POINTER(ctx)]),
        "at_destroy": (None, [ctx]),
        "at_last_error": (C.c_char_p, []),
        "at_kernel_launches": (u64, []),
        "at_get_mics": (C.c_int, [ctx, vp]),
        "at_get_lut": (C.c_int, [ctx, vp]),
        "at_shape": (C.c_int, [ctx] + [C.POINTER(i32)] * 5),
        "at_localize_device": (C.c_int, [ctx, vp, vp, sz, C.POINTER(AtOutputs), vp]),
        "at_localize_host": (C.c_int, [ctx, vp, vp, sz, C.POINTER(AtOutputs)]),
        "at_peer_enable": (C.c_int, [ctx, i32]),
        "at_copy_async": (C.c_int, [ctx, vp, vp, sz, vp]),
        "at_host_alloc": (C.c_int, [ctx, sz, C.POINTER(vp)]),
        "at_host_free": (C.c_int, [ctx, vp]),
        "at_shared_alloc": (C.c_int, [ctx, sz, C.POINTER(vp), C.c_char_p]),
        "at_shared_open": (C.c_int, [ctx, C.c_char_p, C.POINTER(vp)]),
        "at_shared_close": (C.c_int, [ctx, vp, i32]),
        "at_localize_host_sharded": (C.c_int, [C.POINTER(ctx), C.c_int, vp, vp, sz, C.POINTER(AtOutputs)]),
        "at_synchronize": (C.c_int, [ctx]),
        "at_average_device": (C.c_int, [ctx, vp, vp, vp, vp, vp, sz, u64, vp]),
        "at_heatmap_device": (C.c_int, [ctx, vp, sz, vp, vp, vp, vp, vp]),
        "at_pair_max_shift": (C.c_int, [ctx, vp]),
        "at_admissible_lags_device": (C.c_int, [ctx, vp, sz, vp, vp]),
        "at_synth_host": (C.c_int, [ctx, u64, C.c_uint32, sz, sz, vp, vp, vp]),
        "at_synth_device": (C.c_int, [ctx, u64, C.c_uint32, sz, sz, vp, vp, vp, vp]),
        "at_microbench": (C.c_int, [ctx, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
        "at_gccphat_device": (C.c_int, [ctx, vp, vp, sz, vp, vp, vp]),
        "at_stream_create": (C.c_int, [ctx, sz, C.POINTER(ctx)]),
        "at_stream_destroy": (None, [ctx]),
        "at_stream_reset": (C.c_int, [ctx, vp]),
        "at_stream_push": (C.c_int, [ctx, vp, sz, vp, vp, vp, vp]),
        "at_set_time_us": (None, [u64]),
        "at_get_time_us": (u64, []),
        # drop-in symbols (reference names)
        "rolling_buffer_init": (None, [vp]),
        "rolling_buffer_push": (None, [vp, C.c_int16]),
        "rolling_buffer_write_out": (None, [vp, vp]),
        "rolling_buffer_get_incoming_power": (C.c_int64, [vp]),
        "rolling_buffer_get_outgoing_power": (C.c_int64, [vp]),
        "buffer_normalize_range": (None, [vp]),
        "buffer_window": (None, [vp]),
        "correlations_init": (None, [vp, vp, vp]),
        "correlations_average": (None, [vp, vp]),
        "microphones_init": (None, []),
    }
    for name, (res, args) in protos.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


def check(rc):
    if rc != AT_OK:
        raise AtError(rc, load().at_last_error().decode(errors="replace"))
