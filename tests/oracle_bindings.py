"""ctypes bindings to the CHECKERS: oracle/liboracle.so (our CPU restatement) and, when it was
built, oracle/_ref/libat_ref.so (the reference's own objects).  Test infrastructure only --
the product package never imports this module."""
import ctypes as C
import os
import re
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")

N_BITS, N, L, NL, M, P = 10, 1024, 46, 93, 3, 3
HALF_W = HALF_H = 50
CELLS = 101 * 101
RATE_HZ, SPEED, PX_PER_M, HEIGHT = 50000.0, 343.0, 24.0, 1.2

u8p = np.ctypeslib.ndpointer(np.uint8, flags="C")
i16p = np.ctypeslib.ndpointer(np.int16, flags="C")
i32p = np.ctypeslib.ndpointer(np.int32, flags="C")
i64p = np.ctypeslib.ndpointer(np.int64, flags="C")
f32p = np.ctypeslib.ndpointer(np.float32, flags="C")


class AtoConfig(C.Structure):
    _fields_ = [("n_mics", C.c_int32), ("n_bits", C.c_int32), ("max_shift", C.c_int32),
                ("window", C.c_void_p), ("window_bits", C.c_int32),
                ("lut", C.c_void_p), ("n_cells", C.c_int32)]


def window_tables():
    """Parse the generated (committed) window header -> {1024: int16[1024], 4096: int16[4096]}."""
    src = open(os.path.join(ROOT, "audio_triangulation_b200", "csrc", "at_window_tables.h")).read()
    out = {}
    for n in (1024, 4096):
        body = src.split(f"AT_WINDOW_{n}[{n}] = {{")[1].split("};")[0]
        out[n] = np.array([int(v) for v in re.findall(r"-?\d+", body)], dtype=np.int16)
        assert out[n].size == n
    return out


def _build():
    subprocess.run(["make", "-s", "-C", ORACLE_DIR], check=True, capture_output=True)


def load_oracle():
    path = os.path.join(ORACLE_DIR, "liboracle.so")
    if not os.path.exists(path):
        _build()
    lib = C.CDLL(path)
    lib.ato_dc_remove.restype = C.c_int64
    lib.ato_dc_remove.argtypes = [i16p, C.c_int, C.c_int, i16p]
    lib.ato_shift8.argtypes = [i16p, C.c_int]
    lib.ato_window.argtypes = [i16p, C.c_int, i16p, C.c_int]
    lib.ato_xcorr.argtypes = [i16p, i16p, C.c_int, C.c_int, i64p, i32p]
    lib.ato_gauss.argtypes = [i64p, C.c_int, C.c_int]
    lib.ato_average.argtypes = [i64p, i32p, C.POINTER(C.c_uint64), i64p, C.c_int, C.c_uint64]
    lib.ato_capture.restype = C.c_long
    lib.ato_capture.argtypes = [u8p, C.c_size_t, C.c_int, C.c_int, i16p, i32p]
    lib.ato_mics_triangle.argtypes = [C.c_float, C.c_float, C.c_float, C.c_int, C.c_int, f32p]
    lib.ato_lut_build.argtypes = [f32p, C.c_int, C.c_int, C.c_float, C.c_float, C.c_int, C.c_int,
                                  C.c_float, C.c_float, u8p]
    lib.ato_lut_build_points.argtypes = [f32p, C.c_int, C.c_int, C.c_float, C.c_float, f32p, C.c_int, u8p]
    lib.ato_heatmap.argtypes = [i64p, u8p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.ato_synth_frames.argtypes = [C.c_uint64, C.c_uint32, C.c_size_t, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    lib.ato_synth_plain.argtypes = [C.c_int]
    lib.ato_localize.argtypes = [C.POINTER(AtoConfig), u8p, C.c_void_p, C.c_size_t,
                                 C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    return lib


def load_ref(fast=False):
    """The reference's own objects, or None when oracle/_ref was not built (no /root/reference)."""
    name = "libat_ref_fast.so" if fast else "libat_ref.so"
    path = os.path.join(ORACLE_DIR, "_ref", name)
    if not os.path.exists(path):
        if os.path.isdir("/root/reference/src"):
            _build()
        if not os.path.exists(path):
            return None
    lib = C.CDLL(path)
    lib.ref_set_time.argtypes = [C.c_uint64]
    lib.ref_layout.argtypes = [i64p]
    lib.ref_mics.argtypes = [f32p]
    lib.ref_frame_stages.argtypes = [u8p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.ref_localize_frames.argtypes = [u8p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_int, C.c_uint64]
    lib.ref_capture.restype = C.c_long
    lib.ref_capture.argtypes = [u8p, C.c_size_t, C.c_void_p]
    # the reference's own functions, callable directly on numpy views of its structs
    vp = C.c_void_p
    for name, res, args in (("rolling_buffer_init", None, [vp]), ("rolling_buffer_push", None, [vp, C.c_int16]),
                            ("rolling_buffer_write_out", None, [vp, vp]),
                            ("rolling_buffer_get_incoming_power", C.c_int64, [vp]),
                            ("rolling_buffer_get_outgoing_power", C.c_int64, [vp]),
                            ("buffer_normalize_range", None, [vp]), ("buffer_window", None, [vp]),
                            ("correlations_init", None, [vp, vp, vp]), ("correlations_average", None, [vp, vp]),
                            ("microphones_init", None, [])):
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    return lib


# numpy views of the reference structs (LP64; sizes asserted against ref_layout in tests)
CORR_DT = np.dtype([("correlations", np.int64, (NL,)), ("best_shift", np.int32), ("_pad", np.int32),
                    ("last_update", np.uint64)])
BUFFER_DT = np.dtype([("buffer", np.int16, (N,)), ("power", np.int64)])
RING_DT = np.dtype([("head", np.int32), ("_pad0", np.int32), ("incoming_power", np.int64),
                    ("incoming_total", np.int64), ("outgoing_power", np.int64),
                    ("outgoing_total", np.int64), ("is_full", np.uint8), ("_pad1", np.uint8),
                    ("buffer", np.int16, (N,)), ("_pad2", np.uint8, (6,))])


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Oracle:
    """Reference-shaped (3 mics, 1024, +-46) convenience wrapper over liboracle.so."""

    def __init__(self, n_mics=3, n_bits=10, max_shift=46, window=None, lut=None, n_cells=CELLS):
        self.lib = load_oracle()
        tabs = window_tables()
        self.window = np.ascontiguousarray(window if window is not None else
                                           tabs[4096 if n_bits > 10 else 1024])
        self.n_mics, self.n_bits, self.L = n_mics, n_bits, max_shift
        self.n_pairs = n_mics * (n_mics - 1) // 2
        self.lut = lut
        self.n_cells = n_cells
        if lut is None and n_mics == 3:
            self.lut = self.reference_lut()

    def mics(self):
        xy = np.zeros(6, np.float32)
        self.lib.ato_mics_triangle(0.132, 0.15, 0.20, 1, 0, xy)
        return xy.reshape(3, 2)

    def reference_lut(self):
        idx = np.zeros((3, 101, 101), np.uint8)
        self.lib.ato_lut_build(self.mics().reshape(-1), 3, self.L, RATE_HZ, SPEED, HALF_W, HALF_H,
                               PX_PER_M, HEIGHT, idx.reshape(-1))
        return idx.reshape(3, -1)

    def synth(self, n_frames, seed=0xA7D10, flags=0, first_frame=0, nthreads=0):
        """Frames of the bench workload generated on the host without the product library (oracle/synth_host.cpp)."""
        adc = np.empty((n_frames, 3, N), np.uint8)
        heads = np.zeros(n_frames, np.int32); cell = np.zeros(n_frames, np.int32)
        self.lib.ato_synth_frames(seed, flags, first_frame, n_frames, adc.ctypes.data, heads.ctypes.data, cell.ctypes.data,
                                  nthreads or (os.cpu_count() or 1))
        return adc, heads, cell

    def localize(self, adc, heads=None, want_corr=True, want_raw=False, want_cell=True, nthreads=1):
        adc = np.ascontiguousarray(adc, np.uint8)
        F = adc.shape[0]
        nl = 2 * self.L + 1
        cfg = AtoConfig(self.n_mics, self.n_bits, self.L, self.window.ctypes.data,
                        int(np.log2(self.window.size)),
                        self.lut.ctypes.data if self.lut is not None else None, self.n_cells)
        lags = np.zeros((F, self.n_pairs), np.int32)
        corr = np.zeros((F, self.n_pairs, nl), np.int64) if want_corr else None
        raw = np.zeros((F, self.n_pairs, nl), np.int64) if want_raw else None
        cell = np.zeros(F, np.int32) if (want_cell and self.lut is not None) else None
        high = np.zeros(F, np.int64) if (want_cell and self.lut is not None) else None
        hp = None if heads is None else np.ascontiguousarray(heads, np.int32)
        self.lib.ato_localize(C.byref(cfg), adc.reshape(-1), _ptr(hp), F, _ptr(lags), _ptr(corr),
                              _ptr(raw), _ptr(cell), _ptr(high), nthreads)
        return dict(lags=lags, corr=corr, raw=raw, cell=cell, highest=high)
