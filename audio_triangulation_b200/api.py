"""Host-side mirror of the C ABI: `Localizer` wraps an at_context; `dropin` exposes the
reference-named per-frame functions (rolling_buffer_write_out, buffer_window, correlations_init,
...) on numpy views of the reference structs.  No compute happens in Python."""
import ctypes as C

import numpy as np

from . import _lib as L

NL_REF = 93
CORR_DT = np.dtype([("correlations", np.int64, (NL_REF,)), ("best_shift", np.int32), ("_pad", np.int32),
                    ("last_update", np.uint64)])
BUFFER_DT = np.dtype([("buffer", np.int16, (1024,)), ("power", np.int64)])
RING_DT = np.dtype([("head", np.int32), ("_pad0", np.int32), ("incoming_power", np.int64),
                    ("incoming_total", np.int64), ("outgoing_power", np.int64),
                    ("outgoing_total", np.int64), ("is_full", np.uint8), ("_pad1", np.uint8),
                    ("buffer", np.int16, (1024,)), ("_pad2", np.uint8, (6,))])

_ALL = ("lags", "corr", "raw", "cell", "highest", "xy", "gate", "classes", "windowed", "power")


def _np_ptr(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


def hemisphere_points(n_az, n_el, radius_m):
    """n_az x n_el directions on the upper hemisphere (at_hemisphere_points): float32 [n_el * n_az, 3]."""
    xyz = np.zeros((n_el * n_az, 3), np.float32)
    L.load().at_hemisphere_points(n_az, n_el, C.c_float(radius_m), C.c_void_p(xyz.ctypes.data))
    return xyz


class DevPtr:
    """A raw device pointer where the API takes a tensor (only .data_ptr() is used)."""

    def __init__(self, ptr):
        self.ptr = int(ptr)

    def data_ptr(self):
        return self.ptr


class SharedArray:
    """Result memory that lives on one GPU and is mapped by the processes driving the others (Localizer.shared_alloc /
    shared_open).  .view(offset, nbytes) gives a DevPtr to pass as an output of localize_device; on the owning process
    .tensor() exposes the bytes to torch (zero copy, __cuda_array_interface__)."""

    def __init__(self, loc, ptr, nbytes, handle, opened):
        self.loc, self.ptr, self.nbytes, self.handle, self.opened = loc, ptr, nbytes, handle, opened
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}

    def view(self, offset, nbytes):
        assert 0 <= offset and offset + nbytes <= self.nbytes
        return DevPtr(self.ptr + offset)

    def tensor(self):
        import torch
        return torch.as_tensor(self, device="cuda:%d" % self.loc.device)

    def close(self):
        if self.ptr:
            L.check(self.loc.lib.at_shared_close(self.loc.ctx, C.c_void_p(self.ptr), 1 if self.opened else 0))
            self.ptr = 0


class Localizer:
    """One at_context: a shape (mics, frame length, lag range), its tables and streams on one GPU."""

    def __init__(self, device=0, n_mics=3, n_bits=10, max_shift=46, kernel="auto", mic_xy=None,
                 sample_rate_hz=50000.0, speed_of_sound=343.0, half_w=50, half_h=50, px_per_m=24.0,
                 height_m=1.2, points=None):
        """points: optional float32 [n, 3] candidate source positions (AT_LUT_POINTS, the 3-D form of the lag look-up
        table, e.g. hemisphere_points()); default: the reference's plane-on-sphere grid."""
        self.lib = L.load()
        cfg = L.AtConfig()
        self.lib.at_config_reference(C.byref(cfg))
        cfg.device, cfg.n_mics, cfg.n_bits, cfg.max_shift = device, n_mics, n_bits, max_shift
        cfg.kernel = L.KERNELS[kernel]
        cfg.sample_rate_hz, cfg.speed_of_sound = sample_rate_hz, speed_of_sound
        cfg.half_w, cfg.half_h, cfg.px_per_m, cfg.height_m = half_w, half_h, px_per_m, height_m
        if mic_xy is not None:
            xy = np.asarray(mic_xy, np.float32).reshape(n_mics, 2)
            cfg.use_reference_triangle = 0
            for m in range(n_mics):
                cfg.mic_xy[m][0], cfg.mic_xy[m][1] = float(xy[m, 0]), float(xy[m, 1])
        elif n_mics != 3:   # our choice (the reference has no other geometry): circle of radius 0.1 m
            cfg.use_reference_triangle = 0
            for m in range(min(n_mics, L.AT_MAX_MICS)):
                ang = 2.0 * np.pi * m / n_mics
                cfg.mic_xy[m][0], cfg.mic_xy[m][1] = float(np.float32(0.1 * np.cos(ang))), float(np.float32(0.1 * np.sin(ang)))
        if points is not None:
            self._points = np.ascontiguousarray(points, np.float32).reshape(-1, 3)
            cfg.lut_mode, cfg.n_points, cfg.points_xyz = L.AT_LUT_POINTS, self._points.shape[0], self._points.ctypes.data
        self.cfg = cfg
        self.ctx = C.c_void_p()
        L.check(self.lib.at_create(C.byref(cfg), C.byref(self.ctx)))
        v = [C.c_int32() for _ in range(5)]
        L.check(self.lib.at_shape(self.ctx, *[C.byref(x) for x in v]))
        self.n_mics, self.n_samples, self.n_pairs, self.n_lags, self.n_cells = [x.value for x in v]
        self.device = device

    def close(self):
        if getattr(self, "ctx", None) and self.ctx.value:
            self.lib.at_destroy(self.ctx)
            self.ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- geometry products
    def mics(self):
        xy = np.zeros((self.n_mics, 2), np.float32)
        L.check(self.lib.at_get_mics(self.ctx, _np_ptr(xy)))
        return xy

    def lut(self):
        lut = np.zeros((self.n_pairs, self.n_cells), np.uint8)
        L.check(self.lib.at_get_lut(self.ctx, _np_ptr(lut)))
        return lut

    # ---- output allocation
    def _shapes(self, F, struct_corr):
        P, NL, M, N = self.n_pairs, self.n_lags, self.n_mics, self.n_samples
        return {"lags": ((F, P), "int32"), "corr": ((F, P, NL + 2 if struct_corr else NL), "int64"),
                "raw": ((F, P, NL), "int64"), "cell": ((F,), "int32"), "highest": ((F,), "int64"),
                "xy": ((F, 2), "float32"), "gate": ((F,), "uint8"), "classes": ((F, self.n_cells), "uint8"),
                "windowed": ((F, M, N), "int16"), "power": ((F, M), "int64")}

    # ---- batched path, device-resident (torch tensors)
    def localize_device(self, adc, heads=None, want=("lags",), out=None, struct_corr=False, stream=None):
        """adc: torch.uint8 CUDA tensor [F, mics, N] (ring order).  Returns dict of CUDA tensors.
        Asynchronous on `stream` (a torch.cuda.Stream; default: the current torch stream)."""
        import torch
        assert adc.is_cuda and adc.dtype == torch.uint8 and adc.is_contiguous()
        F = adc.shape[0]
        assert adc.numel() == F * self.n_mics * self.n_samples
        shapes = self._shapes(F, struct_corr)
        res = {} if out is None else out
        o = L.AtOutputs()
        o.corr_layout = L.AT_CORR_STRUCT if struct_corr else L.AT_CORR_PACKED
        for k in want:
            if k == "stats":   # int64[5] counters the kernel adds to (zeroed here when first allocated)
                if k not in res:
                    res[k] = torch.zeros(5, dtype=torch.int64, device=adc.device)
            elif k not in res:
                res[k] = torch.empty(shapes[k][0], dtype=getattr(torch, shapes[k][1]), device=adc.device)
            setattr(o, k, res[k].data_ptr())
        st = stream if stream is not None else torch.cuda.current_stream(adc.device)
        hp = None if heads is None else C.c_void_p(heads.data_ptr())
        L.check(self.lib.at_localize_device(self.ctx, C.c_void_p(adc.data_ptr()), hp, F, C.byref(o),
                                            C.c_void_p(st.cuda_stream)))
        return res

    # ---- batched path from host memory (numpy arrays or pinned torch CPU tensors)
    def localize_host(self, adc, heads=None, want=("lags",), out=None, struct_corr=False):
        F = adc.shape[0]
        shapes = self._shapes(F, struct_corr)
        res = {} if out is None else out
        o = L.AtOutputs()
        o.corr_layout = L.AT_CORR_STRUCT if struct_corr else L.AT_CORR_PACKED
        for k in want:
            if k not in res:
                res[k] = np.empty(shapes[k][0], dtype=shapes[k][1])
            setattr(o, k, _host_ptr(res[k]))
        L.check(self.lib.at_localize_host(self.ctx, C.c_void_p(_host_ptr(adc)),
                                          None if heads is None else C.c_void_p(_host_ptr(heads)), F, C.byref(o)))
        return res

    def shared_alloc(self, nbytes):
        """Device memory of this context's GPU that other processes can map (at_shared_alloc).  Returns a SharedArray."""
        ptr, h = C.c_void_p(), C.create_string_buffer(64)
        L.check(self.lib.at_shared_alloc(self.ctx, nbytes, C.byref(ptr), h))
        return SharedArray(self, ptr.value, nbytes, h.raw, opened=False)

    def shared_open(self, handle, nbytes):
        """Map another process's SharedArray (its 64-byte .handle) for this context's GPU (at_shared_open)."""
        ptr = C.c_void_p()
        L.check(self.lib.at_shared_open(self.ctx, bytes(handle), C.byref(ptr)))
        return SharedArray(self, ptr.value, nbytes, bytes(handle), opened=True)

    def copy_async(self, dst, src, nbytes, stream):
        """Device-to-device copy on a torch stream; dst / src: tensors or DevPtr (e.g. a SharedArray view)."""
        L.check(self.lib.at_copy_async(self.ctx, C.c_void_p(dst.data_ptr()), C.c_void_p(src.data_ptr()), nbytes,
                                       C.c_void_p(stream.cuda_stream)))

    def peer_enable(self, peer_device):
        """Allow this context's kernels to store into memory of `peer_device` (frame-sharded runs write their results
        straight into rank 0's arrays)."""
        L.check(self.lib.at_peer_enable(self.ctx, int(peer_device)))

    def synchronize(self):
        L.check(self.lib.at_synchronize(self.ctx))

    # ---- temporal stage / likelihood map on device arrays
    def average_device(self, est, est_best, est_time, fresh, gate, now_us, stream=None):
        import torch
        st = stream if stream is not None else torch.cuda.current_stream(est.device)
        A = est.shape[0]
        L.check(self.lib.at_average_device(self.ctx, est.data_ptr(), est_best.data_ptr(), est_time.data_ptr(),
                                           fresh.data_ptr(), None if gate is None else gate.data_ptr(), A,
                                           int(now_us), C.c_void_p(st.cuda_stream)))

    def heatmap_device(self, corr, want=("cell", "highest"), stream=None):
        import torch
        st = stream if stream is not None else torch.cuda.current_stream(corr.device)
        A = corr.shape[0]
        shapes = self._shapes(A, False)
        res = {k: torch.empty(shapes[k][0], dtype=getattr(torch, shapes[k][1]), device=corr.device) for k in want}
        g = lambda k: res[k].data_ptr() if k in res else None  # noqa: E731
        L.check(self.lib.at_heatmap_device(self.ctx, corr.data_ptr(), A, g("cell"), g("highest"), g("xy"),
                                           g("classes"), C.c_void_p(st.cuda_stream)))
        return res

    def pair_max_shift(self):
        """Physically admissible |lag| per pair: ceil(distance * sample rate / speed of sound), clipped to max_shift."""
        out = np.zeros(self.n_pairs, np.int32)
        L.check(self.lib.at_pair_max_shift(self.ctx, out.ctypes.data))
        return out

    def admissible_lags_device(self, curves, stream=None):
        """First-max arg-max of int64 curves [F][pairs][2L+1] inside each pair's admissible window -> int32 [F][pairs]."""
        import torch
        assert curves.is_cuda and curves.dtype == torch.int64 and curves.is_contiguous()
        F = curves.shape[0]
        lags = torch.empty((F, self.n_pairs), dtype=torch.int32, device=curves.device)
        st = stream if stream is not None else torch.cuda.current_stream(curves.device)
        L.check(self.lib.at_admissible_lags_device(self.ctx, curves.data_ptr(), F, lags.data_ptr(), C.c_void_p(st.cuda_stream)))
        return lags

    def gccphat_device(self, adc, heads=None, want_peak=False, stream=None):
        """FFT / GCC-PHAT variant of the TDOA stage (crossover study; not a reference algorithm)."""
        import torch
        F = adc.shape[0]
        lags = torch.empty((F, self.n_pairs), dtype=torch.int32, device=adc.device)
        peak = torch.empty((F, self.n_pairs), dtype=torch.float32, device=adc.device) if want_peak else None
        st = stream if stream is not None else torch.cuda.current_stream(adc.device)
        L.check(self.lib.at_gccphat_device(self.ctx, adc.data_ptr(), None if heads is None else heads.data_ptr(), F,
                                           lags.data_ptr(), None if peak is None else peak.data_ptr(),
                                           C.c_void_p(st.cuda_stream)))
        return (lags, peak) if want_peak else lags

    # ---- synthetic frames
    def synth_host(self, n_frames, seed=0xA7D10, flags=0, first_frame=0):
        adc = np.empty((n_frames, self.n_mics, self.n_samples), np.uint8)
        heads = np.empty(n_frames, np.int32)
        cells = np.empty(n_frames, np.int32)
        L.check(self.lib.at_synth_host(self.ctx, seed, flags, first_frame, n_frames, _np_ptr(adc), _np_ptr(heads),
                                       _np_ptr(cells)))
        return adc, heads, cells

    def synth_device(self, n_frames, seed=0xA7D10, flags=0, first_frame=0, out=None, stream=None):
        import torch
        dev = torch.device("cuda", self.device)
        adc = out if out is not None else torch.empty((n_frames, self.n_mics, self.n_samples), dtype=torch.uint8, device=dev)
        heads = torch.empty(n_frames, dtype=torch.int32, device=dev)
        cells = torch.empty(n_frames, dtype=torch.int32, device=dev)
        st = stream if stream is not None else torch.cuda.current_stream(dev)
        L.check(self.lib.at_synth_device(self.ctx, seed, flags, first_frame, n_frames, adc.data_ptr(), heads.data_ptr(),
                                         cells.data_ptr(), C.c_void_p(st.cuda_stream)))
        return adc, heads, cells

    def microbench(self, which):
        g, mhz = C.c_double(), C.c_double()
        L.check(self.lib.at_microbench(self.ctx, L.UBENCH[which], C.byref(g), C.byref(mhz)))
        return g.value, mhz.value

    def kernel_launches(self):
        return int(self.lib.at_kernel_launches())


class Stream:
    """Device-resident capture state of many arrays (at_stream): push sample blocks, get onsets + captured rings."""

    def __init__(self, loc, n_arrays):
        self.loc, self.n_arrays = loc, n_arrays
        self.h = C.c_void_p()
        L.check(loc.lib.at_stream_create(loc.ctx, n_arrays, C.byref(self.h)))

    def close(self):
        if self.h.value:
            self.loc.lib.at_stream_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reset(self):
        import torch
        L.check(self.loc.lib.at_stream_reset(self.h, C.c_void_p(torch.cuda.current_stream().cuda_stream)))

    def push(self, samples, want_frames=True, out=None):
        """samples: torch.uint8 CUDA [A, ticks, mics].  Returns dict(fired [A] int32 (1-based tick or -1),
        frames [A, mics, N] uint8 ring order, heads [A] int32); frames/heads rows are valid where fired > 0."""
        import torch
        assert samples.is_cuda and samples.dtype == torch.uint8 and samples.is_contiguous()
        A, ticks = samples.shape[0], samples.shape[1]
        assert A == self.n_arrays
        res = {} if out is None else out
        dev = samples.device
        if "fired" not in res:
            res["fired"] = torch.empty(A, dtype=torch.int32, device=dev)
        if want_frames and "frames" not in res:
            res["frames"] = torch.zeros((A, self.loc.n_mics, self.loc.n_samples), dtype=torch.uint8, device=dev)
            res["heads"] = torch.zeros(A, dtype=torch.int32, device=dev)
        st = torch.cuda.current_stream(dev)
        L.check(self.loc.lib.at_stream_push(self.h, samples.data_ptr(), ticks, res["fired"].data_ptr(),
                                            res["frames"].data_ptr() if want_frames else None,
                                            res["heads"].data_ptr() if want_frames else None, C.c_void_p(st.cuda_stream)))
        return res


def _host_ptr(a):
    if isinstance(a, np.ndarray):
        assert a.flags["C_CONTIGUOUS"]
        return a.ctypes.data
    assert not a.is_cuda and a.is_contiguous()   # torch CPU tensor (pinned or not)
    return a.data_ptr()


class _DropIn:
    """The reference-named entry points on numpy structured scalars (RING_DT / BUFFER_DT / CORR_DT)."""

    def __init__(self):
        self._lib = None

    @property
    def lib(self):
        if self._lib is None:
            self._lib = L.load()
        return self._lib

    @staticmethod
    def _p(rec):
        return C.c_void_p(rec.ctypes.data)

    def new_ring(self):
        r = np.zeros(1, RING_DT)
        self.lib.rolling_buffer_init(self._p(r))
        return r

    def rolling_buffer_init(self, r): self.lib.rolling_buffer_init(self._p(r))
    def rolling_buffer_push(self, r, s): self.lib.rolling_buffer_push(self._p(r), int(s))
    def rolling_buffer_write_out(self, r, b): self.lib.rolling_buffer_write_out(self._p(r), self._p(b))
    def rolling_buffer_get_incoming_power(self, r): return int(self.lib.rolling_buffer_get_incoming_power(self._p(r)))
    def rolling_buffer_get_outgoing_power(self, r): return int(self.lib.rolling_buffer_get_outgoing_power(self._p(r)))
    def buffer_normalize_range(self, b): self.lib.buffer_normalize_range(self._p(b))
    def buffer_window(self, b): self.lib.buffer_window(self._p(b))
    def correlations_init(self, c, a, b): self.lib.correlations_init(self._p(c), self._p(a), self._p(b))
    def correlations_average(self, e, n): self.lib.correlations_average(self._p(e), self._p(n))
    def set_time_us(self, t): self.lib.at_set_time_us(int(t))

    def microphones_init(self):
        self.lib.microphones_init()
        pt = (C.c_float * 2)
        return np.array([list(pt.in_dll(self.lib, n)) for n in ("mic_a_location", "mic_b_location", "mic_c_location")],
                        np.float32)


dropin = _DropIn()
