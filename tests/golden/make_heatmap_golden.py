#!/usr/bin/env python3
"""Generate tests/golden/ref_heatmap.npz from the REFERENCE's own presentation code: oracle/_ref/libat_ref_hm.so is
src/components/vga/vga.h (vga_init_heatmap / vga_draw_heatmap, vga_heatmap.h:48-135) compiled unmodified on the host
by oracle/Makefile (oracle/hm_host.c supplies the SDK shim and no-op VGA primitives).

Contents: the lag look-up table the reference builds (heat_idx_ab/ac/bc), its microphone coordinates, and the colour
class of every cell after vga_draw_heatmap for a set of curve triples: the golden frames' post-Gaussian curves
(ref_vectors.npz, themselves produced by the reference's correlations.c), the averaged estimate, and synthetic edge
cases (all zero, all negative, ties, int64-large values, a single spike).

Run here (the container with /root/reference):  python tests/golden/make_heatmap_golden.py
The .npz is committed; the GPU box has no /root/reference and uses these vectors.
"""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))


def load_hm():
    path = os.path.join(ROOT, "oracle", "_ref", "libat_ref_hm.so")
    assert os.path.exists(path), "build oracle/_ref first (make -C oracle)"
    lib = C.CDLL(path)
    lib.hm_dims.argtypes = [C.POINTER(C.c_int)] * 3 + [C.POINTER(C.c_int)]
    lib.hm_init.argtypes = [C.c_void_p, C.c_void_p]
    lib.hm_draw.argtypes = [C.c_void_p, C.c_void_p]
    lib.hm_draw.restype = C.c_int
    return lib


def edge_curves(nl, rng):
    out = []
    out.append(np.zeros((3, nl), np.int64))                                   # silence: every cell ties at 0
    out.append(-rng.integers(1, 10**9, (3, nl)).astype(np.int64))              # all negative: thresholds of a negative maximum
    c = np.zeros((3, nl), np.int64); c[:, 46] = 5; out.append(c)               # tiny spike at lag 0 (threshold rounding)
    c = np.full((3, nl), 1000, np.int64); out.append(c)                        # flat positive: every cell WHITE
    c = rng.integers(-2**40, 2**40, (3, nl)).astype(np.int64); out.append(c)   # large magnitudes
    c = rng.integers(0, 64, (3, nl)).astype(np.int64); out.append(c)           # small integers: many ties
    c = np.zeros((3, nl), np.int64); c[0, 46 + 7] = 10**12; c[1, 46 - 12] = 10**12; c[2, 46 - 19] = 10**12; out.append(c)
    c = np.zeros((3, nl), np.int64); c[0, 0] = 7; c[1, nl - 1] = 9; c[2, 5] = -3; out.append(c)   # peaks outside the LUT's range
    return out


def main():
    lib = load_hm()
    w, h, nl = C.c_int(), C.c_int(), C.c_int()
    colors = (C.c_int * 5)()
    lib.hm_dims(C.byref(w), C.byref(h), C.byref(nl), colors)
    W, H, NL = w.value, h.value, nl.value
    mic = np.zeros(6, np.float32); lut = np.zeros((3, H, W), np.uint8)
    lib.hm_init(mic.ctypes.data, lut.ctypes.data)
    ref = np.load(os.path.join(HERE, "ref_vectors.npz"))
    rng = np.random.default_rng(20261018)
    curves = [ref["corr"][k] for k in range(ref["corr"].shape[0])]            # per-frame post-Gaussian curves
    curves.append(ref["avg_est"])                                              # the EMA estimate the firmware actually draws
    curves += edge_curves(NL, rng)
    curves = np.ascontiguousarray(np.stack(curves), np.int64)
    classes = np.zeros((curves.shape[0], H * W), np.uint8)
    for k in range(curves.shape[0]):
        bad = lib.hm_draw(curves[k].ctypes.data, classes[k].ctypes.data)
        assert bad == 0, f"fillRect log disagrees with heat_colors on {bad} cells"
    out = dict(dims=np.array([W, H, NL], np.int32), colors=np.array(list(colors), np.int32), mics=mic.reshape(3, 2),
               lut=lut.reshape(3, H * W), curves=curves, classes=classes)
    path = os.path.join(HERE, "ref_heatmap.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items()}, "bytes", os.path.getsize(path))
    print("index ranges:", [(int(lut[p].min()) - 46, int(lut[p].max()) - 46) for p in range(3)],
          "distinct triples:", len(np.unique(lut.reshape(3, -1).T, axis=0)))


if __name__ == "__main__":
    main()
