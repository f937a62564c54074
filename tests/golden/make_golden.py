#!/usr/bin/env python3
"""Generate tests/golden/ref_vectors.npz by driving the REFERENCE's own objects
(oracle/_ref/libat_ref.so, compiled unmodified from /root/reference/src by oracle/Makefile).

Run here (the container with /root/reference):  python tests/golden/make_golden.py
The .npz is committed; the GPU box has no /root/reference and uses these vectors.
"""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from frames import burst_frames, kat_frames  # noqa: E402
from oracle_bindings import BUFFER_DT, CORR_DT, RING_DT, load_ref  # noqa: E402


def main():
    ref = load_ref()
    assert ref is not None, "build oracle/_ref first (make -C oracle)"
    names, kats = kat_frames()
    bursts, delays = burst_frames(24, seed=20261018)
    adc = np.concatenate([kats, bursts])
    K = adc.shape[0]
    rng = np.random.default_rng(7)
    heads = rng.integers(0, 1024, K).astype(np.int32)
    heads[: len(names)] = 0
    heads[len(names) + 1] = 8          # aligned non-zero head
    heads[len(names) + 2] = 1023

    out = dict(adc=adc, heads=heads, kat_names=np.array(names), burst_delays=delays)
    # (a) chronological frames through every stage
    dc = np.zeros((K, 3, 1024), np.int16); sh = np.zeros_like(dc); wn = np.zeros_like(dc)
    pw = np.zeros((K, 3), np.int64); corr = np.zeros((K, 3), CORR_DT)
    ref.ref_set_time(424242)
    for k in range(K):
        ref.ref_frame_stages(adc[k].reshape(-1), 0, dc[k].ctypes.data, pw[k].ctypes.data, sh[k].ctypes.data,
                             wn[k].ctypes.data, corr[k].ctypes.data)
    out.update(after_dc=dc, after_shift=sh, after_window=wn, power=pw,
               corr=corr["correlations"].copy(), best_shift=corr["best_shift"].copy(),
               last_update=corr["last_update"].copy())
    # (b) same frames entering through a ring with a non-zero head must give identical results
    corr_h = np.zeros((K, 3), CORR_DT)
    for k in range(K):
        ref.ref_frame_stages(adc[k].reshape(-1), int(heads[k]), None, None, None, None, corr_h[k].ctypes.data)
    assert (corr_h["correlations"] == corr["correlations"]).all()

    # (c) temporal average: chain the burst frames' curves into one estimate per pair
    est = np.zeros(3, CORR_DT)
    times = []
    t = 1_000_000
    for k in range(len(names), K):
        fresh = corr[k].copy()
        t += int(rng.integers(20_000, 900_000))
        times.append(t)
        ref.ref_set_time(t)
        for p in range(3):
            ref.correlations_average(est[p:p + 1].ctypes.data, fresh[p:p + 1].ctypes.data)
    out.update(avg_times=np.array(times, np.uint64), avg_est=est["correlations"].copy(),
               avg_best=est["best_shift"].copy(), avg_last=est["last_update"].copy())

    # (d) capture loop with onset gate on a recorded triple stream
    n = 6000
    stream = np.clip(128 + rng.normal(0, 1.5, (n, 3)), 0, 255)
    burst = np.convolve(rng.normal(0, 70, 700), np.ones(3) / 3, "same") * np.hanning(700)
    for m, d in enumerate((0, 9, -6)):
        stream[2500 + d:3200 + d, m] += burst
    stream = np.clip(np.round(stream), 0, 255).astype(np.uint8)
    rings = np.zeros(3, RING_DT)
    fired = ref.ref_capture(stream.reshape(-1), n, rings.ctypes.data)
    out.update(cap_stream=stream, cap_fired=np.int64(fired), cap_head=rings["head"].copy(),
               cap_ring=rings["buffer"].copy(),
               cap_sums=np.stack([rings[k] for k in ("incoming_power", "incoming_total", "outgoing_power", "outgoing_total")]))
    quiet = np.full((3000, 3), 128, np.uint8)
    quiet_rings = np.zeros(3, RING_DT)
    out.update(cap_quiet_fired=np.int64(ref.ref_capture(quiet.reshape(-1), 3000, quiet_rings.ctypes.data)))

    # (e) geometry + layout facts
    mics = np.zeros(6, np.float32); ref.ref_mics(mics)
    lay = np.zeros(12, np.int64); ref.ref_layout(lay)
    out.update(mics=mics, layout=lay)
    path = os.path.join(HERE, "ref_vectors.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", K, "frames; capture fired at", fired)
    print("KAT lags:", {nm: corr["best_shift"][i].tolist() for i, nm in enumerate(names)})


if __name__ == "__main__":
    main()
