#!/usr/bin/env python3
"""Generate tests/golden/sc_protothread.npz: a recorded ADC triple stream and the transcript that the
REFERENCE's own protothread_sample_and_compute (src/sample_compute.h, compiled unmodified into
oracle/_ref/sc_ref together with the reference's buffer.c / rolling_buffer.c / correlations.c by
oracle/Makefile) prints for it -- one line per gated frame with the fresh and averaged lags and a
checksum of every correlations_t.

Run here (the container with /root/reference):  python tests/golden/make_sc_golden.py
The .npz is committed; the GPU box has no /root/reference and checks libat_b200.so's drop-in symbols
(oracle/_ref/sc_b200 = the same sample_compute.h linked against libat_b200.so) against this transcript.
"""
import os
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SC_REF = os.path.join(ROOT, "oracle", "_ref", "sc_ref")


def make_stream(n_ticks=30000, seed=5):
    """Background noise with bursts at random times and per-microphone delays (uint8 triples A,B,C)."""
    rng = np.random.default_rng(seed)
    x = 128 + rng.normal(0, 1.5, (n_ticks, 3))
    t = 1500
    while t + 800 < n_ticks:
        burst = np.convolve(rng.normal(0, 50, 600), np.ones(3) / 3, "same") * np.hanning(600)
        for m, d in enumerate(rng.integers(-12, 13, 3)):
            x[t + d:t + d + 600, m] += burst
        t += int(rng.integers(2500, 6000))
    return np.clip(np.round(x), 0, 255).astype(np.uint8)


def run_sc(binary, stream):
    with tempfile.NamedTemporaryFile(suffix=".bin") as f:
        stream.tofile(f.name)
        r = subprocess.run([binary, f.name], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    return r.stdout


def main():
    assert os.path.exists(SC_REF), "build oracle/_ref/sc_ref first (make -C oracle)"
    stream = make_stream()
    transcript = run_sc(SC_REF, stream)
    events = [ln for ln in transcript.splitlines() if ln.startswith("event")]
    assert len(events) >= 5 and transcript.rstrip().endswith("END ticks=%d" % stream.shape[0])
    np.savez_compressed(os.path.join(HERE, "sc_protothread.npz"), stream=stream, transcript=np.array(transcript))
    print("wrote sc_protothread.npz: %d ticks, %d gated events" % (stream.shape[0], len(events)))


if __name__ == "__main__":
    main()
