"""The reference's OWN orchestration code as the test driver.

oracle/sc_host.c includes the reference's src/sample_compute.h unmodified (capture loop, onset gate,
write-out, <<8, window, three correlations_init, TDOA gate, correlations_average; sample_compute.h:45-150) and
feeds it a recorded ADC triple stream.  oracle/Makefile links it twice: _ref/sc_ref against the reference's own
buffer.c / rolling_buffer.c / correlations.c, _ref/sc_b200 against libat_b200.so's drop-in symbols -- nothing
else differs.  tests/golden/sc_protothread.npz holds the stream and sc_ref's transcript
(tests/golden/make_sc_golden.py).

 * CPU: sc_ref (when built here) reproduces the committed transcript; our restated oracle, walked over the same
   stream, reproduces every gated event of the transcript (tick, fresh lags, checksums of the three
   post-Gaussian curves) -- a pin of the oracle on the reference's real call order.
 * GPU: sc_b200's transcript is identical to the golden one, byte for byte: fresh and averaged lags and the
   checksums of all six correlations_t for every gated frame.
"""
import os
import subprocess
import tempfile

import numpy as np
import pytest

from oracle_bindings import N, ROOT

SC_REF = os.path.join(ROOT, "oracle", "_ref", "sc_ref")
SC_B200 = os.path.join(ROOT, "oracle", "_ref", "sc_b200")


@pytest.fixture(scope="module")
def sc_golden():
    g = np.load(os.path.join(ROOT, "tests", "golden", "sc_protothread.npz"))
    return g["stream"], str(g["transcript"])


def run_sc(binary, stream):
    with tempfile.NamedTemporaryFile(suffix=".bin") as f:
        stream.tofile(f.name)
        r = subprocess.run([binary, f.name], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.returncode, r.stderr[-2000:])
    return r.stdout


def parse(transcript):
    ev = []
    for ln in transcript.splitlines():
        if not ln.startswith("event"):
            continue
        w = ln.split()
        ev.append(dict(tick=int(w[3]), new=[int(v) for v in w[5:8]], avg=[int(v) for v in w[9:12]],
                       sum_new=[int(v) for v in w[13:16]], sum_avg=[int(v) for v in w[17:20]]))
    return ev


def checksum(curve):
    s = 0
    for v in curve:                       # sc_host.c: s = s * 31 + (uint64)c, modulo 2^64
        s = (s * 31 + (int(v) & 0xFFFFFFFFFFFFFFFF)) & 0xFFFFFFFFFFFFFFFF
    return s


def test_sc_ref_reproduces_golden_transcript(sc_golden):
    if not os.path.exists(SC_REF):
        pytest.skip("oracle/_ref/sc_ref not built (needs /root/reference at build time)")
    stream, transcript = sc_golden
    assert run_sc(SC_REF, stream) == transcript


def test_oracle_walk_matches_reference_protothread(oracle, sc_golden):
    """Our restatement (ato_capture + ato_localize) over the stream == the events the reference's protothread saw."""
    stream, transcript = sc_golden
    events = parse(transcript)
    assert len(events) >= 5
    got, pos = [], 0
    while pos < stream.shape[0]:
        rings = np.zeros((3, N), np.int16)
        head = np.zeros(1, np.int32)
        fired = oracle.lib.ato_capture(np.ascontiguousarray(stream[pos:]).reshape(-1), stream.shape[0] - pos, 3, 10,
                                       rings.reshape(-1), head)
        if fired < 0:
            break
        # the triple that fired the gate is still in dma_sample_array when the next capture starts
        # (sample_compute.h:67-69 reads it again before the first busy_wait_until)
        pos += fired - 1
        res = oracle.localize(rings.astype(np.uint8)[None], heads=head.copy())
        lags = res["lags"][0]
        if int((lags.astype(np.int64) ** 2).sum()) > 4:           # sample_compute.h:124-134
            got.append(dict(tick=pos + 1, new=lags.tolist(), sum_new=[checksum(c) for c in res["corr"][0]]))
    assert [g["tick"] for g in got] == [e["tick"] for e in events]
    for g, e in zip(got, events):
        assert g["new"] == e["new"]
        assert g["sum_new"] == e["sum_new"]


@pytest.mark.gpu
def test_reference_protothread_on_dropin_library(sc_golden):
    """sample_compute.h linked against libat_b200.so prints exactly what it prints with the reference objects."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    if not os.path.exists(SC_B200):
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "_ref/sc_b200"], check=False)
    assert os.path.exists(SC_B200), "oracle/_ref/sc_b200 missing: run __graft_entry__.build() where /root/reference exists"
    stream, transcript = sc_golden
    got = run_sc(SC_B200, stream)
    assert parse(got) == parse(transcript)
    assert got == transcript
    if os.path.exists(SC_REF):
        assert run_sc(SC_REF, stream) == got
