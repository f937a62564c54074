"""CPU: differential test of the oracle restatement against the reference's own objects
(oracle/_ref/libat_ref.so, built from /root/reference/src unmodified).  Skipped when oracle/_ref
was never built; the golden-vector test covers that case."""
import numpy as np

from frames import burst_frames
from oracle_bindings import CORR_DT, RING_DT, N


def test_random_frames_bit_exact(oracle, ref):
    adc, _ = burst_frames(300, seed=99)
    corr = np.zeros((adc.shape[0], 3), CORR_DT)
    lags = np.zeros((adc.shape[0], 3), np.int32)
    ref.ref_localize_frames(adc.reshape(-1), adc.shape[0], lags.ctypes.data, corr.ctypes.data, 2, 5)
    r = oracle.localize(adc)
    assert (r["lags"] == lags).all()
    assert (r["corr"] == corr["correlations"]).all()
    assert (corr["last_update"] == 5).all()


def test_ring_push_and_powers(oracle, ref):
    import ctypes as C
    from oracle_bindings import load_oracle
    rng = np.random.default_rng(3)
    samples = rng.integers(0, 256, 2500).astype(np.int16)
    ring = np.zeros(1, RING_DT)
    ref.rolling_buffer_init(ring.ctypes.data)

    class R(C.Structure):
        _fields_ = [("head", C.c_int32), ("full", C.c_int32), ("n_bits", C.c_int32),
                    ("in_pow", C.c_int64), ("in_tot", C.c_int64), ("out_pow", C.c_int64), ("out_tot", C.c_int64),
                    ("buf", C.c_void_p)]
    lib = load_oracle()
    store = np.zeros(N, np.int16)
    r = R()
    lib.ato_ring_init(C.byref(r), store.ctypes.data_as(C.c_void_p), 10)
    lib.ato_ring_incoming.restype = lib.ato_ring_outgoing.restype = C.c_int64
    for i, s in enumerate(samples):
        ref.rolling_buffer_push(ring.ctypes.data, int(s))
        lib.ato_ring_push(C.byref(r), C.c_int16(int(s)))
        if i % 97 == 0 or i == len(samples) - 1:
            assert r.head == ring["head"][0] and bool(r.full) == bool(ring["is_full"][0])
            assert lib.ato_ring_incoming(C.byref(r)) == ref.rolling_buffer_get_incoming_power(ring.ctypes.data)
            assert lib.ato_ring_outgoing(C.byref(r)) == ref.rolling_buffer_get_outgoing_power(ring.ctypes.data)
    assert (store == ring["buffer"][0]).all()


def test_ref_fast_build_same_lags(ref):
    """The -O3/AVX2 timing build of the reference gives the same lags as the parity build."""
    from oracle_bindings import load_ref
    fast = load_ref(fast=True)
    adc, _ = burst_frames(64, seed=11)
    a = np.zeros((64, 3), np.int32); b = np.zeros((64, 3), np.int32)
    ref.ref_localize_frames(adc.reshape(-1), 64, a.ctypes.data, None, 1, 0)
    fast.ref_localize_frames(adc.reshape(-1), 64, b.ctypes.data, None, 4, 0)
    assert (a == b).all()
