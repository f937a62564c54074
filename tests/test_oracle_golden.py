"""CPU: the oracle restatement (oracle/at_oracle.c) against the committed golden vectors that
tests/golden/make_golden.py produced from the reference's own compiled objects.  Bit-exact."""
import ctypes as C

import numpy as np

from oracle_bindings import CELLS, L, N, NL, Oracle


def test_layout_facts(golden):
    lay = golden["layout"]
    # sizeof/offsetof of buffer_t, rolling_buffer_t, correlations_t; MAX_SHIFT, lags, BUFFER_SIZE
    assert lay.tolist() == [2056, 2048, 2096, 8, 40, 42, 760, 744, 752, 46, 93, 1024]


def test_kat_lags(golden):
    names = [str(x) for x in golden["kat_names"]]
    best = golden["best_shift"]
    assert best[names.index("silence")].tolist() == [-46, -46, -46]      # strict '>' keeps the first lag
    assert best[names.index("impulse")].tolist() == [11, -20, -31]
    assert best[names.index("burst_dB7_dC-12")].tolist() == [7, -12, -19]
    k = names.index("impulse")
    assert golden["after_window"][k, 0, 500] == 25564 and golden["power"][k, 0] == 10000
    assert golden["corr"][k, 0, 11 + L] == 654412864
    k = names.index("fullscale_square")   # int16 wrap: -127/+128 -> -32512/-32768
    assert set(np.unique(golden["after_shift"][k, 0]).tolist()) == {-32512, -32768}


def test_stages_match_reference(oracle, golden):
    adc = golden["adc"]
    tab = oracle.window
    for k in range(adc.shape[0]):
        for m in range(3):
            ring = adc[k, m].astype(np.int16)
            out = np.zeros(N, np.int16)
            pw = oracle.lib.ato_dc_remove(ring, 10, 0, out)
            assert pw == golden["power"][k, m]
            assert (out == golden["after_dc"][k, m]).all()
            oracle.lib.ato_shift8(out, N)
            assert (out == golden["after_shift"][k, m]).all()
            oracle.lib.ato_window(out, 10, tab, 10)
            assert (out == golden["after_window"][k, m]).all()


def test_localize_matches_reference(oracle, golden):
    r = oracle.localize(golden["adc"], want_raw=True)
    assert (r["lags"] == golden["best_shift"]).all()
    assert (r["corr"] == golden["corr"]).all()          # post-Gaussian curves, bit-exact
    # ring order with arbitrary heads gives the same results (rolling_buffer.c:48-62)
    adc, heads = golden["adc"], golden["heads"]
    rolled = np.stack([np.roll(adc[k], int(heads[k]), axis=-1) for k in range(adc.shape[0])])
    r2 = oracle.localize(rolled, heads=heads)
    assert (r2["lags"] == golden["best_shift"]).all() and (r2["corr"] == golden["corr"]).all()
    # threads do not change anything
    r3 = oracle.localize(golden["adc"], nthreads=4)
    assert (r3["corr"] == golden["corr"]).all() and (r3["cell"] == r["cell"]).all()


def test_average_chain(oracle, golden):
    n_kat = len(golden["kat_names"])
    est = np.zeros((3, NL), np.int64); best = np.zeros(3, np.int32)
    last = [C.c_uint64(0) for _ in range(3)]
    for j, k in enumerate(range(n_kat, golden["adc"].shape[0])):
        for p in range(3):
            oracle.lib.ato_average(est[p], best[p:p + 1], C.byref(last[p]), np.ascontiguousarray(golden["corr"][k, p]),
                                   L, int(golden["avg_times"][j]))
    assert (est == golden["avg_est"]).all()
    assert (best == golden["avg_best"]).all()
    assert [x.value for x in last] == golden["avg_last"].tolist()


def test_capture_gate(oracle, golden):
    stream = golden["cap_stream"]
    rings = np.zeros((3, N), np.int16)
    head = np.zeros(1, np.int32)
    fired = oracle.lib.ato_capture(stream.reshape(-1), stream.shape[0], 3, 10, rings.reshape(-1), head)
    assert fired == int(golden["cap_fired"]) and fired > 1024
    assert head[0] == golden["cap_head"][0]
    assert (rings == golden["cap_ring"]).all()
    quiet = np.full((3000, 3), 128, np.uint8)
    assert oracle.lib.ato_capture(quiet.reshape(-1), 3000, 3, 10, rings.reshape(-1), head) == -1 == int(golden["cap_quiet_fired"])


def test_geometry_and_lut(oracle, golden):
    assert (oracle.mics().reshape(-1) == golden["mics"]).all()       # microphones.c, float32 bit-exact
    np.testing.assert_allclose(oracle.mics(), [[-0.088096, 0.05], [0.043904, 0.05], [0.044192, -0.1]], atol=1e-6)
    lut = oracle.reference_lut().astype(int) - L
    # SURVEY 3.3 probe facts (independent float32 restatement): index ranges and distinct triples
    assert [int(abs(lut[p]).max()) for p in range(3)] == [17, 27, 19]
    assert len(set(map(tuple, lut.T.tolist()))) == 2469
    assert lut.shape == (3, CELLS)
    centre = 50 * 101 + 50
    assert abs(lut[:, centre]).max() <= 1     # straight above the array: near-zero lags


def test_heatmap_definition(oracle):
    rng = np.random.default_rng(5)
    corr = rng.integers(-10**9, 10**12, (3, NL)).astype(np.int64)
    lut = oracle.reference_lut()
    like = corr[0][lut[0]] + corr[1][lut[1]] + corr[2][lut[2]]
    hi = np.zeros(1, np.int64); cell = np.zeros(1, np.int32); cls = np.zeros(CELLS, np.uint8)
    oracle.lib.ato_heatmap(corr.reshape(-1), lut.reshape(-1), 3, CELLS, L, hi.ctypes.data, cell.ctypes.data, cls.ctypes.data)
    assert hi[0] == like.max() and cell[0] == int(np.argmax(like))
    top = int(like.max())
    exp = np.where(like >= (top * 63) >> 6, 15, np.where(like >= (top * 31) >> 5, 3,
                   np.where(like >= (top * 15) >> 4, 8, np.where(like >= (top * 7) >> 3, 5, 0))))
    assert (cls == exp).all()


# ---------------------------------------------------------------- a19 / a20 pinned to the reference's own vga_heatmap.h
def test_lut_is_the_reference_lut(oracle, golden_hm):
    """vga_init_heatmap (vga_heatmap.h:48-93), compiled unmodified, built golden_hm['lut']; the restated ato_lut_build and
    ato_mics_triangle must reproduce it and the reference's microphone coordinates exactly."""
    assert golden_hm["dims"].tolist() == [101, 101, NL]
    assert golden_hm["colors"].tolist() == [15, 3, 8, 5, 0]               # WHITE GREEN RED BLUE BLACK (vga16_graphics.h:31-34)
    assert (oracle.mics().view(np.uint32) == golden_hm["mics"].view(np.uint32)).all()
    assert (oracle.reference_lut() == golden_hm["lut"]).all()


def test_classes_are_the_reference_classes(oracle, golden_hm):
    """vga_draw_heatmap (vga_heatmap.h:95-135) coloured every cell for 41 curve triples (golden frames, the EMA estimate,
    edge cases: silence, negative maxima, ties, peaks outside the LUT); ato_heatmap must give the same class per cell,
    and its highest_L / first cell must be consistent with the reference's look-up table."""
    lut = np.ascontiguousarray(golden_hm["lut"])
    for k in range(golden_hm["curves"].shape[0]):
        corr = np.ascontiguousarray(golden_hm["curves"][k])
        hi = np.zeros(1, np.int64); cell = np.zeros(1, np.int32); cls = np.zeros(CELLS, np.uint8)
        oracle.lib.ato_heatmap(corr.reshape(-1), lut.reshape(-1), 3, CELLS, L, hi.ctypes.data, cell.ctypes.data, cls.ctypes.data)
        assert (cls == golden_hm["classes"][k]).all(), f"curve set {k}"
        like = corr[0][lut[0]] + corr[1][lut[1]] + corr[2][lut[2]]
        assert hi[0] == like.max() and cell[0] == int(np.argmax(like))


def test_host_generator_self_consistent(oracle):
    """oracle/synth_host.cpp tabulates the source sequence per frame; the bytes must equal the plain per-byte evaluation of
    the generator header (at_synth.h), known-answer frames and random ring heads included, whatever the thread count."""
    a, ha, ca = oracle.synth(600, flags=2 | 4, nthreads=3)
    oracle.lib.ato_synth_plain(1)
    try:
        b, hb, cb = oracle.synth(600, flags=2 | 4, nthreads=1)
    finally:
        oracle.lib.ato_synth_plain(0)
    assert (a == b).all() and (ha == hb).all() and (ca == cb).all()
    assert (a[0, 0] == 128).all() and (a[0, 1] == 131).all()          # frame 0 = the silence KAT
    c, _, _ = oracle.synth(100, first_frame=500, flags=2 | 4)
    assert (c == a[500:]).all()                                       # counter-based: any sub-range gives the same frames


def test_points_lut_reduces_to_the_reference_lut(oracle, golden_hm):
    """The 3-D form of the lag look-up table (arbitrary candidate positions, ato_lut_build_points) fed with the reference's
    own candidate set -- the 101 x 101 plane grid projected onto the 1.2 m sphere, in the reference's float32 arithmetic
    (vga_heatmap.h:52-60) -- must reproduce the reference's table bit for bit."""
    f = np.float32
    y, x = np.divmod(np.arange(CELLS), 101)
    xm = (x - 50).astype(f) / f(24.0); ym = (50 - y).astype(f) / f(24.0); zm = np.full(CELLS, 1.2, f)
    k = f(1.2) / np.sqrt(zm * zm + xm * xm + ym * ym, dtype=f)       # hypot3f(z, x, y): same summation order
    pts = np.ascontiguousarray(np.stack([xm * k, ym * k, zm * k], 1), f)
    idx = np.zeros((3, CELLS), np.uint8)
    oracle.lib.ato_lut_build_points(oracle.mics().reshape(-1), 3, L, 50000.0, 343.0, pts.reshape(-1), CELLS, idx.reshape(-1))
    assert (idx == golden_hm["lut"]).all()
