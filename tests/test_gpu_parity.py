"""GPU: parity of the CUDA path (through the C ABI) against the oracle and the golden vectors
produced by the reference's own objects.  Integer/index results bit-exact; the two float-derived
products (post-Gaussian curves, plane coordinates) are bit-exact too by construction (host-libm
factor table, IEEE round-to-nearest ops) and are asserted with tolerance 0."""
import ctypes as C

import numpy as np
import pytest

from frames import burst_frames, kat_frames
from oracle_bindings import CELLS, CORR_DT, HALF_H, HALF_W, HEIGHT, PX_PER_M, RATE_HZ, SPEED, L, N, NL, Oracle

pytestmark = pytest.mark.gpu

ALL = ("lags", "corr", "raw", "cell", "highest", "xy", "gate", "classes", "windowed", "power")
KERNELS = ("imad", "imma", "umma")


def _torch():
    import torch
    return torch


def make_loc(kernel, **kw):
    """Localizer pinned to one kernel variant; skip when that variant has no instantiation for the shape."""
    import audio_triangulation_b200 as at
    torch = _torch()
    loc = at.Localizer(device=0, kernel=kernel, **kw)
    probe = torch.zeros((1, loc.n_mics, loc.n_samples), dtype=torch.uint8, device="cuda")
    try:
        loc.localize_device(probe, want=("lags",))
        torch.cuda.synchronize()
    except at.AtError as e:
        if e.code == -1:
            pytest.skip(str(e))
        raise
    return loc


def run_device(loc, adc_np, heads_np=None, want=ALL, struct_corr=False):
    torch = _torch()
    adc = torch.from_numpy(np.ascontiguousarray(adc_np)).cuda()
    heads = None if heads_np is None else torch.from_numpy(np.ascontiguousarray(heads_np, np.int32)).cuda()
    try:
        res = loc.localize_device(adc, heads, want=want, struct_corr=struct_corr)
    except Exception as e:
        if "no instantiation" in str(e) or "IMMA kernel has no" in str(e):
            pytest.skip(str(e))
        raise
    torch.cuda.synchronize()
    return {k: v.cpu().numpy() for k, v in res.items()}


def oracle_classes(oracle, corr):
    out = np.zeros((corr.shape[0], CELLS), np.uint8)
    for f in range(corr.shape[0]):
        oracle.lib.ato_heatmap(np.ascontiguousarray(corr[f]).reshape(-1), oracle.lut.reshape(-1), 3, CELLS, L, None, None,
                               out[f].ctypes.data)
    return out


# ---------------------------------------------------------------- geometry
def test_geometry_and_lut_bit_exact(loc, oracle, golden):
    assert (loc.mics().reshape(-1) == golden["mics"]).all()          # vs reference microphones_init
    assert (loc.lut() == oracle.reference_lut()).all()               # vs restated vga_init_heatmap
    assert loc.n_cells == CELLS and loc.n_lags == NL


# ---------------------------------------------------------------- golden vectors, every product
@pytest.mark.parametrize("kernel", KERNELS)
def test_golden_frames_every_output(kernel, oracle, golden):
    loc = make_loc(kernel)
    adc = golden["adc"]
    r = run_device(loc, adc)
    o = oracle.localize(adc, want_raw=True)
    assert (r["lags"] == golden["best_shift"]).all()
    assert (r["windowed"] == golden["after_window"]).all()
    assert (r["power"] == golden["power"]).all()
    assert (r["raw"] == o["raw"]).all()
    assert (r["corr"] == golden["corr"]).all()                       # post-Gaussian, tolerance 0
    assert (r["cell"] == o["cell"]).all() and (r["highest"] == o["highest"]).all()
    gate = (golden["best_shift"].astype(np.int64) ** 2).sum(1) > 4   # sample_compute.h:124-134
    assert (r["gate"] == gate).all()
    cx, cy = o["cell"] % 101, o["cell"] // 101
    xy = np.stack([(cx - HALF_W).astype(np.float32) / np.float32(PX_PER_M),
                   (HALF_H - cy).astype(np.float32) / np.float32(PX_PER_M)], 1)
    assert (r["xy"] == xy).all()                                     # tolerance 0 (IEEE float division)
    assert (r["classes"] == oracle_classes(oracle, o["corr"])).all()


@pytest.mark.parametrize("kernel", KERNELS)
def test_struct_layout_output(kernel, golden):
    import audio_triangulation_b200 as at
    loc = make_loc(kernel)
    at.load().at_set_time_us(987654321)
    r = run_device(loc, golden["adc"], want=("corr", "lags"), struct_corr=True)
    at.load().at_set_time_us(2**64 - 1)
    rec = np.ascontiguousarray(r["corr"]).view(CORR_DT).reshape(-1, 3)
    assert (rec["correlations"] == golden["corr"]).all()
    assert (rec["best_shift"] == golden["best_shift"]).all()
    assert (rec["last_update"] == 987654321).all() and (rec["_pad"] == 0).all()


@pytest.mark.parametrize("kernel", KERNELS)
def test_ring_heads(kernel, golden):
    loc = make_loc(kernel)
    adc, heads = golden["adc"], golden["heads"]
    rolled = np.stack([np.roll(adc[k], int(heads[k]), axis=-1) for k in range(adc.shape[0])])
    r = run_device(loc, rolled, heads, want=("lags", "corr", "windowed", "power"))
    assert (r["lags"] == golden["best_shift"]).all() and (r["corr"] == golden["corr"]).all()
    assert (r["windowed"] == golden["after_window"]).all() and (r["power"] == golden["power"]).all()


# ---------------------------------------------------------------- randomized differential test
@pytest.mark.parametrize("kernel", KERNELS)
def test_random_frames_vs_oracle(kernel, oracle):
    loc = make_loc(kernel)
    adc, _ = burst_frames(1500, seed=4242)
    r = run_device(loc, adc, want=("lags", "corr", "cell", "highest", "raw"))
    o = oracle.localize(adc, want_raw=True, nthreads=8)
    for k in ("lags", "raw", "corr", "cell", "highest"):
        assert (r[k] == o[k]).all(), k


@pytest.mark.parametrize("kernel", KERNELS)
def test_synthetic_batch_vs_oracle(kernel, oracle):
    """20k frames from the product's own generator (device) checked frame by frame."""
    torch = _torch()
    loc = make_loc(kernel)
    F = 20000
    adc, heads, cells = loc.synth_device(F, flags=2 | 4)      # random heads + KAT frames
    res = loc.localize_device(adc, heads, want=("lags", "cell", "highest"))
    torch.cuda.synchronize()
    o = oracle.localize(adc.cpu().numpy(), heads=heads.cpu().numpy(), want_corr=False, nthreads=16)
    assert (res["lags"].cpu().numpy() == o["lags"]).all()
    assert (res["cell"].cpu().numpy() == o["cell"]).all()
    assert (res["highest"].cpu().numpy() == o["highest"]).all()
    # KAT frame 0 is silence: every pair reports the first lag, -46
    assert res["lags"][0].tolist() == [-46, -46, -46]


# ---------------------------------------------------------------- the harness' generator
def test_synth_host_and_device_identical(loc):
    for flags in (0, 1, 2, 4, 7):
        a_h, h_h, c_h = loc.synth_host(48, flags=flags, first_frame=5 if flags == 0 else 0)
        a_d, h_d, c_d = loc.synth_device(48, flags=flags, first_frame=5 if flags == 0 else 0)
        assert (a_d.cpu().numpy() == a_h).all(), flags
        assert (h_d.cpu().numpy() == h_h).all() and (c_d.cpu().numpy() == c_h).all()
    assert len(np.unique(a_h)) > 100


def test_synth_sources_are_localized(loc):
    """Sanity of the harness, not parity: with integer delays the TDOA of most frames is the
    generator's delay difference, and the likelihood arg-max lands near the source cell."""
    torch = _torch()
    adc, heads, cells = loc.synth_device(4000, flags=1, first_frame=100)
    res = loc.localize_device(adc, None, want=("lags", "cell"))
    torch.cuda.synchronize()
    lags = res["lags"].cpu().numpy()
    lut = loc.lut().astype(int) - L
    expect = lut[:, cells.cpu().numpy()].T            # LUT lag of the true cell
    close = (np.abs(lags - expect) <= 1).all(1).mean()
    assert close > 0.9, close


# ---------------------------------------------------------------- drop-in symbols
def test_dropin_functions_match_reference(golden):
    import audio_triangulation_b200 as at
    from audio_triangulation_b200.api import BUFFER_DT, CORR_DT as CDT, RING_DT
    d = at.dropin
    assert (d.microphones_init().reshape(-1) == golden["mics"]).all()
    n_kat = len(golden["kat_names"])
    for k in (1, 2, 3, n_kat, n_kat + 1, n_kat + 2, n_kat + 9):
        bufs = []
        for m in range(3):
            ring = np.zeros(1, RING_DT)
            head = int(golden["heads"][k])
            ring["head"] = head
            ring["is_full"] = 1
            ring["buffer"][0] = np.roll(golden["adc"][k, m].astype(np.int16), head)
            b = np.zeros(1, BUFFER_DT)
            d.rolling_buffer_write_out(ring, b)
            assert (b["buffer"][0] == golden["after_dc"][k, m]).all() and b["power"][0] == golden["power"][k, m]
            d.buffer_normalize_range(b)
            assert (b["buffer"][0] == golden["after_shift"][k, m]).all()
            d.buffer_window(b)
            assert (b["buffer"][0] == golden["after_window"][k, m]).all()
            bufs.append(b)
        d.set_time_us(555)
        for p, (i, j) in enumerate(((0, 1), (0, 2), (1, 2))):
            c = np.zeros(1, CDT)
            d.correlations_init(c, bufs[i], bufs[j])
            assert (c["correlations"][0] == golden["corr"][k, p]).all()
            assert c["best_shift"][0] == golden["best_shift"][k, p] and c["last_update"][0] == 555
    # arbitrary int16 content (not reachable from 8-bit ADC data), incl. -32768 * -32768 products
    rng = np.random.default_rng(8)
    a = np.zeros(1, BUFFER_DT); b = np.zeros(1, BUFFER_DT)
    a["buffer"][0] = rng.integers(-32768, 32768, N); b["buffer"][0] = rng.integers(-32768, 32768, N)
    a["buffer"][0][:64] = -32768; b["buffer"][0][:80] = -32768
    c = np.zeros(1, CDT)
    d.correlations_init(c, a, b)
    from oracle_bindings import load_oracle
    lib = load_oracle()
    ref_c = np.zeros(NL, np.int64); best = np.zeros(1, np.int32)
    lib.ato_xcorr(np.ascontiguousarray(a["buffer"][0]), np.ascontiguousarray(b["buffer"][0]), N, L, ref_c, best)
    lib.ato_gauss(ref_c, L, int(best[0]))
    assert c["best_shift"][0] == best[0] and (c["correlations"][0] == ref_c).all()
    d.set_time_us(2**64 - 1)


def test_dropin_average_chain(golden):
    import audio_triangulation_b200 as at
    from audio_triangulation_b200.api import CORR_DT as CDT
    d = at.dropin
    n_kat = len(golden["kat_names"])
    est = [np.zeros(1, CDT) for _ in range(3)]
    for j, k in enumerate(range(n_kat, golden["adc"].shape[0])):
        d.set_time_us(int(golden["avg_times"][j]))
        for p in range(3):
            fresh = np.zeros(1, CDT)
            fresh["correlations"][0] = golden["corr"][k, p]
            d.correlations_average(est[p], fresh)
    d.set_time_us(2**64 - 1)
    for p in range(3):
        assert (est[p]["correlations"][0] == golden["avg_est"][p]).all()
        assert est[p]["best_shift"][0] == golden["avg_best"][p] and est[p]["last_update"][0] == golden["avg_last"][p]


def test_oracle_synth_equals_product_synth(loc, oracle):
    """bench.py's reference arm generates its inputs with oracle/synth_host.cpp (it must not map the product library);
    they are the product generator's frames byte for byte, on the device and on the host."""
    torch = _torch()
    F = 3000
    for flags in (0, 2 | 4, 1):
        adc, heads, cell = loc.synth_device(F, flags=flags, first_frame=12345)
        torch.cuda.synchronize()
        o_adc, o_heads, o_cell = oracle.synth(F, flags=flags, first_frame=12345)
        assert (adc.cpu().numpy() == o_adc).all() and (heads.cpu().numpy() == o_heads).all() and (cell.cpu().numpy() == o_cell).all()


# ---------------------------------------------------------------- a19 / a20 against the reference's own vga_heatmap.h
def test_lut_and_classes_equal_reference_heatmap_code(loc, golden, golden_hm):
    """tests/golden/ref_heatmap.npz was produced by vga_init_heatmap / vga_draw_heatmap compiled unmodified
    (oracle/hm_host.c).  The device-built geometry and look-up table, the stand-alone likelihood map and the fused
    kernel's `classes` output must reproduce the reference's table and its colour of every cell."""
    torch = _torch()
    assert (loc.mics().view(np.uint32) == golden_hm["mics"].view(np.uint32)).all()
    assert (loc.lut().reshape(3, -1) == golden_hm["lut"]).all()
    curves = golden_hm["curves"]
    hm = loc.heatmap_device(torch.from_numpy(curves).cuda(), want=("cell", "highest", "classes"))
    torch.cuda.synchronize()
    assert (hm["classes"].cpu().numpy() == golden_hm["classes"]).all()
    lut = golden_hm["lut"]
    like = curves[:, 0][:, lut[0]] + curves[:, 1][:, lut[1]] + curves[:, 2][:, lut[2]]          # [K][cells]
    assert (hm["highest"].cpu().numpy() == like.max(axis=1)).all() and (hm["cell"].cpu().numpy() == like.argmax(axis=1)).all()
    # the fused kernel on the golden frames: its per-frame curves are golden["corr"] = golden_hm["curves"][:K]
    K = golden["adc"].shape[0]
    assert (curves[:K] == golden["corr"]).all()
    got = loc.localize_device(torch.from_numpy(golden["adc"]).cuda(), want=("classes", "cell", "highest"))
    torch.cuda.synchronize()
    assert (got["classes"].cpu().numpy() == golden_hm["classes"][:K]).all()
    assert (got["cell"].cpu().numpy() == like[:K].argmax(axis=1)).all() and (got["highest"].cpu().numpy() == like[:K].max(axis=1)).all()


# ---------------------------------------------------------------- temporal stage + map on device arrays
def test_average_device_many_dt(loc, oracle, golden):
    """120 000 random time differences (microseconds to minutes): the batched EMA equals the reference formula
    (correlations.c:42-49, host libm) bit for bit for every one of them, gated arrays untouched."""
    torch = _torch()
    n_kat = len(golden["kat_names"])
    A = 40_000
    rng = np.random.default_rng(77)
    pick = rng.integers(n_kat, golden["adc"].shape[0], (2, A))
    est = golden["corr"][pick[0]].copy(); fresh = golden["corr"][pick[1]].copy()
    now = 200_000_000
    scale = 10.0 ** rng.uniform(0, 8.3, (A, 3))                   # dt from 1 us to 200 s
    times = (now - np.minimum(scale, now).astype(np.uint64)).astype(np.uint64)
    times[:100] = now                                              # dt = 0
    times[100:200, :] = times[100:200, :1]                         # arrays whose pairs share their stamp
    gate = (rng.random(A) > 0.1).astype(np.uint8)
    d_est = torch.from_numpy(est).cuda(); d_best = torch.zeros((A, 3), dtype=torch.int32, device="cuda")
    d_time = torch.from_numpy(times.view(np.int64)).cuda()
    loc.average_device(d_est, d_best, d_time, torch.from_numpy(fresh).cuda(), torch.from_numpy(gate).cuda(), now)
    torch.cuda.synchronize()
    exp = est.copy(); exp_best = np.zeros((A, 3), np.int32)
    lib = oracle.lib
    for a in np.nonzero(gate)[0]:
        for p in range(3):
            t = C.c_uint64(int(times[a, p]))
            lib.ato_average(exp[a, p], exp_best[a, p:p + 1], C.byref(t), fresh[a, p], L, now)
    got = d_est.cpu().numpy()
    bad = np.nonzero((got != exp).any(axis=2))
    assert bad[0].size == 0, f"{bad[0].size} (array, pair) entries differ, first: array {bad[0][0]} pair {bad[1][0]}"
    assert (d_best.cpu().numpy()[gate == 1] == exp_best[gate == 1]).all()


def test_average_and_heatmap_device(loc, oracle, golden):
    torch = _torch()
    n_kat = len(golden["kat_names"])
    A = 64
    rng = np.random.default_rng(12)
    pick = rng.integers(n_kat, golden["adc"].shape[0], (2, A))
    est = golden["corr"][pick[0]].copy(); fresh = golden["corr"][pick[1]].copy()
    times = rng.integers(0, 2_000_000, (A, 3)).astype(np.uint64)
    gate = (rng.random(A) > 0.3).astype(np.uint8)
    now = 2_500_000
    d_est = torch.from_numpy(est).cuda(); d_best = torch.zeros((A, 3), dtype=torch.int32, device="cuda")
    d_time = torch.from_numpy(times.view(np.int64)).cuda()
    loc.average_device(d_est, d_best, d_time, torch.from_numpy(fresh).cuda(), torch.from_numpy(gate).cuda(), now)
    torch.cuda.synchronize()
    exp = est.copy(); exp_best = np.zeros((A, 3), np.int32); exp_time = times.copy()
    for a in range(A):
        if not gate[a]:
            continue
        for p in range(3):
            t = C.c_uint64(int(times[a, p]))
            oracle.lib.ato_average(exp[a, p], exp_best[a, p:p + 1], C.byref(t), np.ascontiguousarray(fresh[a, p]), L, now)
            exp_time[a, p] = t.value
    got = d_est.cpu().numpy()
    assert (got == exp).all()            # decay factors come from the host libm: bit-equal by construction
    assert (d_best.cpu().numpy()[gate == 1] == exp_best[gate == 1]).all()
    assert (d_time.cpu().numpy().view(np.uint64) == exp_time).all()
    hm = loc.heatmap_device(torch.from_numpy(exp).cuda(), want=("cell", "highest", "xy", "classes"))
    torch.cuda.synchronize()
    for a in range(A):
        hi = np.zeros(1, np.int64); cell = np.zeros(1, np.int32); cls = np.zeros(CELLS, np.uint8)
        oracle.lib.ato_heatmap(exp[a].reshape(-1), oracle.lut.reshape(-1), 3, CELLS, L, hi.ctypes.data, cell.ctypes.data,
                               cls.ctypes.data)
        assert hm["cell"][a].item() == cell[0] and hm["highest"][a].item() == hi[0]
        assert (hm["classes"][a].cpu().numpy() == cls).all()


# ---------------------------------------------------------------- host API, sharding, edge cases
def test_host_api_and_sharding_match_device_api(loc):
    import audio_triangulation_b200 as at
    from audio_triangulation_b200 import _lib
    torch = _torch()
    F = 5003                                # ragged: not a multiple of the grid or the chunk
    adc, heads, _ = loc.synth_device(F, flags=2)
    dev = loc.localize_device(adc, heads, want=("lags", "cell", "corr"))
    torch.cuda.synchronize()
    adc_h, heads_h = adc.cpu().numpy(), heads.cpu().numpy()
    import os
    os.environ["AT_CHUNK_FRAMES"] = "1000"  # force several chunks through both stream slots
    try:
        host = loc.localize_host(adc_h, heads_h, want=("lags", "cell", "corr"))
        pinned = torch.from_numpy(adc_h).pin_memory()
        host2 = loc.localize_host(pinned, heads_h, want=("lags",))
        # two contexts on the same GPU stand in for two GPUs
        loc2 = at.Localizer(device=0)
        lags = np.zeros((F, 3), np.int32); cell = np.zeros(F, np.int32)
        o = _lib.AtOutputs(); o.lags = lags.ctypes.data; o.cell = cell.ctypes.data
        ctxs = (C.c_void_p * 2)(loc.ctx, loc2.ctx)
        _lib.check(loc.lib.at_localize_host_sharded(ctxs, 2, adc_h.ctypes.data, heads_h.ctypes.data, F, C.byref(o)))
    finally:
        del os.environ["AT_CHUNK_FRAMES"]
    for k in ("lags", "cell", "corr"):
        assert (host[k] == dev[k].cpu().numpy()).all(), k
    assert (host2["lags"] == host["lags"]).all()
    assert (lags == host["lags"]).all() and (cell == host["cell"]).all()
    # whole correlations_t structs, pinned in and out, 13 chunks rotating over the three slots (copy-in, kernel and copy-out
    # streams chained by events), twice in a row on the same slots
    os.environ["AT_CHUNK_FRAMES"] = "400"
    try:
        devs = loc.localize_device(adc, heads, want=("lags", "corr", "xy"), struct_corr=True)
        torch.cuda.synchronize()
        outp = {"lags": torch.empty((F, 3), dtype=torch.int32).pin_memory(),
                "corr": torch.empty((F, 3, 95), dtype=torch.int64).pin_memory(),
                "xy": torch.empty((F, 2), dtype=torch.float32).pin_memory()}
        for _ in range(2):
            outp["corr"].zero_()
            loc.localize_host(pinned, heads_h, want=("lags", "corr", "xy"), out=outp, struct_corr=True)
            dcorr, hcorr = devs["corr"].cpu().numpy(), outp["corr"].numpy()
            assert (hcorr[:, :, :94] == dcorr[:, :, :94]).all()      # curves and best_shift; the time stamp differs per call
            assert (outp["lags"].numpy() == devs["lags"].cpu().numpy()).all()
            assert (outp["xy"].numpy() == devs["xy"].cpu().numpy()).all()
    finally:
        del os.environ["AT_CHUNK_FRAMES"]


def test_host_api_config4_tcgen05_chunks():
    """8 mics x 4096 through at_localize_host: several chunks on both stream slots, each a launch of the tcgen05 kernel
    (512 TMEM columns allocated and released per CTA), equal to the device API."""
    import os
    torch = _torch()
    loc = make_loc("auto", n_mics=8, n_bits=12, max_shift=46)
    F = 701
    adc, heads, _ = loc.synth_device(F, flags=2, seed=3)
    dev = loc.localize_device(adc, heads, want=("lags", "raw"))
    torch.cuda.synchronize()
    os.environ["AT_CHUNK_FRAMES"] = "150"
    try:
        host = loc.localize_host(adc.cpu().numpy(), heads.cpu().numpy(), want=("lags", "raw"))
    finally:
        del os.environ["AT_CHUNK_FRAMES"]
    for k in ("lags", "raw"):
        assert (host[k] == dev[k].cpu().numpy()).all(), k
    loc.close()


def test_edge_cases(loc):
    import audio_triangulation_b200 as at
    torch = _torch()
    empty = torch.empty((0, 3, N), dtype=torch.uint8, device="cuda")
    assert loc.localize_device(empty, want=("lags",))["lags"].shape == (0, 3)
    one, _, _ = loc.synth_device(1)
    r = loc.localize_device(one, want=("lags",))
    torch.cuda.synchronize()
    assert r["lags"].shape == (1, 3)
    with pytest.raises(at.AtError):
        at.Localizer(device=0, n_mics=9)
    with pytest.raises(at.AtError):
        at.Localizer(device=0, max_shift=200)
    with pytest.raises(at.AtError):                       # in range, but no fused kernel is instantiated for it: refused at create
        at.Localizer(device=0, n_mics=5)
    before = loc.kernel_launches()
    loc.localize_device(one, want=("lags",))
    # one fused launch per call -- two for the tcgen05 kernel of the reference shape (certified pass + exact pass over its list)
    assert loc.kernel_launches() - before in (1, 2)


@pytest.mark.parametrize("kernel", KERNELS)
def test_ragged_small_batches(kernel, oracle):
    """Batch sizes around the warp / CTA granularity of the kernels (1..9, 127..130 frames), every frame checked."""
    loc = make_loc(kernel)
    adc_all, heads_all, _ = loc.synth_device(130, flags=2 | 4, seed=5)
    torch = _torch()
    o = oracle.localize(adc_all.cpu().numpy(), heads=heads_all.cpu().numpy(), want_corr=False)
    for F in (1, 2, 3, 4, 5, 7, 8, 9, 127, 128, 129, 130):
        r = loc.localize_device(adc_all[:F].contiguous(), heads_all[:F].contiguous(), want=("lags", "cell", "highest"))
        torch.cuda.synchronize()
        assert (r["lags"].cpu().numpy() == o["lags"][:F]).all(), F
        assert (r["cell"].cpu().numpy() == o["cell"][:F]).all() and (r["highest"].cpu().numpy() == o["highest"][:F]).all(), F


@pytest.mark.parametrize("kernel", ("auto", "imma"))
def test_bounded_search_paths_are_exercised(kernel, oracle):
    """The likelihood search must take its routes on suitable data and still equal the oracle's full scan: clean bursts,
    heavy-noise frames, and flat frames.  Routes: peak-tuple look-up / full scan over all LUT tuples for the tcgen05 kernel
    (AUTO), plus the first and the widened box of the warp-scope bounded search for the mma.sync kernel."""
    torch = _torch()
    loc = make_loc(kernel)
    rng = np.random.default_rng(17)
    clean, _ = burst_frames(256, seed=3)
    noisy = rng.integers(0, 256, (256, 3, N), dtype=np.uint8)                      # white noise: inconsistent peaks
    weak, _ = burst_frames(256, seed=4)
    weak = np.clip(128 + (weak.astype(int) - 128) // 6 + rng.integers(-6, 7, weak.shape), 0, 255).astype(np.uint8)
    flat = np.full((64, 3, N), 77, np.uint8)
    adc = np.concatenate([clean, noisy, weak, flat])
    d = torch.from_numpy(adc).cuda()
    r = loc.localize_device(d, want=("lags", "cell", "highest", "stats"))
    torch.cuda.synchronize()
    o = oracle.localize(adc, want_corr=False, nthreads=8)
    assert (r["cell"].cpu().numpy() == o["cell"]).all() and (r["highest"].cpu().numpy() == o["highest"]).all()
    st = r["stats"].cpu().numpy()
    assert st[:4].sum() == adc.shape[0] and st[3] > 0 and st[2] > 0, st          # look-up and full scan used
    if kernel == "imma":
        assert st[0] + st[1] > 0, st                                              # and a bounded box
    # without `highest` the tensor kernel may certify the arg-max from nine of the twelve digit products (no l.l):
    # same lags and cells, and the clean bursts must take that route while white noise and flat frames must not
    r2 = loc.localize_device(d, want=("lags", "cell", "stats"))
    torch.cuda.synchronize()
    assert (r2["cell"].cpu().numpy() == o["cell"]).all() and (r2["lags"].cpu().numpy() == o["lags"]).all()
    st2 = r2["stats"].cpu().numpy()
    assert st2[:4].sum() == adc.shape[0] and 0 < st2[4] <= st2[3] and st2[4] < adc.shape[0], st2


# ---------------------------------------------------------------- 3-D candidate sets (no reference pin beyond the table's arithmetic)
def test_points_lut_reference_candidates_equal_reference_table(golden_hm):
    """AT_LUT_POINTS with the reference's own candidate positions gives the reference's table (and the same cells)."""
    import audio_triangulation_b200 as at
    torch = _torch()
    f = np.float32
    y, x = np.divmod(np.arange(CELLS), 101)
    xm = (x - 50).astype(f) / f(24.0); ym = (50 - y).astype(f) / f(24.0); zm = np.full(CELLS, 1.2, f)
    k = f(1.2) / np.sqrt(zm * zm + xm * xm + ym * ym, dtype=f)
    pts = np.stack([xm * k, ym * k, zm * k], 1).astype(f)
    loc = at.Localizer(device=0, points=pts)
    assert (loc.lut() == golden_hm["lut"]).all()
    ref = at.Localizer(device=0)
    adc, heads, _ = ref.synth_device(600, flags=2 | 4)
    a = loc.localize_device(adc, heads, want=("lags", "cell", "highest")); b = ref.localize_device(adc, heads, want=("lags", "cell", "highest"))
    torch.cuda.synchronize()
    for k_ in ("lags", "cell", "highest"):
        assert torch.equal(a[k_], b[k_]), k_


def test_hemisphere_lut_8_mics_position(oracle):
    """Config 4 gets a position: 8 microphones x 4096 samples (28 pairs), candidate directions on a hemisphere (azimuth x
    elevation at 2 m), likelihood arg-max over the 28 curves inside the tcgen05 kernel's epilogue.  Table, cell and
    highest_L equal the oracle's restatement; the synthetic sources are found."""
    import audio_triangulation_b200 as at
    torch = _torch()
    n_az, n_el, R = 72, 12, 2.0
    pts = at.hemisphere_points(n_az, n_el, R)
    M, nb, Ls = 8, 12, 46
    loc = at.Localizer(device=0, n_mics=M, n_bits=nb, max_shift=Ls, points=pts)
    mics = loc.mics()
    P, n = M * (M - 1) // 2, pts.shape[0]
    idx = np.zeros((P, n), np.uint8)
    oracle.lib.ato_lut_build_points(mics.reshape(-1), M, Ls, 50000.0, 343.0, pts.reshape(-1), n, idx.reshape(-1))
    assert (loc.lut() == idx).all()
    F = 96
    adc, heads, truth = loc.synth_device(F, flags=2, seed=11)
    r = loc.localize_device(adc, heads, want=("lags", "cell", "highest", "xy"))
    torch.cuda.synchronize()
    o = Oracle(n_mics=M, n_bits=nb, max_shift=Ls, lut=idx, n_cells=n).localize(adc.cpu().numpy(), heads=heads.cpu().numpy(),
                                                                               want_corr=False, nthreads=16)
    assert (r["lags"].cpu().numpy() == o["lags"]).all()
    assert (r["cell"].cpu().numpy() == o["cell"]).all() and (r["highest"].cpu().numpy() == o["highest"]).all()
    cell = r["cell"].cpu().numpy(); t = truth.cpu().numpy()
    assert (r["xy"].cpu().numpy().view(np.uint32) == pts[cell][:, :2].view(np.uint32)).all()
    # direction error: angle between the estimated and the true unit vectors (a planar array resolves azimuth well and
    # elevation coarsely; most estimates must fall within 15 degrees)
    cosang = (pts[cell] * pts[t]).sum(1) / (R * R)
    assert (np.degrees(np.arccos(np.clip(cosang, -1, 1))) < 15.0).mean() > 0.7


# ---------------------------------------------------------------- other shapes (no reference pin)
@pytest.mark.parametrize("shape", [(8, 12, 46, 1500), (8, 10, 46, 6000)])
def test_umma_8mic_large_batch_equals_mma_sync_kernel(shape):
    """Many frames per CTA (TMEM slot and plane-buffer re-use, the last pass' shared atomics): the tcgen05 kernel
    against the mma.sync CTA kernel, itself oracle-checked below.  No reference pin for 8 microphones."""
    M, nb, Ls, F = shape
    torch = _torch()
    res = {}
    for kernel in ("imma", "umma"):
        loc = make_loc(kernel, n_mics=M, n_bits=nb, max_shift=Ls)
        adc, heads, _ = loc.synth_device(F, flags=2, seed=5)
        r = loc.localize_device(adc, heads, want=("lags", "raw", "corr"))
        torch.cuda.synchronize()
        res[kernel] = {k: v.cpu() for k, v in r.items()}
        loc.close()
    for k in ("lags", "raw", "corr"):
        assert torch.equal(res["imma"][k], res["umma"][k]), k


@pytest.mark.parametrize("kernel", KERNELS + ("auto",))
@pytest.mark.parametrize("shape", [(8, 12, 46, 24), (3, 10, 44, 64), (4, 10, 46, 64), (3, 12, 46, 32), (8, 10, 46, 48)])
def test_generalised_shapes_vs_oracle(kernel, shape):
    M, nb, Ls, F = shape
    loc = make_loc(kernel, n_mics=M, n_bits=nb, max_shift=Ls, sample_rate_hz=48000.0 if Ls == 44 else 50000.0)
    torch = _torch()
    adc, heads, _ = loc.synth_device(F, flags=2, seed=77)
    try:
        res = loc.localize_device(adc, heads, want=("lags", "raw", "corr", "cell", "highest", "windowed"))
    except Exception as e:
        if "instantiation" in str(e):
            pytest.skip(str(e))
        raise
    torch.cuda.synchronize()
    lib = Oracle().lib
    P = M * (M - 1) // 2
    lut = np.zeros((P, CELLS), np.uint8)
    lib.ato_lut_build(loc.mics().reshape(-1), M, Ls, 48000.0 if Ls == 44 else RATE_HZ, SPEED, HALF_W, HALF_H, PX_PER_M,
                      HEIGHT, lut.reshape(-1))
    assert (loc.lut() == lut).all()
    o = Oracle(n_mics=M, n_bits=nb, max_shift=Ls, lut=lut).localize(adc.cpu().numpy(), heads=heads.cpu().numpy(),
                                                                    want_raw=True, nthreads=8)
    for k in ("lags", "raw", "corr", "cell", "highest"):
        assert (res[k].cpu().numpy() == o[k]).all(), k


# ---------------------------------------------------------------- BASELINE config 2 at full size
def test_full_batch_properties(loc, oracle):
    """2^20 frames (3.2 GB of ADC bytes) on one GPU: spot-check against the oracle and verify
    size-independent properties (batch-split invariance, ring-rotation invariance)."""
    torch = _torch()
    F = 1 << 20
    adc, heads, cells = loc.synth_device(F, flags=4)
    lags = loc.localize_device(adc, None, want=("lags", "cell"))
    torch.cuda.synchronize()
    idx = np.concatenate([np.arange(8), np.random.default_rng(0).integers(0, F, 3000), [F - 1]])
    sel = torch.from_numpy(idx).cuda()
    o = oracle.localize(adc[sel].cpu().numpy(), want_corr=False, nthreads=16)
    assert (lags["lags"][sel].cpu().numpy() == o["lags"]).all()
    assert (lags["cell"][sel].cpu().numpy() == o["cell"]).all()
    # split invariance: second half alone == second half of the whole
    half = loc.localize_device(adc[F // 2:], None, want=("lags",))
    torch.cuda.synchronize()
    assert torch.equal(half["lags"], lags["lags"][F // 2:])
    # rotation invariance: store a slice rotated by per-frame heads
    n = 4096
    h = torch.randint(0, N, (n,), device="cuda", dtype=torch.int32)
    ar = torch.arange(N, device="cuda").view(1, 1, N)
    src = (ar - h.view(n, 1, 1)) % N
    rolled = torch.gather(adc[:n], 2, src.expand(n, 3, N).long()).contiguous()
    rot = loc.localize_device(rolled, h, want=("lags",))
    torch.cuda.synchronize()
    assert torch.equal(rot["lags"], lags["lags"][:n])


def test_config4_batch_properties():
    """BASELINE config 4 (8 mics x 4096 samples, 28 pairs; no reference pin) at a batch of 2^13 frames through the
    kernel AUTO selects (tcgen05): oracle spot check, batch-split invariance and ring-rotation invariance."""
    torch = _torch()
    M, nb, Ls, F = 8, 12, 46, 1 << 13
    n_s = 1 << nb
    loc = make_loc("auto", n_mics=M, n_bits=nb, max_shift=Ls)
    adc, _, _ = loc.synth_device(F, seed=11)
    res = loc.localize_device(adc, None, want=("lags", "raw"))
    torch.cuda.synchronize()
    idx = np.concatenate([np.arange(4), np.random.default_rng(1).integers(0, F, 40), [F - 1]])
    sel = torch.from_numpy(idx).cuda()
    o = Oracle(n_mics=M, n_bits=nb, max_shift=Ls, lut=None).localize(adc[sel].cpu().numpy(), want_raw=True, want_cell=False, nthreads=16)
    assert (res["lags"][sel].cpu().numpy() == o["lags"]).all()
    assert (res["raw"][sel].cpu().numpy() == o["raw"]).all()
    half = loc.localize_device(adc[F // 2:], None, want=("lags",))
    torch.cuda.synchronize()
    assert torch.equal(half["lags"], res["lags"][F // 2:])
    n = 512
    h = torch.randint(0, n_s, (n,), device="cuda", dtype=torch.int32)
    ar = torch.arange(n_s, device="cuda").view(1, 1, n_s)
    src = (ar - h.view(n, 1, 1)) % n_s
    rolled = torch.gather(adc[:n], 2, src.expand(n, M, n_s).long()).contiguous()
    rot = loc.localize_device(rolled, h, want=("lags",))
    torch.cuda.synchronize()
    assert torch.equal(rot["lags"], res["lags"][:n])
    loc.close()


def test_certified_argmax_across_signal_levels(loc):
    """The default tensor kernel certifies the arg-max from nine of the twelve digit products when its bound allows and
    falls back to the exact path otherwise.  Sweep the signal level from barely above the noise to full scale so that
    frames sit on both sides of the bound, and demand lags and cells identical to the integer-pipe kernel (which has
    no shortcut and is itself oracle-checked) on every frame."""
    torch = _torch()
    F = 1 << 16
    adc, _, _ = loc.synth_device(F, seed=99)
    scale = (torch.arange(F, device="cuda") % 33).view(F, 1, 1).float() / 32.0          # 0 .. 1 in 33 steps
    dimmed = (128.0 + (adc.float() - 128.0) * scale).round().clamp(0, 255).to(torch.uint8).contiguous()
    got = loc.localize_device(dimmed, want=("lags", "cell", "stats"))
    ref_loc = make_loc("imad")
    exp = ref_loc.localize_device(dimmed, want=("lags", "cell"))
    torch.cuda.synchronize()
    assert torch.equal(got["lags"], exp["lags"]) and torch.equal(got["cell"], exp["cell"])
    st = got["stats"].cpu().numpy()
    assert 0 < st[4] < F, st          # both routes were taken
    ref_loc.close()


def test_admissible_lag_windows(loc):
    """Per-pair admissible lag windows (an extension -- the reference scans +-46 for every pair):
    window = ceil(distance * fs / c), and the windowed first-max arg-max equals a numpy restatement on the raw curves."""
    torch = _torch()
    lim = loc.pair_max_shift()
    mics = loc.mics().astype(np.float32)
    exp_lim = []
    for i in range(3):
        for j in range(i + 1, 3):
            d = np.float32(np.sqrt(np.float32((mics[i, 0] - mics[j, 0]) ** 2 + (mics[i, 1] - mics[j, 1]) ** 2)))
            exp_lim.append(min(L, int(np.ceil(np.float32(d * np.float32(RATE_HZ)) / np.float32(SPEED)))))
    assert lim.tolist() == exp_lim == [20, 30, 22]          # sides 0.132 / 0.20 / 0.15 m at 50 kHz, 343 m/s
    names, kats = kat_frames()
    bursts, _ = burst_frames(200, seed=12, max_delay=45)    # delays beyond the windows: the two arg-maxes must differ somewhere
    adc = np.concatenate([kats, bursts])
    r = loc.localize_device(torch.from_numpy(adc).cuda(), want=("lags", "raw"))
    got = loc.admissible_lags_device(r["raw"]).cpu().numpy()
    raw = r["raw"].cpu().numpy()
    exp = np.stack([np.argmax(raw[:, p, L - lim[p]:L + lim[p] + 1], axis=1) - lim[p] for p in range(3)], axis=1)   # argmax = first max
    assert (got == exp).all()
    full = r["lags"].cpu().numpy()
    inside = np.abs(full) <= lim[None, :]
    assert (got[inside] == full[inside]).all() and (~inside).any()
