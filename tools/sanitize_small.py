"""Small all-paths run: every kernel variant, every output, ragged batch sizes, ring heads, the 8-microphone kernels,
streaming, the drop-in symbols -- each result compared with the oracle.  Run it against the debug build
(AT_LIB_VARIANT=checked, in-kernel assertions; tests/test_gpu_checked.py does) or under compute-sanitizer where that
tool is open."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import audio_triangulation_b200 as at
from oracle_bindings import Oracle   # the checker

ALL = ("lags", "corr", "raw", "cell", "highest", "xy", "gate", "classes", "windowed", "power")
orc = Oracle()
for kernel in ("auto", "imma", "umma", "imad"):
    loc = at.Localizer(kernel=kernel)
    adc, heads, _ = loc.synth_device(1203, flags=2 | 4)                 # ragged: several frames per CTA, a partial last round
    exp = orc.localize(adc.cpu().numpy(), heads=heads.cpu().numpy(), want_raw=True)
    for F in (1, 5, 67, 1203):
        r = loc.localize_device(adc[:F].contiguous(), heads[:F].contiguous(), want=ALL)
        r2 = loc.localize_device(adc[:F].contiguous(), heads[:F].contiguous(), want=("lags", "cell", "xy", "gate"))
        torch.cuda.synchronize()
        for k in ("lags", "corr", "raw", "cell", "highest"):
            assert (r[k].cpu().numpy() == exp[k][:F]).all(), (kernel, F, k)
        assert (r2["lags"].cpu().numpy() == exp["lags"][:F]).all() and (r2["cell"].cpu().numpy() == exp["cell"][:F]).all(), (kernel, F)
    # white noise and flat frames: the routes that certify nothing / scan everything
    noise = torch.randint(0, 256, (300, 3, 1024), dtype=torch.uint8, device="cuda")
    noise[:20] = 77
    e2 = orc.localize(noise.cpu().numpy(), want_corr=False)
    r3 = loc.localize_device(noise, want=("lags", "cell", "xy"))
    torch.cuda.synchronize()
    assert (r3["lags"].cpu().numpy() == e2["lags"]).all() and (r3["cell"].cpu().numpy() == e2["cell"]).all(), kernel
    print(kernel, "ok", r["lags"][:2].tolist(), int(r["cell"][0]))
    loc.close()
# 8-microphone kernels (tcgen05 for 4096 / 1024 samples, mma.sync CTA kernel), hemisphere candidates
for nb, kernel in ((12, "auto"), (10, "umma"), (10, "imma")):
    loc = at.Localizer(n_mics=8, n_bits=nb, kernel=kernel, points=at.hemisphere_points(24, 6, 2.0))
    adc, heads, _ = loc.synth_device(37, flags=2)
    r = loc.localize_device(adc, heads, want=("lags", "raw", "cell", "highest", "xy"))
    torch.cuda.synchronize()
    o = Oracle(n_mics=8, n_bits=nb, max_shift=46, lut=loc.lut(), n_cells=loc.n_cells).localize(
        adc.cpu().numpy(), heads=heads.cpu().numpy(), want_raw=True, want_corr=False, nthreads=8)
    assert (r["lags"].cpu().numpy() == o["lags"]).all() and (r["raw"].cpu().numpy() == o["raw"]).all(), (nb, kernel)
    assert (r["cell"].cpu().numpy() == o["cell"]).all(), (nb, kernel)
    print("8 mics", nb, kernel, "ok")
    loc.close()
# FFT / GCC-PHAT variant, both inverse forms, ragged sizes (partial column groups)
for M, nb, F in ((8, 12, 13), (3, 10, 131), (4, 10, 40)):
    loc = at.Localizer(n_mics=M, n_bits=nb)
    adc, heads, _ = loc.synth_device(F, flags=2)
    res = {}
    for inv in ("dft", "fft"):
        os.environ["AT_GCC_INVERSE"] = inv
        res[inv] = loc.gccphat_device(adc, heads).cpu().numpy()
    os.environ.pop("AT_GCC_INVERSE")
    assert (res["dft"] == res["fft"]).mean() > 0.98, (M, nb)
    print("gcc-phat", M, nb, "ok")
    loc.close()
loc = at.Localizer()
adc, heads, _ = loc.synth_device(67, flags=2 | 4)
st = at.Stream(loc, 5)
x = torch.randint(100, 156, (5, 1024, 3), dtype=torch.uint8, device="cuda")
for _ in range(3):
    st.push(x)
est = torch.zeros((5, 3, 93), dtype=torch.int64, device="cuda")
loc.heatmap_device(est, want=("cell", "highest", "xy", "classes"))
h = loc.localize_host(adc.cpu().numpy(), heads.cpu().numpy(), want=("lags", "cell"))
d = at.dropin
b = np.zeros(1, at.api.BUFFER_DT); d.buffer_window(b); d.buffer_normalize_range(b)
c = np.zeros(1, at.api.CORR_DT); d.correlations_init(c, b, b); d.correlations_average(c, c)
torch.cuda.synchronize()
print("checked run ok" if os.environ.get("AT_LIB_VARIANT") == "checked" else "run ok")
