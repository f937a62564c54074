"""A/B of library builds on config 4 (8 x 4096) and the white-noise / whole-struct modes of the reference shape, one call."""
import os, subprocess, sys
code = r'''
import os, sys, torch
sys.path.insert(0, os.getcwd())
import audio_triangulation_b200 as at
def t(loc, adc, want, n=6, **kw):
    out = {}
    for _ in range(2): loc.localize_device(adc, want=want, out=out, **kw)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): loc.localize_device(adc, want=want, out=out, **kw)
    b.record(); torch.cuda.synchronize()
    return adc.shape[0] / (a.elapsed_time(b) / n) / 1e3
loc = at.Localizer(n_mics=8, n_bits=12, points=at.hemisphere_points(72, 12, 2.0))
adc, _, _ = loc.synth_device(1 << 14)
r = [t(loc, adc, ("lags",)), t(loc, adc, ("lags", "cell", "xy"))]
loc.close()
loc = at.Localizer()
adc, _, _ = loc.synth_device(1 << 18, flags=16)
r.append(t(loc, adc, ("lags", "cell", "xy")))
adc, _, _ = loc.synth_device(1 << 18)
r.append(t(loc, adc, ("lags", "corr"), struct_corr=True))
r.append(t(loc, adc, ("lags", "cell", "xy")))
print(" ".join("%.2f" % x for x in r))
'''
for v in sys.argv[1:]:
    o = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=dict(os.environ, AT_LIB_VARIANT=("" if v == "main" else v)), timeout=200)
    print("%-6s c4 lags / c4 lags+pos / white noise / struct / normal (M frames/s): %s" % (v, o.stdout.strip().splitlines()[-1] if o.stdout.strip() else o.stderr[-300:]))
