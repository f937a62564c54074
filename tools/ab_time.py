"""A/B timing of library builds inside ONE gpurun call (boxes differ by several per cent): alternates the variants
(AT_LIB_VARIANT names, '' = the product) in subprocesses, several rounds, and prints the median of each.
usage: python tools/ab_time.py base '' [rounds]"""
import os, subprocess, sys, statistics
variants = [v for v in sys.argv[1:] if not v.isdigit()] or ["base", ""]
rounds = int([v for v in sys.argv[1:] if v.isdigit()][0]) if any(v.isdigit() for v in sys.argv[1:]) else 3
code = r'''
import os, sys, torch
sys.path.insert(0, os.getcwd())
import audio_triangulation_b200 as at
loc = at.Localizer(kernel="umma")
F = 1 << 19
adc, _, _ = loc.synth_device(F)
res = []
for want in (("lags",), ("lags", "cell", "xy")):
    out = {}
    for _ in range(3): loc.localize_device(adc, want=want, out=out)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20): loc.localize_device(adc, want=want, out=out)
    b.record(); torch.cuda.synchronize()
    res.append(F / (a.elapsed_time(b) / 20) / 1e3)
print("%.1f %.1f" % tuple(res))
'''
acc = {v: [] for v in variants}
for r in range(rounds):
    for v in variants:
        env = dict(os.environ, AT_LIB_VARIANT=v)
        o = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=120)
        try:
            acc[v].append(tuple(float(x) for x in o.stdout.split()[-2:]))
        except Exception:
            print("variant %r failed: %s" % (v, o.stderr[-300:]))
for v in variants:
    if acc[v]:
        print("variant %-8r lags %.1f  lags+cell+xy %.1f  M frames/s (median of %d)" % (v, statistics.median(x[0] for x in acc[v]), statistics.median(x[1] for x in acc[v]), len(acc[v])))
