"""Timing experiment: kernel time with phases skipped (AT_DEBUG_SKIP, results are then wrong)."""
import os, sys, subprocess, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import sys, torch
sys.path.insert(0, %r)
import audio_triangulation_b200 as at
loc = at.Localizer()
F = 1 << 19
adc, _, _ = loc.synth_device(F)
for want in (("lags",), ("lags", "cell", "xy")):
    out = {}
    for _ in range(3): loc.localize_device(adc, want=want, out=out)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5): loc.localize_device(adc, want=want, out=out)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 5
    print("   want=%%-22s %%.3f ms  %%.1f Mframes/s" %% ("+".join(want), ms, F / ms / 1e3))
''' % ROOT
for skip, name in ((0, "full"), (1, "no prep"), (2, "no MMA loop"), (4, "no epilogue"), (3, "no prep, no MMA"), (6, "no MMA, no epilogue"), (5, "MMA only")):
    env = dict(os.environ, AT_DEBUG_SKIP=str(skip))
    print(name, flush=True)
    subprocess.run([sys.executable, "-c", code], env=env)
