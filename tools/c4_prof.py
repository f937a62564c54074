"""Cycle account of the group epilogue in the 8-microphone kernel (needs make VARIANT=prof): AT_LIB_VARIANT=prof python tools/c4_prof.py"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import audio_triangulation_b200 as at
loc = at.Localizer(n_mics=8, n_bits=12, points=at.hemisphere_points(72, 12, 2.0))
adc, _, _ = loc.synth_device(1 << 13)
out = {}
for want in (("lags",), ("lags", "cell", "xy")):
    for _ in range(2): loc.localize_device(adc, want=want, out=out)
    torch.cuda.synchronize()
    os.environ["AT_PROF_PRINT"] = "1"
    print("want =", want, flush=True)
    loc.localize_device(adc, want=want, out=out)
    torch.cuda.synchronize()
    os.environ.pop("AT_PROF_PRINT")
