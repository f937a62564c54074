import os, sys, torch
sys.path.insert(0, os.getcwd())
import audio_triangulation_b200 as at
for nb, F in ((12, 1 << 14), (10, 1 << 16)):
    loc = at.Localizer(n_mics=8, n_bits=nb)
    for flags in (0, 2):
        adc, heads, _ = loc.synth_device(F, flags=flags)
        out = {}
        for _ in range(3): loc.localize_device(adc, heads if flags else None, want=("lags",), out=out)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10): loc.localize_device(adc, heads if flags else None, want=("lags",), out=out)
        b.record(); torch.cuda.synchronize()
        print("8 x %d, %s heads: %.2f M frames/s" % (1 << nb, "random" if flags else "aligned", F / (a.elapsed_time(b) / 10) / 1e3))
