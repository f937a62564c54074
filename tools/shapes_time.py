"""Throughput of the other supported shapes (lags only), tensor vs integer-pipe kernel."""
import sys, os, json, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import audio_triangulation_b200 as at
for M, nb, L, F in ((3, 12, 46, 1 << 16), (3, 10, 44, 1 << 18)):
    for kernel in ("imad", "imma"):
        loc = at.Localizer(kernel=kernel, n_mics=M, n_bits=nb, max_shift=L, sample_rate_hz=48000.0 if L == 44 else 50000.0)
        adc, _, _ = loc.synth_device(F)
        out = {}
        for _ in range(2): loc.localize_device(adc, want=("lags",), out=out)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5): loc.localize_device(adc, want=("lags",), out=out)
        b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 5
        print(M, 1 << nb, L, kernel, "%.3f ms  %.3f Mframes/s" % (ms, F / ms / 1e3), flush=True)
        loc.close()
