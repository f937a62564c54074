"""Direct-vs-FFT crossover (BASELINE config 4): 8 mics x 4096 samples, 28 pairs.
Measured: direct integer correlation on tensor cores (auto = tcgen05 where it exists, imma = mma.sync; +-46 lags), on
the integer pipe (imad), and the hand-written FFT/GCC-PHAT variant (cost independent of the lag range).  Model: the
direct tensor form computes 128-lag tiles (93 used), so its time scales with ceil((2L+1+padding)/128); the crossover
lag range follows."""
import json, math, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import audio_triangulation_b200 as at

def timeit(fn, reps=3):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps

out = {}
for M, nb, F in ((8, 12, 8192), (3, 10, 1 << 17)):
    N = 1 << nb
    row = {}
    for kernel in ("auto", "imma", "imad"):
        loc = at.Localizer(kernel=kernel, n_mics=M, n_bits=nb)
        adc, _, _ = loc.synth_device(F)
        o = {}
        ms = timeit(lambda: loc.localize_device(adc, want=("lags",), out=o))
        row[kernel + "_frames_per_s"] = F / ms * 1e3
        loc.close()
    loc = at.Localizer(n_mics=M, n_bits=nb)
    adc, _, _ = loc.synth_device(F)
    ms = timeit(lambda: loc.gccphat_device(adc))
    row["gccphat_frames_per_s"] = F / ms * 1e3
    ratio = max(row["auto_frames_per_s"], row["imma_frames_per_s"]) / row["gccphat_frames_per_s"]
    # direct tensor cost ~ tiles(L) = ceil((2L + 1 + 3) / 128) (PAD alignment), FFT cost constant
    tiles = math.floor(ratio)
    row["direct_over_fft_at_L46"] = ratio
    row["crossover_lag_range_model"] = "FFT wins once the direct form needs more than %d 128-lag tiles, i.e. L > ~%d samples (%.2f m aperture at 50 kHz)" % (
        tiles, (128 * tiles - 4) // 2, ((128 * tiles - 4) // 2) / 50000.0 * 343.0)
    out["%d mics x %d" % (M, N)] = row
    print(M, N, row, flush=True)
json.dump(out, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "crossover.json"), "w"), indent=1)
