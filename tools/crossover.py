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
    os.environ["AT_GCC_INVERSE"] = "fft"
    ms = timeit(lambda: loc.gccphat_device(adc))
    os.environ.pop("AT_GCC_INVERSE")
    row["gccphat_fft_inverse_frames_per_s"] = F / ms * 1e3
    direct = max(row["auto_frames_per_s"], row["imma_frames_per_s"])
    row["direct_over_gccphat_at_L46"] = direct / row["gccphat_frames_per_s"]
    row["direct_over_gccphat_fft_inverse_at_L46"] = direct / row["gccphat_fft_inverse_frames_per_s"]
    # The direct tensor form works on 128-row lag tiles (2L + 1 lags + 15 rows of phase skew per tile row block): its time
    # scales with ceil((2L + 16) / 128).  The FFT-inverse form does not depend on L (below N / 16); the contraction form
    # needs one more 128-row operand block per 128 lags on its inverse side only.
    tiles = math.floor(row["direct_over_gccphat_fft_inverse_at_L46"])
    row["crossover_model"] = ("with inverse FFTs the variant wins once the direct form needs more than %d lag tiles, i.e. L > ~%d samples "
                              "(%.2f m aperture at 50 kHz); at the reference's +-46 lags the direct form is %.1fx ahead of the contraction form") % (
        tiles, (128 * tiles - 16) // 2, ((128 * tiles - 16) // 2) / 50000.0 * 343.0, row["direct_over_gccphat_at_L46"])
    out["%d mics x %d" % (M, N)] = row
    print(M, N, row, flush=True)
json.dump(out, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "crossover.json"), "w"), indent=1)
