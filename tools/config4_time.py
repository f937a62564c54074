"""BASELINE config 4 timing: 8 mics (28 pairs) x 4096-sample frames, direct integer cross-correlation on the
integer pipe (imad), on the legacy tensor-core path (imma, mma.sync, CTA per frame) and on tcgen05 (umma, polyphase
Hankel form, 8 mics only).  No reference counterpart (parity unpinned)."""
import sys, os, json, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import audio_triangulation_b200 as at
res = {}
for M, nb, F in ((8, 12, 1 << 14), (8, 10, 1 << 16), (4, 10, 1 << 17)):
    for kernel in ("imad", "imma", "umma") if M == 8 else ("imad", "imma"):
        loc = at.Localizer(kernel=kernel, n_mics=M, n_bits=nb)
        adc, _, _ = loc.synth_device(F)
        out = {}
        for _ in range(2): loc.localize_device(adc, want=("lags",), out=out)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5): loc.localize_device(adc, want=("lags",), out=out)
        b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 5
        P = M * (M - 1) // 2
        macs = P * sum((1 << nb) - abs(s) for s in range(-46, 47))
        res["%dx%d %s" % (M, 1 << nb, kernel)] = {"frames_per_s": F / ms * 1e3, "ms": ms, "useful_int16_TMAC_per_s": macs * F / ms / 1e9}
        print(M, 1 << nb, kernel, "%.3f ms  %.3f Mframes/s  %.1f T int16-MAC/s" % (ms, F / ms / 1e3, macs * F / ms / 1e9), flush=True)
        lags_ref = out["lags"].clone() if kernel == "imad" else lags_ref
        if kernel != "imad": print("   lags identical to imad:", bool(torch.equal(out["lags"], lags_ref)))
        loc.close()
json.dump(res, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "config4_time.json"), "w"), indent=1)
