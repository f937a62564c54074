"""Kernel-only timing of the fused path at 2^19 frames for lags-only and lags+cell+xy."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import audio_triangulation_b200 as at
loc = at.Localizer(kernel=sys.argv[1] if len(sys.argv) > 1 else "auto")
F = 1 << 19
adc, _, _ = loc.synth_device(F)
for want in (("lags",), ("lags", "cell", "xy")):
    out = {}
    for _ in range(3): loc.localize_device(adc, want=want, out=out)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10): loc.localize_device(adc, want=want, out=out)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 10
    print("want=%-22s %.3f ms  %.1f Mframes/s  (skip=%s)" % ("+".join(want), ms, F / ms / 1e3, os.environ.get("AT_DEBUG_SKIP", "0")))
st = loc.localize_device(adc, want=("lags", "cell", "xy", "stats"))["stats"]
torch.cuda.synchronize()
print("search routes [first box, widened, full scan, peak-tuple look-up, of which certified without l.l]:", st.cpu().tolist())
