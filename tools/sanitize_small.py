"""Small all-paths run for compute-sanitizer (memcheck): every kernel variant, every output, heads, streaming."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import audio_triangulation_b200 as at
ALL = ("lags", "corr", "raw", "cell", "highest", "xy", "gate", "classes", "windowed", "power")
for kernel in ("imma", "umma", "imad"):
    loc = at.Localizer(kernel=kernel)
    adc, heads, _ = loc.synth_device(67, flags=2 | 4)
    r = loc.localize_device(adc, heads, want=ALL)
    r2 = loc.localize_device(adc, None, want=("lags", "cell", "xy"))
    torch.cuda.synchronize()
    print(kernel, r["lags"][:2].tolist(), int(r["cell"][0]))
loc = at.Localizer()
st = at.Stream(loc, 5)
x = torch.randint(100, 156, (5, 1024, 3), dtype=torch.uint8, device="cuda")
for _ in range(3):
    st.push(x)
est = torch.zeros((5, 3, 93), dtype=torch.int64, device="cuda")
loc.heatmap_device(est, want=("cell", "highest", "xy", "classes"))
h = loc.localize_host(adc.cpu().numpy(), heads.cpu().numpy(), want=("lags", "cell"))
d = at.dropin
b = np.zeros(1, at.api.BUFFER_DT); d.buffer_window(b); d.buffer_normalize_range(b)
c = np.zeros(1, at.api.CORR_DT); d.correlations_init(c, b, b); d.correlations_average(c, c)
torch.cuda.synchronize()
print("done")
