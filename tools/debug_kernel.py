"""GPU debugging aid: run one kernel variant on the golden frames and report which products differ."""
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import audio_triangulation_b200 as at
from oracle_bindings import Oracle

kernel = sys.argv[1] if len(sys.argv) > 1 else "imma"
g = np.load(os.path.join(ROOT, "tests", "golden", "ref_vectors.npz"))
loc = at.Localizer(kernel=kernel)
adc = torch.from_numpy(g["adc"]).cuda()
r = loc.localize_device(adc, want=("lags", "raw", "windowed", "power", "corr"))
torch.cuda.synchronize()
r = {k: v.cpu().numpy() for k, v in r.items()}
o = Oracle().localize(g["adc"], want_raw=True)
print("windowed ok:", (r["windowed"] == g["after_window"]).all(), " power ok:", (r["power"] == g["power"]).all())
bad = np.argwhere(r["windowed"] != g["after_window"])
print("windowed mismatches:", len(bad), bad[:5].tolist())
if len(bad):
    f, m, i = bad[0]
    print(" got", r["windowed"][f, m, i:i+8], "exp", g["after_window"][f, m, i:i+8], "adc", g["adc"][f, m, i:i+8])
print("raw ok:", (r["raw"] == o["raw"]).all(), " lags ok:", (r["lags"] == o["lags"]).all())
badr = np.argwhere(r["raw"] != o["raw"])
print("raw mismatches:", len(badr), "of", r["raw"].size, badr[:6].tolist())
f = 3
for p in range(3):
    print("frame", f, "pair", p, "got", r["raw"][f, p, 40:52].tolist())
    print("             exp", o["raw"][f, p, 40:52].tolist())
    ratio = r["raw"][f, p].astype(float) / np.where(o["raw"][f, p] == 0, 1, o["raw"][f, p])
    print("   ratio", np.round(ratio[40:52], 4).tolist())
    # is got a shifted / reversed version of exp?
    e = o["raw"][f, p]; gt = r["raw"][f, p]
    for sh in range(-8, 9):
        if (np.roll(e, sh)[10:80] == gt[10:80]).all(): print("   == exp rolled by", sh)
    if (e[::-1][10:80] == gt[10:80]).all(): print("   == exp reversed")
print("lags got", r["lags"][:6].tolist(), "exp", o["lags"][:6].tolist())
