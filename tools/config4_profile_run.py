"""Launch target for the ncu launch list of BASELINE config 4 (8 mics x 4096 samples): the kernel AUTO selects."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import audio_triangulation_b200 as at
loc = at.Localizer(n_mics=8, n_bits=12)
adc, _, _ = loc.synth_device(1 << 14)
out = {}
for _ in range(5): loc.localize_device(adc, want=("lags",), out=out)
torch.cuda.synchronize()
