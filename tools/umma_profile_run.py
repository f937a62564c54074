"""Launch the tcgen05 kernels a few times (target of an ncu capture): 8 mics x 4096 (config 4) and 3 mics x 1024."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import audio_triangulation_b200 as at
for M, nb, F in ((8, 12, 1 << 13), (3, 10, 1 << 18)):
    loc = at.Localizer(kernel="umma", n_mics=M, n_bits=nb)
    adc, _, _ = loc.synth_device(F)
    out = {}
    for _ in range(3): loc.localize_device(adc, want=("lags",), out=out)
    torch.cuda.synchronize()
    loc.close()
