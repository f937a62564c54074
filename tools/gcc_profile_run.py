"""Launch the GCC-PHAT kernels a few times (target of an ncu capture).  usage: python tools/gcc_profile_run.py [mics bits frames]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import audio_triangulation_b200 as at
M, nb, F = (int(x) for x in sys.argv[1:4]) if len(sys.argv) > 3 else (8, 12, 256)
loc = at.Localizer(n_mics=M, n_bits=nb)
adc, _, _ = loc.synth_device(F)
for _ in range(3): loc.gccphat_device(adc)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10): loc.gccphat_device(adc)
b.record(); torch.cuda.synchronize()
print("%d mics x %d samples, %d frames: %.3f M frames/s" % (M, 1 << nb, F, F * 10 / a.elapsed_time(b) / 1e3))
loc.close()
