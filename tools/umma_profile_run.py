"""Launch a fused kernel a few times (target of an ncu capture).
usage: python tools/umma_profile_run.py [3|8] [want ...]   3: reference shape (3 mics x 1024), 8: config 4 (8 mics x 4096)"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import audio_triangulation_b200 as at
which = sys.argv[1] if len(sys.argv) > 1 else "3"
want = tuple(sys.argv[2:]) or ("lags", "cell", "xy")
M, nb, F = (8, 12, 1 << 13) if which == "8" else (3, 10, 1 << 18)
loc = at.Localizer(kernel=os.environ.get("AT_KERNEL", "auto"), n_mics=M, n_bits=nb)
adc, _, _ = loc.synth_device(F)
out = {}
for _ in range(3): loc.localize_device(adc, want=want if M == 3 else ("lags",), out=out)
torch.cuda.synchronize()
loc.close()
