"""Per-role cycle accounting of the tcgen05 kernel (needs `make -C audio_triangulation_b200/csrc VARIANT=prof`).
Run as: AT_LIB_VARIANT=prof AT_PROF_PRINT=1 python tools/umma_prof.py [want ...]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import audio_triangulation_b200 as at
loc = at.Localizer(kernel="umma")
F = 1 << 19
flags = int(os.environ.get("AT_SYNTH_FLAGS", "0"))            # 16: white noise (every frame takes the exact pass and the full scan)
adc, _, _ = loc.synth_device(F, flags=flags)
want = tuple(sys.argv[1:]) or ("lags", "cell", "xy")
os.environ.pop("AT_PROF_PRINT", None)
out = {}
for _ in range(2): loc.localize_device(adc, want=want, out=out)
torch.cuda.synchronize()
os.environ["AT_PROF_PRINT"] = "1"
print("want =", want, " sections: mma-issue [loop, wait ready, wait empty, issue]; prep [loop, load+mean, wait planes free, prep+store, "
      "meta+fence+arrive]; epilogue [loop, wait full, TMEM+butterflies, spill+bar, arg-max, bar, decide+store]", flush=True)
loc.localize_device(adc, want=want, out=out)
torch.cuda.synchronize()
