"""GPU: the plain-C host harness (harness/at_harness.c, linked against libat_b200.so's C ABI) runs the reference's call
sequence through the drop-in symbols, then the batched path, and reports that both agree."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_harness_runs_and_agrees():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    exe = os.path.join(ROOT, "harness", "at_harness")
    if not os.path.exists(exe):
        subprocess.run(["make", "-C", os.path.join(ROOT, "harness")], check=True)
    r = subprocess.run([exe, "8192"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "[agree]" in r.stdout and "[MISMATCH]" not in r.stdout
    m = re.search(r"([0-9.]+) % within 5 cells of the source; (\d+) kernel launches", r.stdout)
    assert m, r.stdout
    assert float(m.group(1)) > 80.0          # the synthetic sources are found
    assert int(m.group(2)) >= 3              # CUDA kernels did the work (no CPU path exists)
