"""GPU: the debug build of the library (make -C audio_triangulation_b200/csrc VARIANT=checked: -DAT_CHECKED turns on
in-kernel assertions on every shared-memory plane / curve / exchange index, TMEM slot ownership and the pipelines'
incremental phase counters; a failed one prints its location and traps) runs every kernel variant on ragged batches
with every output and must agree with the oracle.  It stands in for compute-sanitizer, which is closed on this pool."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_checked_build_runs_clean():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    # incremental: a no-op when __graft_entry__.build() has already produced an up-to-date libat_b200_checked.so
    subprocess.run(["make", "-j8", "-C", os.path.join(ROOT, "audio_triangulation_b200", "csrc"), "VARIANT=checked"], check=True)
    env = dict(os.environ, AT_LIB_VARIANT="checked")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sanitize_small.py")], capture_output=True, text=True,
                       timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "AT_CHECK failed" not in r.stdout + r.stderr
    assert "checked run ok" in r.stdout
