import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.dirname(os.path.abspath(__file__))):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle_bindings import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def ref():
    """The reference's own compiled objects (oracle/_ref/libat_ref.so); skip if never built."""
    from oracle_bindings import load_ref
    lib = load_ref()
    if lib is None:
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    return lib


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "ref_vectors.npz"))


@pytest.fixture(scope="session")
def golden_hm():
    """Lag LUT and per-cell colour classes produced by the reference's own vga_heatmap.h (tests/golden/make_heatmap_golden.py)."""
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "ref_heatmap.npz"))


@pytest.fixture(scope="session")
def loc():
    """Reference-shape Localizer on cuda:0 (GPU tests only)."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import audio_triangulation_b200 as at
    return at.Localizer(device=0)
