"""audio_triangulation_b200 -- B200-native per-frame localization path of Audio-Triangulation.

The compute lives in libat_b200.so (hand-written sm_100a CUDA behind a C ABI, include/at_b200.h).
This package is the thin host-side mirror used by tests and bench.py: torch supplies device
memory and streams, nothing else.
"""
from ._lib import AtError, load  # noqa: F401
from .api import Localizer, Stream, dropin, hemisphere_points  # noqa: F401

__all__ = ["Localizer", "Stream", "dropin", "hemisphere_points", "load", "AtError"]
