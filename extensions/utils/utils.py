"""Overlay module for `from extensions.utils.utils import compute_raydirs` (models/autoencoder.py:19, :240): with this
repository in front of an unmodified ava-256 checkout on PYTHONPATH the ray generator, too, runs on the B200 library
(no reference-native code left on the render path)."""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from ava256_b200.raydirs import ComputeRaydirs, compute_raydirs  # noqa: E402,F401
