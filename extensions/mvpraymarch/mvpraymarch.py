"""Overlay module: put this repository in front of an unmodified ava-256 checkout on PYTHONPATH and
`from extensions.mvpraymarch.mvpraymarch import mvpraymarch` (models/raymarchers/mvpraymarcher.py:14) resolves
here -- `extensions/` has no __init__.py in the reference either, so it is a namespace package and
`extensions.utils` keeps resolving to the reference's ray generator.

`mvpraymarch` is a plain Python function with the reference's exact parameter list because its caller filters
render options by `mvpraymarch.__code__.co_varnames` (mvpraymarcher.py:45).
"""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from ava256_b200.op import MVPRaymarch, mvpraymarch  # noqa: E402,F401
