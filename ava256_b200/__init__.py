"""Importable alias of the hyphen-named package directory `ava-256_b200/` (a hyphen is not a valid identifier).

`import ava256_b200.op` resolves `ava-256_b200/op.py`; nothing else lives here.
"""
import os as _os

__path__.append(_os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "..", "ava-256_b200"))
from .version import __version__  # noqa: E402,F401
