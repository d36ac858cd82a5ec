"""`pybnesian` - import-line alias of :mod:`pybnesian_b200`.

The reference is ONE flat pybind11 module named ``pybnesian`` (/root/reference/pybnesian/lib.cpp:22-51); code written
against it (``import pybnesian as pbn``; the reference's own tests/ directory) runs unmodified on the B200 path with
this package first on ``sys.path``.  Everything is re-exported from pybnesian_b200; nothing is implemented here.
"""
from pybnesian_b200 import *  # noqa: F401,F403
from pybnesian_b200 import __version__, parallel, default_context, Context  # noqa: F401
import pybnesian_b200 as _impl

__all__ = [n for n in dir(_impl) if not n.startswith("_")]


def __getattr__(name):
    # classes of the reference that lie outside the KDE / CKDE hot path (out of scope) are absent on purpose
    raise AttributeError("pybnesian (B200 hot-path build) has no attribute %r: outside the scope of SURVEY.md section 8" % name)
