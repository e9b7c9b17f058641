"""``geometry.sampling`` under the reference's module name (reference: geometry/sampling.py).
Alias of ``matryodshka_b200.geometry.sampling``."""
from matryodshka_b200.geometry.sampling import *  # noqa: F401,F403
from matryodshka_b200.geometry import sampling as _impl

__all__ = [n for n in dir(_impl) if not n.startswith("_")]
