"""``geometry.spherical`` under the reference's module name (reference: geometry/spherical.py).
Alias of ``matryodshka_b200.geometry.spherical``."""
from matryodshka_b200.geometry.spherical import *  # noqa: F401,F403
from matryodshka_b200.geometry import spherical as _impl

__all__ = [n for n in dir(_impl) if not n.startswith("_")]
