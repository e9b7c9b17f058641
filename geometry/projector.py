"""``geometry.projector`` under the reference's module name (reference: geometry/projector.py).
Alias of ``matryodshka_b200.geometry.projector``."""
from matryodshka_b200.geometry.projector import *  # noqa: F401,F403
from matryodshka_b200.geometry import projector as _impl

__all__ = [n for n in dir(_impl) if not n.startswith("_")]
