"""Drop-in module names of the reference's ``geometry`` package (``import geometry.projector as pj``,
/root/reference/matryodshka/msi.py:26-31): thin aliases of ``matryodshka_b200.geometry``."""
from . import projector, sampling, spherical  # noqa: F401
