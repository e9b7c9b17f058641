"""matryodshka_b200 -- Blackwell-native multi-sphere-image (MSI) inference path.

Host-side mirror of brownvc/matryodshka's ``matryodshka.msi.MSI`` and
``geometry.{projector,spherical,sampling}`` call surface over hand-written
sm_100a CUDA kernels (``csrc/``) behind the C-ABI of ``include/msi_b200.h``.
No CPU fallback: every compute entry point raises if the CUDA library is
missing or no GPU is present.
"""
__version__ = "0.1.0"
