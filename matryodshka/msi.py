"""``matryodshka.msi`` under the reference's module name (reference: matryodshka/msi.py, class MSI :33).
Alias of ``matryodshka_b200.msi``: the sm_100a kernels behind torch.ops.msi.* / include/msi_b200.h."""
from matryodshka_b200.msi import MSI, MSIConfig  # noqa: F401

__all__ = ["MSI", "MSIConfig"]
