"""``matryodshka.nets`` under the reference's module name (reference: matryodshka/nets.py:387-534).
Alias of ``matryodshka_b200.nets``."""
from matryodshka_b200.nets import (ARCH, layer_channels, layer_geometry, layer_shapes, msi_coord_train_net,  # noqa: F401
                                   msi_train_net, net_flops)

__all__ = ["msi_coord_train_net", "msi_train_net", "ARCH", "layer_channels", "layer_geometry", "layer_shapes", "net_flops"]
