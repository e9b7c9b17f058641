"""Drop-in module names of the reference (``from matryodshka.msi import MSI``, /root/reference/test.py:27,
matryodshka/msi.py:26-31): thin aliases of the ``matryodshka_b200`` mirror.  Nothing is implemented here."""
