"""Host-side mirror of the reference's ``geometry`` package (projector, spherical, sampling)
for the MSI inference path."""
from . import projector, sampling, spherical  # noqa: F401
