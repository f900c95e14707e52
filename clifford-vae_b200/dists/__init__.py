"""Drop-in for the reference package ``dists`` (reference dists/__init__.py:1-15)."""
from dists.clifford import (
    PowerSpherical,
    HypersphericalUniform,
    CliffordTorusUniform,
    CliffordTorusDistribution,
    CliffordPowerSphericalDistribution,
)

__all__ = [
    "PowerSpherical",
    "HypersphericalUniform",
    "CliffordTorusUniform",
    "CliffordTorusDistribution",
    "CliffordPowerSphericalDistribution",
]
