"""Drop-in for the reference module ``dists.clifford`` (reference dists/clifford.py): same class
names and constructor signatures, implemented by clifford_b200.distributions on sm_100a kernels."""
from clifford_b200.distributions import (  # noqa: F401
    HypersphericalUniform,
    PowerSpherical,
    CliffordTorusUniform,
    CliffordTorusDistribution,
    CliffordPowerSphericalDistribution,
    _kl_ps_uniform,
    _kl_vm_uniform,
    _kl_powerspherical_uniform,
)
