"""Drop-in for reference vmf/hyperspherical_vae/distributions/hyperspherical_uniform.py."""
from clifford_b200.vmf import HypersphericalUniform  # noqa: F401
