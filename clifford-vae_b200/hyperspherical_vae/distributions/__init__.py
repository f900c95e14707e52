from .von_mises_fisher import VonMisesFisher
from .hyperspherical_uniform import HypersphericalUniform

__all__ = ["VonMisesFisher", "HypersphericalUniform"]
