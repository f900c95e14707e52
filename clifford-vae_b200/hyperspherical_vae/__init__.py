"""Drop-in for the reference package ``hyperspherical_vae`` (reference vmf/hyperspherical_vae)."""
from . import distributions
from . import ops

__all__ = ["distributions", "ops"]
