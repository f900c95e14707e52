"""Drop-in for reference vmf/hyperspherical_vae/distributions/von_mises_fisher.py."""
from clifford_b200.vmf import VonMisesFisher, _kl_vmf_uniform  # noqa: F401
