"""Drop-in for reference vmf/hyperspherical_vae/ops: the Bessel-ratio bound used by VonMisesFisher.mean.
The reference's IveFunction (host SciPy, ops/ive.py:7-46) is replaced by the device kernel behind
VonMisesFisher.entropy()/_log_normalization()."""
from clifford_b200.vmf import _ive_fraction_approx2 as ive_fraction_approx2  # noqa: F401

__all__ = ["ive_fraction_approx2"]
