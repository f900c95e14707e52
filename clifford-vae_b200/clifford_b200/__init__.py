"""clifford_b200 -- B200-native (sm_100a) latent hot path of momalekabid/clifford-vae.

Python host side over the C-ABI library ``libclifford_b200.so`` (include/clifford_b200.h):
``distributions`` mirrors reference ``dists/clifford.py``, ``vmf`` mirrors
``hyperspherical_vae.distributions``, ``vsa`` mirrors ``utils/vsa.py:9-96``.
"""
from . import _lib  # noqa: F401

__all__ = ["_lib", "ops", "distributions", "vsa", "vmf"]
