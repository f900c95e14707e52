"""Drop-in for the reference package ``utils`` restricted to the hot path (``utils.vsa``).

When the environment variable CLIFFORD_VAE_REFERENCE_ROOT points at a checkout of the reference,
its ``utils/`` directory is appended to this package's search path so ``utils.wandb_utils`` (plots,
logging -- out of scope here) still resolves while ``utils.vsa`` comes from this repo.
"""
import os as _os

from .vsa import bind as vsa_bind, unbind as vsa_unbind, invert as vsa_invert  # noqa: F401

_ref = _os.environ.get("CLIFFORD_VAE_REFERENCE_ROOT")
if _ref and _os.path.isdir(_os.path.join(_ref, "utils")):
    __path__.append(_os.path.join(_ref, "utils"))
