"""Drop-in for the reference package ``utils`` restricted to the hot path (``utils.vsa``).

``utils.wandb_utils`` (plots, logging -- out of scope here) keeps loading from the reference checkout: its
``utils/`` directory is appended to this package's search path.  The checkout is found through the environment
variable CLIFFORD_VAE_REFERENCE_ROOT or, failing that, on ``sys.path`` (every reference driver appends its
repository root there before importing ``utils.*``: mnist/mnist_clifpws.py:18, cnn/cifar10_train.py:21).
"""
import os as _os
import sys as _sys

from .vsa import bind as vsa_bind, unbind as vsa_unbind, invert as vsa_invert  # noqa: F401

_here = _os.path.dirname(_os.path.abspath(__file__))


def _reference_utils_dir():
    roots = [_os.environ.get("CLIFFORD_VAE_REFERENCE_ROOT")] + list(_sys.path)
    for root in roots:
        if not root:
            continue
        cand = _os.path.join(_os.path.abspath(root), "utils")
        if cand != _here and _os.path.isfile(_os.path.join(cand, "wandb_utils.py")):
            return cand
    return None


_ref_utils = _reference_utils_dir()
if _ref_utils is not None and _ref_utils not in __path__:
    __path__.append(_ref_utils)


def __getattr__(name):
    """The reference's ``utils/__init__.py`` also re-exports logging helpers from ``wandb_utils`` and a few
    one-liners; resolve them lazily from the reference checkout so ``from utils import WandbLogger`` keeps working."""
    if name in ("WandbLogger", "test_self_binding", "compute_class_means", "evaluate_mean_vector_cosine",
                "test_vsa_operations"):
        from . import wandb_utils
        return getattr(wandb_utils, name)
    raise AttributeError(f"module 'utils' has no attribute {name!r}")
