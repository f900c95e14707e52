"""Drop-in for the hot-path part of the reference module ``utils.vsa`` (reference utils/vsa.py:9-96)."""
from clifford_b200.vsa import (  # noqa: F401
    hrr_init,
    unitary_init,
    normalize_vectors,
    bind,
    invert,
    unbind,
    bundle,
    permute_vector,
    unpermute_vector,
    similarity,
)
