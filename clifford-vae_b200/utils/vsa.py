"""Drop-in for the reference module ``utils.vsa``: the ten ops (reference utils/vsa.py:9-96) and the three
experiment-harness entry points every driver imports from it (utils/vsa.py:99-167, 224-398, 402-630)."""
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
from clifford_b200.harness import (  # noqa: F401
    test_bundle_capacity,
    test_binding_unbinding_pairs,
    test_per_class_bundle_capacity_k_items,
)
