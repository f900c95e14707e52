"""Draw injection for parity tests of UNMODIFIED callers.

The reference's models call ``q_z.rsample()`` themselves (mnist/mlp_vae.py:110, cnn/models.py:231), so a test that
runs them unchanged cannot pass the private ``_base_draws=`` argument.  ``with injected_draws(d1, d2, ...)`` queues
base-variate tuples; every ``rsample()`` of a drop-in distribution that is called without explicit draws consumes the
next one (thread-local, FIFO).  Outside such a block nothing is queued and ``rsample`` uses the device generator.
"""
from __future__ import annotations

import contextlib
import threading

_state = threading.local()


def take():
    q = getattr(_state, "queue", None)
    if q:
        return q.pop(0)
    return None


@contextlib.contextmanager
def injected_draws(*draws):
    old = getattr(_state, "queue", None)
    _state.queue = list(draws)
    try:
        yield
    finally:
        _state.queue = old
