"""Permissive stand-in for matplotlib (not installed in this image) so the reference's drivers and
scripts, which import it at module level (utils/vsa.py:4, utils/wandb_utils.py:6), can be imported by
the boundary tests.  Every attribute of every submodule is a MagicMock: nothing is ever drawn."""
import sys
import types
from unittest.mock import MagicMock


class _Stub(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return MagicMock(name=f"{self.__name__}.{name}")


_me = sys.modules[__name__]
for _sub in ("pyplot", "colors", "image", "gridspec", "cm", "patches", "lines", "ticker", "font_manager", "animation", "widgets"):
    _m = _Stub(f"matplotlib.{_sub}")
    sys.modules[f"matplotlib.{_sub}"] = _m
    setattr(_me, _sub, _m)


def __getattr__(name):
    if name.startswith("__"):
        raise AttributeError(name)
    return MagicMock(name=f"matplotlib.{name}")
