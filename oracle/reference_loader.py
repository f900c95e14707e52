"""TEST / BENCH INFRASTRUCTURE -- import the UNMODIFIED reference modules (from the checkout when present, else
from the copy `oracle/stage_reference.py` staged under `oracle/_ref/`) under private module names, so they never
shadow the drop-in packages `dists`, `utils`, `hyperspherical_vae` of this repository.

    ref = load()            # None when neither location exists
    ref.clifford            # dists/clifford.py           (CliffordPowerSphericalDistribution, PowerSpherical, ...)
    ref.vsa                 # utils/vsa.py                (bind, unbind, bundle, ..., test_* harnesses)
    ref.vmf                 # hyperspherical_vae.distributions (VonMisesFisher, HypersphericalUniform)
    ref.root, ref.kind      # where it came from: "checkout" | "staged"
    models(ref, "cnn", dists=<module>)   # cnn/models.py or mnist/mlp_vae.py bound to a chosen `dists.clifford`
"""
from __future__ import annotations

import importlib
import importlib.util
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
STAGED = os.path.join(HERE, "_ref")
CHECKOUT = os.environ.get("CLIFFORD_VAE_REFERENCE_ROOT", "/root/reference")

_cache = {}


def _stub_matplotlib():
    try:
        import matplotlib.pyplot  # noqa: F401
        return
    except Exception:
        pass
    for m in ("matplotlib", "matplotlib.pyplot"):        # utils/vsa.py:4 imports it at module level
        sys.modules.setdefault(m, types.ModuleType(m))


def _load_file(name: str, path: str):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def find_root():
    for root, kind in ((CHECKOUT, "checkout"), (STAGED, "staged")):
        if os.path.isfile(os.path.join(root, "dists", "clifford.py")) and os.path.isfile(os.path.join(root, "utils", "vsa.py")):
            return root, kind
    return None, None


def load():
    if "ref" in _cache:
        return _cache["ref"]
    root, kind = find_root()
    if root is None:
        _cache["ref"] = None
        return None
    _stub_matplotlib()
    ref = types.SimpleNamespace(root=root, kind=kind)
    ref.clifford = _load_file("_cvref_dists_clifford", os.path.join(root, "dists", "clifford.py"))
    ref.vsa = _load_file("_cvref_utils_vsa", os.path.join(root, "utils", "vsa.py"))
    # the vMF package imports itself by its absolute name (von_mises_fisher.py:5-8): import it under that name with
    # the drop-in `hyperspherical_vae` modules set aside, then re-register it under a private prefix
    vmf_dir = os.path.join(root, "vmf")
    ref.vmf = None
    if os.path.isdir(os.path.join(vmf_dir, "hyperspherical_vae")):
        def _is_h(k):
            return k == "hyperspherical_vae" or k.startswith("hyperspherical_vae.")
        saved = {k: sys.modules.pop(k) for k in list(sys.modules) if _is_h(k)}
        sys.path.insert(0, vmf_dir)
        try:
            ref.vmf = importlib.import_module("hyperspherical_vae.distributions")
        finally:
            sys.path.remove(vmf_dir)
            loaded = {k: sys.modules.pop(k) for k in list(sys.modules) if _is_h(k)}
            sys.modules.update(saved)
            for k, v in loaded.items():
                sys.modules["_cvref_" + k] = v
    _cache["ref"] = ref
    return ref


def models(ref, which: str, dists_clifford=None, hvae_distributions=None):
    """Load the reference's model file (`cnn` -> cnn/models.py, `mnist` -> mnist/mlp_vae.py) with its
    `from dists.clifford import ...` (cnn/models.py:10-15, mnist/mlp_vae.py:11-16) and its lazy
    `from hyperspherical_vae.distributions import ...` (mlp_vae.py:85-88) resolved to the given modules: the drop-in
    ones (default: whatever `dists.clifford` imports to, i.e. this repository's when it is on sys.path) or the
    reference's own (pass ref.clifford / ref.vmf) for the reference-on-GPU comparator."""
    rel = {"cnn": ("cnn", "models.py"), "mnist": ("mnist", "mlp_vae.py")}[which]
    path = os.path.join(ref.root, *rel)
    tag = "ref" if dists_clifford is not None else "dropin"
    name = f"_cvref_{which}_models_{tag}"
    if name in sys.modules:
        return sys.modules[name]
    saved = {k: sys.modules.get(k) for k in ("dists", "dists.clifford", "hyperspherical_vae",
                                             "hyperspherical_vae.distributions",
                                             "hyperspherical_vae.distributions.hyperspherical_uniform")}
    try:
        if dists_clifford is not None:
            pkg = types.ModuleType("dists")
            pkg.clifford = dists_clifford
            pkg.__path__ = []
            sys.modules["dists"], sys.modules["dists.clifford"] = pkg, dists_clifford
        mod = _load_file(name, path)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return mod
