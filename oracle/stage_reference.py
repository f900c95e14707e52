"""TEST / BENCH INFRASTRUCTURE -- stage the reference's hot-path and model files so they travel to the GPU box.

The reference is pure Python; `gpurun` ships only this repository, so `/root/reference` does not exist where the
benchmarks run.  `__graft_entry__.build()` calls `stage()` in the build container: the files listed below are copied
VERBATIM from the reference checkout into `oracle/_ref/` (git-ignored -> never in history, not gpurun-ignored -> it
travels like a built .so).  `bench.py --impl reference` / `cpu_baseline` then time the reference's OWN classes
(`dists/clifford.py:295-327`, `utils/vsa.py:43-46`; `cpu_baseline.kind = "reference"`), and the `vae_train_step` leg
and tests/test_gpu_reference_models.py run the reference's own models (`cnn/models.py:134-315`,
`mnist/mlp_vae.py:19-190`) on top of the drop-in distributions.  The product package never imports any of it.

    python oracle/stage_reference.py [reference_root]
"""
from __future__ import annotations

import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
DEFAULT_REF = os.environ.get("CLIFFORD_VAE_REFERENCE_ROOT", "/root/reference")

# hot path + the two model files that call it (no drivers, no plotting / logging utilities)
FILES = [
    "dists/__init__.py",
    "dists/clifford.py",
    "utils/vsa.py",
    "vmf/hyperspherical_vae/__init__.py",
    "vmf/hyperspherical_vae/distributions/__init__.py",
    "vmf/hyperspherical_vae/distributions/von_mises_fisher.py",
    "vmf/hyperspherical_vae/distributions/hyperspherical_uniform.py",
    "vmf/hyperspherical_vae/ops/__init__.py",
    "vmf/hyperspherical_vae/ops/ive.py",
    "mnist/__init__.py",
    "mnist/mlp_vae.py",
    "cnn/__init__.py",
    "cnn/models.py",
]


def stage(ref_root: str = DEFAULT_REF, dest: str = DEST) -> bool:
    """Copy FILES from ref_root to dest.  Returns False (and leaves dest alone) when the checkout is absent."""
    if not os.path.isfile(os.path.join(ref_root, "dists", "clifford.py")):
        return False
    for rel in FILES:
        src = os.path.join(ref_root, rel)
        if not os.path.isfile(src):
            continue
        dst = os.path.join(dest, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
    with open(os.path.join(dest, "STAGED_FROM"), "w") as f:
        f.write(ref_root + "\n")
    return True


if __name__ == "__main__":
    ok = stage(sys.argv[1] if len(sys.argv) > 1 else DEFAULT_REF)
    print("staged into", DEST if ok else "(nothing: reference checkout not found)")
