"""Stage the files of the reference checkout that the tests / bench need on the GPU box under `baseline/_ref/`
(git-ignored, NOT gpurun-ignored: it travels with the snapshot like the built .so).

Nothing is copied into the tracked tree.  Staged: `scripts/main.py` (run unmodified by tools/run_main.py) and the
hot-path modules of the reference's `vsrd` package that `oracle/ref_import.py` imports unmodified for the CPU
reference arm of bench.py.  A manifest with SHA-256 digests is written so the tests can prove "unmodified".
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE_ROOT = os.environ.get("VSRD_REFERENCE_ROOT", "/root/reference")
STAGE_ROOT = os.path.join(ROOT, "baseline", "_ref")
FILES = [
    "scripts/main.py",
    "vsrd/utils.py",
    "vsrd/rendering/__init__.py", "vsrd/rendering/renderers.py", "vsrd/rendering/samplers.py",
    "vsrd/rendering/sdfs.py", "vsrd/rendering/utils.py",
    "vsrd/models/fields/__init__.py", "vsrd/models/fields/hyper_distance_field.py",
    "vsrd/models/fields/hyper_radiance_field.py",
    "vsrd/models/encoders/__init__.py", "vsrd/models/encoders/sinusoidal_encoder.py",
    "vsrd/models/encoders/tensorial_encoder.py",
    "vsrd/models/detectors/box_parameters.py",
    "vsrd/operations/__init__.py", "vsrd/operations/geometric_operations.py", "vsrd/operations/kitti360_operations.py",
    "LICENSE",
] + [f"vsrd/modules/{name}.py" for name in (     # imported by models/encoders/tensorial_encoder.py at package import
    "__init__", "attention", "drop_path", "grad_scale", "grid_sampler", "layer_scale", "packing_block",
    "plane_sweep_stereo", "sinkhorn_knopp", "spatial_propagation", "squeeze_excitation", "utils")]
# digest of the script the parity / drop-in claims are about (skmhrk1209/VSRD @ 68765a4)
MAIN_PY_SHA256 = "45368120015307176e46484d54b3a44585a4d2db774e1aba511e0ea4f3b851f5"


def sha256(path: str) -> str:
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def reference_root() -> str | None:
    """The live checkout if mounted, else the staged copy, else None."""
    for root in (REFERENCE_ROOT, STAGE_ROOT):
        if os.path.isfile(os.path.join(root, "scripts", "main.py")):
            return root
    return None


def stage() -> str | None:
    if not os.path.isfile(os.path.join(REFERENCE_ROOT, "scripts", "main.py")):
        return None
    manifest = {}
    for rel in FILES:
        src, dst = os.path.join(REFERENCE_ROOT, rel), os.path.join(STAGE_ROOT, rel)
        if not os.path.isfile(src):
            continue
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if not os.path.isfile(dst) or sha256(dst) != sha256(src):
            shutil.copyfile(src, dst)
        manifest[rel] = sha256(dst)
    with open(os.path.join(STAGE_ROOT, "MANIFEST.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)
    return STAGE_ROOT


if __name__ == "__main__":
    print(stage())
