"""TEST INFRASTRUCTURE ONLY — import the *unmodified* reference hot-path modules.

Works only where the upstream checkout is mounted (the build container: /root/reference).  It is
used by `tests/golden/make_golden.py` to produce the committed fixtures and by the optional
`tests/test_oracle_vs_reference.py`; nothing on the GPU box needs it.

`import vsrd` in the reference eagerly pulls skimage / pycocotools (vsrd/__init__.py:1-11), which
are not installed, so we register empty namespace packages whose `__path__` points into the
checkout and import only the leaf modules the hot path needs.

The field-composition closures live inside `scripts/main.py::train` (main.py:433-523) and cannot be
imported; `main_closures()` extracts those nested function definitions from the script's AST and
compiles them *verbatim* against a caller-supplied namespace, so the goldens are produced by the
reference's own source text rather than by a paraphrase.
"""
from __future__ import annotations

import ast
import importlib
import os
import sys
import types

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _find_root() -> str:
    """The live checkout ($VSRD_REFERENCE_ROOT or /root/reference) if mounted, else the copy tools/stage_reference.py
    staged under baseline/_ref (git-ignored; it travels to the GPU box with the snapshot)."""
    candidates = [os.environ.get("VSRD_REFERENCE_ROOT"), "/root/reference", os.path.join(_REPO, "baseline", "_ref")]
    for root in candidates:
        if root and os.path.isfile(os.path.join(root, "scripts", "main.py")):
            return root
    return "/root/reference"


REFERENCE_ROOT = _find_root()
_PREFIX = "vsrd"  # the reference's own package name


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "scripts", "main.py"))


def _stub(name: str, rel: str):
    mod = types.ModuleType(name)
    mod.__path__ = [os.path.join(REFERENCE_ROOT, rel)]
    mod.__package__ = name
    sys.modules[name] = mod
    return mod


class reference_modules:
    """Context manager: temporarily bind `vsrd*` in sys.modules to the reference checkout.

    Our own drop-in package is also called `vsrd`; entering this context swaps it out and leaving
    restores it, so both can be used from one test process.
    """

    def __enter__(self):
        if not available():
            raise RuntimeError(f"reference checkout not found at {REFERENCE_ROOT}")
        self._saved = {k: v for k, v in sys.modules.items() if k == _PREFIX or k.startswith(_PREFIX + ".")}
        for k in self._saved:
            del sys.modules[k]
        _stub("vsrd", "vsrd")
        _stub("vsrd.models", "vsrd/models")
        _stub("vsrd.models.detectors", "vsrd/models/detectors")
        ns = types.SimpleNamespace()
        ns.utils = importlib.import_module("vsrd.utils")
        ns.rendering = importlib.import_module("vsrd.rendering")
        ns.fields = importlib.import_module("vsrd.models.fields")
        ns.encoders = importlib.import_module("vsrd.models.encoders.sinusoidal_encoder")
        ns.box_parameters = importlib.import_module("vsrd.models.detectors.box_parameters")
        ns.geometric_operations = importlib.import_module("vsrd.operations.geometric_operations")
        return ns

    def __exit__(self, *exc):
        for k in [k for k in sys.modules if k == _PREFIX or k.startswith(_PREFIX + ".")]:
            del sys.modules[k]
        sys.modules.update(self._saved)
        return False


_CLOSURE_NAMES = (
    "residual_distance_field", "residual_composition", "instance_field",
    "soft_union", "hard_union", "hierarchical_wrapper",
)


def main_closures(namespace: dict) -> dict:
    """Compile the nested factory functions of scripts/main.py:433-523 against `namespace`.

    `namespace` must provide the free variables those functions read at call time:
    `torch`, `nn`, `config`, `models`, `num_instances`.
    """
    path = os.path.join(REFERENCE_ROOT, "scripts", "main.py")
    with open(path) as f:
        tree = ast.parse(f.read(), filename=path)
    found = {}
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name in _CLOSURE_NAMES and node.name not in found:
            found[node.name] = node
    missing = set(_CLOSURE_NAMES) - set(found)
    if missing:
        raise RuntimeError(f"could not find {sorted(missing)} in {path}")
    module = ast.Module(body=[found[n] for n in _CLOSURE_NAMES], type_ignores=[])
    ast.fix_missing_locations(module)
    scope = dict(namespace)
    exec(compile(module, path, "exec"), scope)
    return {n: scope[n] for n in _CLOSURE_NAMES}
