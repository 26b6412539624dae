"""Run the reference's UNMODIFIED `scripts/main.py` on top of the drop-in `vsrd` package.

    python tools/run_main.py --config configs/synthetic/vsrd/drive_0000_synthetic/config.json --train \
        [--set optimization.num_steps=60 --set logging.scalar_intervals=10 ...] [--workdir DIR] [--main PATH]

What this launcher does — and all it does:
  * finds main.py (`--main`, else $VSRD_REFERENCE_ROOT/scripts/main.py, else the staged baseline/_ref/scripts/main.py),
    checks its SHA-256 against the digest of the upstream file (tools/stage_reference.py) and refuses anything else;
  * puts this repository first on sys.path so `import vsrd` resolves to the drop-in package, and tools/shims on the
    path only if `inflection` is not installed;
  * fills in the torchrun environment for a single process when it is not launched by torchrun;
  * `--set a.b=c` writes a derived config (JSON leaf overrides) to `<workdir>/configs/<...>/config.json`; main.py derives
    its ckpts/ logs/ outs/ directories from the config path (main.py:130-132), so outputs land under `<workdir>`;
  * executes the script with `runpy.run_path(..., run_name="__main__")` and argv `--launcher torchrun --config ... --train`.
The script's source is never edited, patched or wrapped.
"""
from __future__ import annotations

import argparse
import json
import os
import runpy
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from tools import stage_reference  # noqa: E402


def find_main(explicit=None) -> str:
    if explicit:
        path = explicit
    else:
        root = stage_reference.reference_root()
        if root is None:
            raise SystemExit("run_main: scripts/main.py not found (neither $VSRD_REFERENCE_ROOT nor baseline/_ref); "
                             "run tools/stage_reference.py where the reference checkout is mounted")
        path = os.path.join(root, "scripts", "main.py")
    digest = stage_reference.sha256(path)
    if digest != stage_reference.MAIN_PY_SHA256:
        raise SystemExit(f"run_main: {path} is not the upstream scripts/main.py (sha256 {digest})")
    return path


def _set_leaf(config: dict, dotted: str, value):
    node = config
    *parents, leaf = dotted.split(".")
    for key in parents:
        node = node[key]
    if leaf not in node and parents[-1:] != ["kwargs"]:           # kwargs may gain keys; anything else must exist
        raise SystemExit(f"run_main: --set {dotted}: no such config leaf")
    node[leaf] = value


def derive_config(config_path: str, overrides, workdir: str) -> str:
    """Copy the config tree entry to `<workdir>/configs/...` with the leaf overrides applied."""
    with open(config_path) as f:
        config = json.load(f)
    for item in overrides:
        dotted, _, raw = item.partition("=")
        try:
            value = json.loads(raw)
        except json.JSONDecodeError:
            value = raw
        _set_leaf(config, dotted, value)
    parts = os.path.abspath(config_path).split(os.sep)
    tail = parts[parts.index("configs"):] if "configs" in parts else ["configs", os.path.basename(config_path)]
    out = os.path.join(workdir, *tail)
    os.makedirs(os.path.dirname(out), exist_ok=True)
    with open(out, "w") as f:
        json.dump(config, f, indent=4)
    return out


def run(config: str, overrides=(), workdir=None, main=None, train=True, device_id=0) -> str:
    """Returns the config path main.py was given (its outputs sit beside it under ckpts/ logs/ outs/)."""
    main_path = find_main(main)
    if workdir or overrides:
        config = derive_config(config, overrides, workdir or os.path.join(ROOT, "gpurun_out", "main_py"))
    try:
        import inflection  # noqa: F401
    except ImportError:
        sys.path.append(os.path.join(ROOT, "tools", "shims"))
    os.environ.setdefault("RANK", "0")
    os.environ.setdefault("LOCAL_RANK", "0")
    os.environ.setdefault("WORLD_SIZE", "1")
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29533")
    argv = [main_path, "--launcher", "torchrun", "--config", config, "--device_id", str(device_id)] + (["--train"] if train else [])
    saved = sys.argv
    sys.argv = argv
    try:
        runpy.run_path(main_path, run_name="__main__")
    finally:
        sys.argv = saved
    return config


if __name__ == "__main__":
    parser = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    parser.add_argument("--config", required=True)
    parser.add_argument("--train", action="store_true")
    parser.add_argument("--set", dest="overrides", action="append", default=[], metavar="key.path=json")
    parser.add_argument("--workdir", default=None)
    parser.add_argument("--main", default=None)
    parser.add_argument("--device_id", type=int, default=0)
    args = parser.parse_args()
    run(args.config, args.overrides, args.workdir, args.main, args.train, args.device_id)
