#!/usr/bin/env python
"""Final checkpoints -> per-frame prediction JSON with confidences, for sequences labelled from the synthetic dataset
(the counterpart of tools/kitti_360/make_predictions.py; arithmetic in vsrd_b200/predictions.py).

    python tools/make_predictions.py --config <config.json used for the run> --ckpt-root <.../ckpts/...> [--out DIR]

Reads `<ckpt-root>/**/step_<num_steps-1>.pt` as scripts/main.py (or tools/label_sequence.py --ckpt-dir) wrote them,
finds each frame in the dataset by its file name, and writes `<out>/<frame>/<relative index>.json`."""
import argparse
import glob
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import vsrd  # noqa: E402
from vsrd_b200 import predictions  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", required=True)
    ap.add_argument("--ckpt-root", required=True)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    config = vsrd.utils.Dict.apply(vsrd.configuration.Configurator.load(args.config))
    dataset = vsrd.utils.import_module(config.datasets.train, globals(), locals())
    last = f"step_{config.optimization.num_steps - 1}.pt"
    out_root = args.out or os.path.join(os.path.dirname(args.ckpt_root.rstrip("/")), "predictions")
    by_name = {os.path.splitext(os.path.basename(dataset.frame_filename(i)))[0]: i for i in range(len(dataset))}
    written = 0
    for path in sorted(glob.glob(os.path.join(args.ckpt_root, "**", last), recursive=True)):
        name = os.path.basename(os.path.dirname(path))
        if name not in by_name:
            print(f"[{path}] no such frame in the dataset")
            continue
        views = dataset[by_name[name]]
        target = views[0]
        boxes = predictions.boxes_from_checkpoint(torch.load(path, map_location="cpu", weights_only=False))
        group = [dict(intrinsic_matrix=v["intrinsic_matrix"], extrinsic_matrix=v["extrinsic_matrix"], boxes_2d=v["boxes_2d"],
                      instance_ids=v["instance_ids"]) for v in views.values()]
        records = predictions.make_frame_predictions(boxes, target["extrinsic_matrix"], target["instance_ids"], group,
                                                     target["image"].shape[-2:])
        for relative_index, record in zip(views.keys(), records):
            filename = os.path.join(out_root, name, f"{relative_index:+d}.json")
            os.makedirs(os.path.dirname(filename), exist_ok=True)
            with open(filename, "w") as file:
                json.dump(record, file, indent=4, sort_keys=False)
            written += 1
    print(json.dumps(dict(prediction_files=written, out=out_root)))


if __name__ == "__main__":
    main()
