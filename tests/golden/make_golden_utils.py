"""Pin `vsrd.utils` to the UNMODIFIED reference module (build container only; needs /root/reference):

    python tests/golden/make_golden_utils.py   ->  tests/golden/utils.npz

Each entry is the output of the reference's own `vsrd/utils.py` on seeded inputs; tests/test_vsrd_utils_cpu.py
replays the same calls through the drop-in module.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import ref_import  # noqa: E402


def cases(utils):
    """name -> ndarray for every pinned call (shared by the maker and the test)."""
    gen = torch.Generator().manual_seed(0)
    out = {}
    for ndim in (1, 2, 3, 4):
        x = torch.rand(*range(2, 2 + ndim), generator=gen)
        out[f"reversed_pad_01_{ndim}d"] = utils.reversed_pad(x, (0, 1))
        if ndim >= 2:
            out[f"reversed_pad_1203_{ndim}d"] = utils.reversed_pad(x, (1, 2, 0, 3), mode="constant", value=-1.0)
    ids = torch.arange(5)
    out["reversed_pad_int"] = utils.reversed_pad(ids, (0, 1))
    x = torch.rand(2, 5, 7, 2, generator=gen) * 300.0
    out["linear_map_ndarray_limit"] = utils.linear_map(x, 0, np.subtract((376, 1408)[::-1], 1), -1.0, 1.0)
    out["linear_map_auto"] = utils.linear_map(x)
    out["linear_map_scalars"] = utils.linear_map(x, 1.0, 7.0, 0.0, 10.0)
    # meters: the call sequence of scripts/main.py:94, 857-861, 1123 with fixed runtimes
    meter = utils.ProgressMeter(30)
    trace = []
    for step in range(7):
        meter.update(forward=0.10 + 0.01 * step)
        meter.update(backward=0.20 - 0.01 * step)
        if step % 2 == 0:
            trace.append([meter.progress(), meter.arrival_seconds(), meter.elapsed_seconds(), *meter.means()] + [0.0] * (3 - len(meter)))
        meter.update(logging=0.05 * (step + 1))
        trace.append([meter.progress(), meter.arrival_seconds(), meter.elapsed_seconds(), *meter.means()])
    out["progress_meter_trace"] = np.asarray(trace, dtype=np.float64)
    out["progress_meter_keys"] = np.asarray(list(meter.keys()))
    stat = utils.StatMeter()
    for v in (1.0, 4.0, 2.5):
        stat.update(a=v, b=2 * v)
    out["stat_meter"] = np.asarray([list(stat.means()), list(stat.variances()), list(stat.counts())], dtype=np.float64)
    # torch_function: nested outputs (tuple of ndarrays) and nested inputs
    fn = utils.torch_function(lambda a, pair: (a + 1, [pair[0] * 2, pair[1].sum()]))
    r = fn(torch.arange(3.0), (torch.ones(2), torch.arange(4.0)))
    out["torch_function_0"] = r[0]
    out["torch_function_1"] = r[1][0]
    out["torch_function_types"] = np.asarray([type(r).__name__, type(r[0]).__name__, type(r[1]).__name__, type(r[1][0]).__name__, type(r[1][1]).__name__])
    # collate: ragged + equal tensors, strings, missing keys, nesting two deep
    samples = [
        {0: dict(image=torch.full((3, 2, 2), 1.0), masks=torch.ones(2, 2, 2), filename="a.png", extra=1),
         1: dict(image=torch.full((3, 2, 2), 2.0), masks=torch.ones(1, 2, 2), filename="b.png")},
        {0: dict(image=torch.full((3, 2, 2), 3.0), masks=torch.ones(3, 2, 2), filename="c.png"),
         1: dict(image=torch.full((3, 2, 2), 4.0), masks=torch.ones(1, 2, 2), filename="d.png")},
    ]
    c = utils.collate_nested_dicts(samples)
    out["collate_keys"] = np.asarray([repr(sorted(c.keys())), repr(sorted(c[0].keys())), repr(sorted(c[1].keys()))])
    out["collate_image0"] = c[0]["image"]
    out["collate_masks1"] = c[1]["masks"]
    out["collate_types"] = np.asarray([type(c[0]["image"]).__name__, type(c[0]["masks"]).__name__, type(c[1]["masks"]).__name__,
                                       type(c[0]["filename"]).__name__, repr(c[0]["filename"]), repr([tuple(m.shape) for m in c[0]["masks"]])])
    single = utils.collate_nested_dicts(samples[:1])
    out["collate_single_types"] = np.asarray([type(single[0]["image"]).__name__, repr(tuple(single[0]["image"].shape)),
                                              type(single[0]["masks"]).__name__, repr(tuple(single[0]["masks"].shape)), repr(single[0]["filename"])])
    # compose / apply / Dict
    out["compose"] = np.asarray(utils.compose(lambda a, b: a + b, lambda v: v * 3, lambda v: v - 1)(2, 5))
    d = utils.Dict.apply({"a": {"b": [{"c": 1}], "t": (1, {"z": 2})}})
    out["dict_apply_types"] = np.asarray([type(d).__name__, type(d.a).__name__, type(d.a.b[0]).__name__, type(d.a.t).__name__, type(d.a.t[1]).__name__])
    return {k: (v.numpy() if isinstance(v, torch.Tensor) else np.asarray(v)) for k, v in out.items()}


if __name__ == "__main__":
    with ref_import.reference_modules() as ref:
        data = cases(ref.utils)
    np.savez(os.path.join(HERE, "utils.npz"), **data)
    print(f"wrote {len(data)} entries")
