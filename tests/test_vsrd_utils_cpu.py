"""`vsrd.utils` pinned to the reference module's outputs (tests/golden/utils.npz, written by
tests/golden/make_golden_utils.py from the unmodified /root/reference/vsrd/utils.py) — ADVICE r1: reversed_pad,
ProgressMeter, collate_nested_dicts, linear_map and torch_function must behave as scripts/main.py expects."""
import importlib.util
import os

import numpy as np
import pytest
import torch

import vsrd

HERE = os.path.dirname(os.path.abspath(__file__))


def _cases():
    spec = importlib.util.spec_from_file_location("make_golden_utils", os.path.join(HERE, "golden", "make_golden_utils.py"))
    module = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(module)
    return module.cases(vsrd.utils)


def test_utils_match_the_reference_module_outputs():
    golden = np.load(os.path.join(HERE, "golden", "utils.npz"))
    ours = _cases()
    assert set(ours) == set(golden.files)
    for key in golden.files:
        want, got = golden[key], ours[key]
        assert want.shape == got.shape, key
        if want.dtype.kind in "US":
            assert want.tolist() == got.tolist(), (key, want.tolist(), got.tolist())
        else:
            assert want.dtype == got.dtype, (key, want.dtype, got.dtype)
            assert np.allclose(want, got, rtol=1e-6, atol=0.0, equal_nan=True), key


def test_reversed_pad_pads_the_first_dimension():
    """main.py:218-251 appends a -1 row so that index -1 ("instance not visible") selects it."""
    x = torch.arange(6.0).reshape(2, 3)
    y = vsrd.utils.reversed_pad(x, (0, 1))
    assert y.shape == (3, 3) and torch.equal(y[:2], x) and torch.equal(y[2], torch.zeros(3))
    assert torch.equal(y[torch.tensor([1, -1])], torch.stack([x[1], torch.zeros(3)]))


def test_progress_meter_replays_the_main_loop_calls():
    """main.py:94 constructs ProgressMeter(len(loader) * num_steps); :857-861, :1123 update three phases per step;
    :943-950 read progress / ETA / keys / means."""
    meter = vsrd.utils.ProgressMeter(2 * 50)
    for _ in range(50):
        meter.update(forward=0.1)
        meter.update(backward=0.2)
        meter.update(logging=0.05)
    assert meter.progress() == pytest.approx(0.5)
    assert dict(zip(meter.keys(), meter.means())) == pytest.approx(dict(forward=0.1, backward=0.2, logging=0.05))
    assert meter.arrival_seconds() == pytest.approx(0.35 * 100 * 0.5)
    assert meter.elapsed_seconds() == pytest.approx(0.35 * 100 * 0.5)
    import datetime, json
    datetime.timedelta(seconds=meter.arrival_seconds())
    json.dumps(dict(zip(meter.keys(), meter.means())))


def test_stop_watch_is_a_stack():
    watch = vsrd.utils.StopWatch()
    watch.start()
    assert watch.restart() >= 0.0 and len(watch.stack) == 1
    assert watch.stop() >= 0.0 and not watch.stack


def test_dict_and_default_dict():
    d = vsrd.utils.Dict.apply({"a": {"b": 1}})
    d.c = 3
    assert d.a.b == 1 and d["c"] == 3
    with pytest.raises(AttributeError):
        d.missing
    dd = vsrd.utils.DefaultDict(vsrd.utils.Dict)
    dd[0].update(x=1)
    assert dd[0].x == 1 and isinstance(dd[5], vsrd.utils.Dict)
    import json
    assert json.loads(json.dumps(d)) == {"a": {"b": 1}, "c": 3}


def test_import_module_evaluates_in_the_callers_scope():
    config = vsrd.utils.Dict.apply({
        "value": "eval:scale * 2",
        "layer": {"function": "torch.nn.Linear", "args": [4], "kwargs": {"out_features": "eval:scale"}},
        "items": [1, "eval:scale + 1"],
    })
    out = vsrd.utils.import_module(config, globals(), dict(scale=3))
    assert out.value == 6 and out["items"] == [1, 4]
    assert isinstance(out.layer, torch.nn.Linear) and out.layer.out_features == 3


def test_mode_switchers_restore():
    m = torch.nn.Linear(1, 1).eval()
    with vsrd.utils.TrainSwitcher(m):
        assert m.training
    assert not m.training
    with vsrd.utils.EvalSwitcher(m.train()):
        assert not m.training
    assert m.training
