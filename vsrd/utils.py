"""`vsrd.utils` — the host-side glue scripts/main.py and the tools call (reference: vsrd/utils.py).

Pure Python / PyTorch, no kernels.  Every helper keeps the reference's name, signature and observable behaviour
(cited per function) so that the unmodified script runs on top of it; tests/test_vsrd_utils_cpu.py pins them to
outputs of the reference module (tests/golden/utils.npz, made by tests/golden/make_golden_utils.py).
"""
import collections
import contextlib
import functools
import importlib
import itertools
import logging
import operator
import os
import time

import numpy as np
import torch


# ---- nested containers --------------------------------------------------------------------------------------------

def apply(function, element):
    """Post-order map over nested lists / tuples / sets / dicts: children first, then `function` on the rebuilt
    container itself (vsrd/utils.py:337-343)."""
    if isinstance(element, (list, tuple, set)):
        element = type(element)(apply(function, v) for v in element)
    if isinstance(element, dict):
        items = [(k, apply(function, v)) for k, v in element.items()]
        if isinstance(element, collections.defaultdict):
            rebuilt = type(element)(element.default_factory)
            rebuilt.update(items)
            element = rebuilt
        else:
            element = type(element)(items)
    return function(element)


class _AttributeAccess:
    """Attribute access onto the mapping's items; `getattr/setattr/delattr` reach the real attributes
    (vsrd/utils.py:16-47)."""

    def getattr(self, key):
        return object.__getattribute__(self, key)

    def setattr(self, key, value):
        object.__setattr__(self, key, value)

    def delattr(self, key):
        object.__delattr__(self, key)

    def __getattr__(self, key):
        try:
            return self[key]
        except KeyError:
            raise AttributeError(key)

    def __setattr__(self, key, value):
        self[key] = value

    def __delattr__(self, key):
        del self[key]

    def __getstate__(self):
        return self.__dict__

    def __setstate__(self, state):
        self.__dict__.update(state)

    @classmethod
    def apply(cls, dictionary):
        return apply(lambda e: cls(e) if isinstance(e, dict) else e, dictionary)


class Dict(_AttributeAccess, dict):
    pass


class DefaultDict(_AttributeAccess, collections.defaultdict):

    def __getattr__(self, key):
        if key.startswith("__") and key.endswith("__"):      # copy / pickle protocol probes must not create items
            raise AttributeError(key)
        return self[key]


# ---- meters (vsrd/utils.py:82-170) ----------------------------------------------------------------------------------

class _Meter(Dict):
    """name -> Dict(mean=..., count=...)."""

    def means(self):
        for stat in self.values():
            yield stat.mean

    def counts(self):
        for stat in self.values():
            yield stat.count


class StatMeter(_Meter):

    def update(self, **items):
        for key, value in items.items():
            stat = self.get(key, Dict(mean=value, variance=0, count=0))
            count = stat.count + 1
            mean = (stat.mean * stat.count + value) / count
            variance = ((stat.mean ** 2 + stat.variance) * stat.count + value ** 2) / count - mean ** 2
            self[key] = Dict(mean=mean, variance=variance, count=count)

    def variances(self):
        for stat in self.values():
            yield stat.variance


class SMAMeter(_Meter):

    def update(self, **items):
        for key, value in items.items():
            stat = self.get(key, Dict(mean=value, count=0))
            self[key] = Dict(mean=(stat.mean * stat.count + value) / (stat.count + 1), count=stat.count + 1)


class EMAMeter(_Meter):
    """Exponential moving average seeded with the first value (vsrd/utils.py:122-140)."""

    def __init__(self, *args, momentum=0.9, **kwargs):
        super().__init__(*args, **kwargs)
        self.setattr("momentum", momentum)

    def update(self, **items):
        momentum = self.getattr("momentum")
        for key, value in items.items():
            stat = self.get(key, Dict(mean=value, count=0))
            self[key] = Dict(mean=stat.mean * momentum + value * (1 - momentum), count=stat.count + 1)


class RuntimeMeter(EMAMeter):
    """Means rescaled to "per occurrence of the rarest key" (vsrd/utils.py:143-148)."""

    def means(self):
        least = min(self.counts())
        for stat in self.values():
            yield stat.mean * stat.count / least


class ProgressMeter(RuntimeMeter):
    """`ProgressMeter(num_steps)`: progress and ETA from the per-phase runtimes main.py feeds it
    (`update(forward=...)`, `update(backward=...)`, `update(logging=...)`; main.py:94, 857-861, 943-950, 1123)."""

    def __init__(self, num_steps, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.setattr("num_steps", num_steps)

    def progress(self):
        return min(self.counts()) / self.getattr("num_steps")

    def elapsed_seconds(self):
        return sum(self.means()) * self.getattr("num_steps") * self.progress()

    def arrival_seconds(self):
        return sum(self.means()) * self.getattr("num_steps") * (1.0 - self.progress())


class StopWatch:
    """Stack of start times: start() pushes, stop() pops and returns the elapsed seconds (vsrd/utils.py:173-187)."""

    def __init__(self):
        self.stack = []

    def start(self):
        self.stack.append(time.time())

    def stop(self):
        return time.time() - self.stack.pop()

    def restart(self):
        value = self.stop()
        self.start()
        return value


class Saver:

    def __init__(self, dirname):
        self.dirname = dirname

    def save(self, filename, **kwargs):
        os.makedirs(self.dirname, exist_ok=True)
        torch.save(kwargs, os.path.join(self.dirname, filename))


class ModeSwitcher(contextlib.ContextDecorator):
    """Sets `.train(mode)` on the models for the duration of the block (vsrd/utils.py:200-224)."""

    def __init__(self, mode, *models):
        self.mode, self.models, self.modes = mode, models, {}

    def __enter__(self):
        for model in self.models:
            self.modes[model] = model.training
            model.train(self.mode)

    def __exit__(self, exception_type, exception_value, traceback):
        for model in self.models:
            model.train(self.modes.pop(model))


class TrainSwitcher(ModeSwitcher):

    def __init__(self, *models):
        super().__init__(True, *models)


class EvalSwitcher(ModeSwitcher):

    def __init__(self, *models):
        super().__init__(False, *models)


class RandomStateRestorer(contextlib.ContextDecorator):

    def __init__(self, mode=True):
        self.mode = mode

    def __enter__(self):
        self.rng_state = torch.get_rng_state()

    def __exit__(self, exception_type, exception_value, traceback):
        if self.mode:
            torch.set_rng_state(self.rng_state)


# ---- config instantiation (vsrd/utils.py:311-334) -------------------------------------------------------------------

def import_function(name):
    module_name, function_name = name.rsplit(".", 1)
    return getattr(importlib.import_module(module_name), function_name)


def import_module(config, globals=None, locals=None):
    """{"function": "pkg.fn", "args": [...], "kwargs": {...}} nodes are called (children first); "eval:<expr>" strings
    are evaluated in the caller's scope; containers are rebuilt with their own type."""
    recurse = functools.partial(import_module, globals=globals, locals=locals)
    if isinstance(config, dict) and "function" in config:
        function = import_function(config["function"])
        args = [recurse(a) for a in config.get("args", [])]
        kwargs = {k: recurse(v) for k, v in config.get("kwargs", {}).items()}
        return function(*args, **kwargs)
    if isinstance(config, (list, tuple, set)):
        return type(config)(recurse(v) for v in config)
    if isinstance(config, dict):
        return type(config)((k, recurse(v)) for k, v in config.items())
    if isinstance(config, str) and config.split(":", 1)[0] == "eval":
        return eval(config.split(":", 1)[1], globals, locals)
    return config


# ---- functional helpers ---------------------------------------------------------------------------------------------

def cycle(iterable):
    while True:
        yield from iterable


def pairwise(iterable):
    prevs, nexts = itertools.tee(iterable)
    next(nexts, None)
    return zip(prevs, nexts)


def compose(function, *functions):
    """compose(f, g, h)(x) == h(g(f(x))) (vsrd/utils.py:368-372).  The chain stays readable on the result
    (`__vsrd_compose__`) so the renderer can see through `compose(field, operator.itemgetter(0))` (main.py:1030)."""
    chain = (function, *functions)

    def composed(*args, **kwargs):
        out = function(*args, **kwargs)
        for f in functions:
            out = f(out)
        return out

    composed.__vsrd_compose__ = chain
    return composed


def multimap(functions, *iterables):
    return map(lambda f, *x: f(*x), functions, *iterables)


def to(element, *args, **kwargs):
    return apply(lambda e: e.to(*args, **kwargs) if isinstance(e, torch.Tensor) else e, element)


def tensor_args(dtype=None, device=None):
    """Decorator: every positional / keyword argument goes through `torch.as_tensor(dtype, device)`
    (vsrd/utils.py:387-395)."""
    convert = functools.partial(torch.as_tensor, dtype=dtype, device=device)

    def decorator(function):
        def wrapper(*args, **kwargs):
            return function(*map(convert, args), **{k: convert(v) for k, v in kwargs.items()})
        return wrapper
    return decorator


def multi_dim(reducer):
    def multi_dim_reducer(inputs, dims, **kwargs):
        for dim in sorted(dims, reverse=True):
            inputs = reducer(inputs, dim=dim, **kwargs)
        return inputs
    return multi_dim_reducer


def unsqueeze(inputs, *dims):
    return functools.reduce(torch.unsqueeze, dims, inputs)


def reversed_pad(inputs, padding, *args, **kwargs):
    """`F.pad` with the (before, after) pairs listed FIRST dimension first and zero-extended to every dimension
    (vsrd/utils.py:426-430): `reversed_pad(x, (0, 1))` appends one row along dim 0 (main.py:218-251)."""
    padding = tuple(padding) + (0,) * (inputs.ndim * 2 - len(padding))
    pairs = [padding[i:i + 2] for i in range(0, len(padding), 2)]
    flat = tuple(p for pair in reversed(pairs) for p in pair)
    return torch.nn.functional.pad(inputs, flat, *args, **kwargs)


def _range_limit(reduce, inputs):
    first = compose(reduce, operator.itemgetter(0))
    return multi_dim(first)(inputs, dims=range(1, inputs.ndim), keepdim=True)


def linear_map(inputs, in_min=None, in_max=None, out_min=0.0, out_max=1.0):
    """Affine map of [in_min, in_max] onto [out_min, out_max]; limits default to the per-sample min / max and may be
    python scalars, ndarrays or tensors (coerced to `inputs`' dtype and device; vsrd/utils.py:433-440, main.py:815)."""
    in_min = _range_limit(torch.min, inputs) if in_min is None else in_min
    in_max = _range_limit(torch.max, inputs) if in_max is None else in_max

    @tensor_args(dtype=inputs.dtype, device=inputs.device)
    def mapped(inputs, in_min, in_max, out_min, out_max):
        return out_min + (out_max - out_min) * (inputs - in_min) / (in_max - in_min)

    return mapped(inputs, in_min, in_max, out_min, out_max)


def _torch_to_numpy(value):
    return value.detach().cpu().numpy() if isinstance(value, torch.Tensor) else value


def _numpy_to_torch(value):
    return torch.from_numpy(value) if isinstance(value, np.ndarray) else value


def torch_function(function):
    """Run a numpy function on tensors: nested tensor arguments -> ndarrays, nested ndarray results -> tensors
    (vsrd/utils.py:619-640; main.py:382-386 wraps scipy's linear_sum_assignment with it)."""

    def wrapper(*args, **kwargs):
        return apply(_numpy_to_torch, function(*apply(_torch_to_numpy, args), **apply(_torch_to_numpy, kwargs)))

    return wrapper


def numpy_function(function):
    """The converse: run a torch function on ndarrays (vsrd/utils.py:643-664)."""

    def wrapper(*args, **kwargs):
        return apply(_torch_to_numpy, function(*apply(_numpy_to_torch, args), **apply(_numpy_to_torch, kwargs)))

    return wrapper


def collate_nested_dicts(inputs):
    """Batch collation for nested sample dicts (vsrd/utils.py:667-690): lists of dicts become dicts of lists over the
    keys common to every sample, repeatedly, and a list of equally-shaped tensors is stacked; everything else (ragged
    tensor lists, strings, numbers) stays a list."""

    def merge(node):
        if isinstance(node, list) and node and all(isinstance(v, dict) for v in node):
            keys = functools.reduce(operator.and_, map(set, node))
            return {k: [v[k] for v in node] for k in node[0] if k in keys}
        return node

    while True:
        merged = apply(merge, inputs)
        if _same_structure(merged, inputs):
            break
        inputs = merged

    def stack(node):
        if isinstance(node, list) and node and all(isinstance(v, torch.Tensor) for v in node):
            if len({tuple(v.shape) for v in node}) == 1:
                return torch.stack(node, dim=0)
        return node

    return apply(stack, merged)


def _same_structure(a, b):
    """Container skeleton equality (the reference compares with `==`, which is ambiguous for multi-element tensors)."""
    if type(a) is not type(b):
        return False
    if isinstance(a, dict):
        return a.keys() == b.keys() and all(_same_structure(a[k], b[k]) for k in a)
    if isinstance(a, (list, tuple)):
        return len(a) == len(b) and all(_same_structure(x, y) for x, y in zip(a, b))
    return True


def get_logger(name, level=logging.INFO):
    """Named logger with one stream handler (vsrd/utils.py:693-703; repeated calls do not stack handlers here)."""
    logger = logging.getLogger(name)
    logger.setLevel(level)
    if not any(getattr(h, "_vsrd_stream", False) for h in logger.handlers):
        handler = logging.StreamHandler()
        handler.setLevel(level)
        handler.setFormatter(logging.Formatter("%(levelname)s: %(asctime)s: %(message)s"))
        handler._vsrd_stream = True
        logger.addHandler(handler)
    return logger
