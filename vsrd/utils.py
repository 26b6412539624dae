"""Pure-Python glue scripts/main.py needs from `vsrd.utils` (reference: vsrd/utils.py).  No kernels."""
import collections
import contextlib
import functools
import importlib
import logging
import os
import time

import numpy as np
import torch


def apply(function, element):
    """Map `function` over the leaves of nested dicts / lists / tuples."""
    if isinstance(element, dict):
        return function(type(element)((k, apply(function, v)) for k, v in element.items())) \
            if not isinstance(element, collections.defaultdict) else function(element)
    if isinstance(element, (list, tuple)):
        return function(type(element)(apply(function, v) for v in element))
    return function(element)


class Dict(dict):
    """dict with attribute access (vsrd/utils.py:16-47)."""

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
        return apply(lambda e: cls(e) if isinstance(e, dict) and not isinstance(e, cls) else e, dictionary)


class DefaultDict(collections.defaultdict):

    def __getattr__(self, key):
        if key.startswith("__"):
            raise AttributeError(key)
        return self[key]

    def __setattr__(self, key, value):
        self[key] = value

    def __delattr__(self, key):
        del self[key]


def compose(*functions):
    """compose(f, g)(x) == g(f(x)).  The chain is kept on the result (`__vsrd_compose__`) so the
    renderer can see through `compose(field, operator.itemgetter(0))` (main.py:1030)."""

    def composed(*args, **kwargs):
        out = functions[0](*args, **kwargs)
        for f in functions[1:]:
            out = f(out)
        return out

    composed.__vsrd_compose__ = functions
    return composed


def torch_function(function):
    """Run a numpy function on tensors: inputs -> numpy, outputs -> tensors (vsrd/utils.py)."""

    @functools.wraps(function)
    def wrapper(*args, **kwargs):
        to_np = lambda t: t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else t
        to_t = lambda a: torch.as_tensor(a) if isinstance(a, (np.ndarray, np.generic, float, int)) else a
        out = function(*map(to_np, args), **{k: to_np(v) for k, v in kwargs.items()})
        return tuple(map(to_t, out)) if isinstance(out, tuple) else to_t(out)

    return wrapper


def linear_map(inputs, in_min, in_max, out_min, out_max):
    return (inputs - in_min) / (in_max - in_min) * (out_max - out_min) + out_min


def reversed_pad(inputs, padding, *args, **kwargs):
    """F.pad with the per-dimension padding given first-dimension-first."""
    flat = [p for pair in reversed(list(padding)) for p in pair]
    return torch.nn.functional.pad(inputs, flat, *args, **kwargs)


def to(element, *args, **kwargs):
    return apply(lambda e: e.to(*args, **kwargs) if isinstance(e, torch.Tensor) else e, element)


def collate_nested_dicts(batch):
    first = batch[0]
    if isinstance(first, dict):
        return type(first)((k, collate_nested_dicts([b[k] for b in batch])) for k in first)
    return torch.utils.data.default_collate(batch)


def import_module(node, globals=None, locals=None):
    """Instantiate a config node: {"function": "pkg.fn", "args": [...], "kwargs": {...}} recursively;
    strings starting with "eval:" are evaluated in the caller's scope (vsrd/utils.py:318-340)."""
    if isinstance(node, str) and node.startswith("eval:"):
        return eval(node[len("eval:"):], globals, locals)
    if isinstance(node, dict) and "function" in node:
        module_name, _, attr = node["function"].rpartition(".")
        function = getattr(importlib.import_module(module_name), attr)
        args = [import_module(a, globals, locals) for a in node.get("args", [])]
        kwargs = {k: import_module(v, globals, locals) for k, v in node.get("kwargs", {}).items()}
        return function(*args, **kwargs)
    if isinstance(node, dict):
        return type(node)((k, import_module(v, globals, locals)) for k, v in node.items())
    if isinstance(node, (list, tuple)):
        return type(node)(import_module(v, globals, locals) for v in node)
    return node


class StopWatch:

    def __init__(self):
        self.start_time = time.time()

    def start(self):
        self.start_time = time.time()

    def restart(self):
        now = time.time()
        elapsed, self.start_time = now - self.start_time, now
        return elapsed

    def stop(self):
        return time.time() - self.start_time


class ProgressMeter:
    """Exponential moving averages of named scalars."""

    def __init__(self, momentum=0.9):
        self.momentum = momentum
        self.values = {}

    def update(self, **items):
        for k, v in items.items():
            self.values[k] = v if k not in self.values else self.momentum * self.values[k] + (1 - self.momentum) * v

    def __getattr__(self, key):
        try:
            return self.__dict__["values"][key]
        except KeyError:
            raise AttributeError(key)


class Saver:

    def __init__(self, dirname):
        self.dirname = dirname

    def save(self, filename, **states):
        os.makedirs(self.dirname, exist_ok=True)
        torch.save(states, os.path.join(self.dirname, filename))


class TrainSwitcher(contextlib.ContextDecorator):

    def __init__(self, *modules, mode=True):
        self.modules, self.mode = modules, mode

    def __enter__(self):
        self.previous = [m.training for m in self.modules]
        for m in self.modules:
            m.train(self.mode)
        return self

    def __exit__(self, *exc):
        for m, was in zip(self.modules, self.previous):
            m.train(was)
        return False


def get_logger(name, filename=None, level=logging.INFO):
    logger = logging.getLogger(name)
    logger.setLevel(level)
    logger.handlers.clear()
    handlers = [logging.StreamHandler()] + ([logging.FileHandler(filename)] if filename else [])
    for handler in handlers:
        handler.setFormatter(logging.Formatter("%(asctime)s: %(message)s"))
        logger.addHandler(handler)
    logger.propagate = False
    return logger
