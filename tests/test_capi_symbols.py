"""CPU-side checks of the C-ABI boundary: the library builds/loads without a GPU and exports every
symbol include/vsrd_b200.h declares; the ctypes structs match the header's layout."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "vsrd_b200.h")


def _declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vsrd_[a-z_0-9]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    from vsrd_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        from vsrd_b200 import build
        build.build()
    return _lib.load()


def test_header_declares_expected_entry_points():
    names = _declared_functions()
    assert "vsrd_field_forward" in names and "vsrd_field_backward" in names
    assert "vsrd_composite_forward" in names and "vsrd_composite_backward" in names
    assert len(names) >= 19
    assert {'vsrd_projection_step', 'vsrd_select_rays', 'vsrd_ray_cdf_build', 'vsrd_step_state_update',
            'vsrd_hyper_forward', 'vsrd_hyper_backward', 'vsrd_decode_boxes', 'vsrd_adam_step'} <= set(names)


def test_library_exports_every_declared_symbol(lib):
    from vsrd_b200 import _lib
    for name in _declared_functions():
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature"
    assert sorted(_lib.SIGNATURES) == _declared_functions()


def test_version_and_error_string_callable_without_gpu(lib):
    assert lib.vsrd_version() == 1
    assert isinstance(lib.vsrd_last_error(), bytes)


def test_struct_layout_matches_header():
    from vsrd_b200 import _lib
    # VsrdScene: 2 x int32, 4 pointers, 2 floats, 1 pointer -> 56 bytes on LP64; VsrdRays: 2 x int32 + 3 pointers
    assert ctypes.sizeof(_lib.VsrdScene) == 56
    assert ctypes.sizeof(_lib.VsrdStepState) == 32
    assert ctypes.sizeof(_lib.VsrdSchedule) == 48
    assert ctypes.sizeof(_lib.VsrdViews) == 32
    assert ctypes.sizeof(_lib.VsrdRays) == 56
    assert ctypes.sizeof(_lib.VsrdRenderParams) == 16
    assert ctypes.sizeof(_lib.VsrdLoss) == 16
    # model entry points (sizes printed by a C program including the header: gcc, LP64)
    assert ctypes.sizeof(_lib.VsrdHyperLayer) == 48 and ctypes.sizeof(_lib.VsrdHyperNet) == 248
    assert ctypes.sizeof(_lib.VsrdHyperLayerGrads) == 40 and ctypes.sizeof(_lib.VsrdHyperNetGrads) == 208
    assert ctypes.sizeof(_lib.VsrdBoxRanges) == 48 and ctypes.sizeof(_lib.VsrdAdamGroups) == 192
    text = open(HEADER).read()
    assert f"#define VSRD_MLP_WEIGHTS {_lib.MLP_WEIGHTS}" in text
    assert f"#define VSRD_GRAD_STRIDE {_lib.GRAD_STRIDE}" in text
    assert f"#define VSRD_MAX_INSTANCES {_lib.MAX_INSTANCES}" in text
    assert f"#define VSRD_MAX_INTERVALS {_lib.MAX_INTERVALS}" in text
    assert f"#define VSRD_HYPER_WIDTH {_lib.HYPER_WIDTH}" in text
    assert f"#define VSRD_MAX_PARAM_GROUPS {_lib.MAX_PARAM_GROUPS}" in text


def test_argument_errors_are_reported_without_gpu(lib):
    """Validation happens before any CUDA call, so the error convention is testable on CPU."""
    from vsrd_b200 import _lib
    scene = _lib.VsrdScene(0, 0, None, None, None, None, 1.0, 100.0, None)
    rays = _lib.VsrdRays(1, 1, None, None, None, None, None, None)
    status = lib.vsrd_field_forward(ctypes.byref(scene), ctypes.byref(rays), None, None)
    assert status != 0
    assert b"num_instances" in lib.vsrd_last_error()
    with pytest.raises(RuntimeError, match="num_instances"):
        _lib.check(status)


def test_ops_refuse_cpu_tensors():
    import torch
    from vsrd_b200 import ops
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        ops.SceneArgs(torch.zeros(2, 3), torch.zeros(2, 3, 3), torch.zeros(2, 3), None, 1.0)
