"""The experimental tcgen05 / TMEM field backward (vsrd_field_bwd_umma.cu) against the shipped kernel.

It is NOT on the product path (slower, and its MLP weight gradients carry bf16 staging error); this test keeps the
measurement repeatable and the record honest: pose gradients and the layer-4 weight gradient (fp32 paths) agree to
1e-4, the tensor-core weight gradients to the bf16 level."""
import pytest
import torch

from tests import fullsize_cases as fc
from vsrd_b200 import ops

pytestmark = pytest.mark.gpu


def _rel(x, y):
    return float((x - y).norm() / x.norm().clamp_min(1e-30))


@pytest.mark.parametrize("rays,sparse", [(96, 0.0), (300, 0.5)])
def test_tcgen05_backward_agrees_with_shipped_kernel(rays, sparse):
    dev = torch.device("cuda", 0)
    inp = fc.scene_inputs("cfg2")
    scene = ops.SceneArgs(*[inp[k].to(dev) for k in fc.GRAD_NAMES], fc.SCHEDULES["mid"]["temperature"], 100.0)
    gen = torch.Generator().manual_seed(3)
    dist = torch.sort(torch.rand(rays, 2 * fc.NUM_SAMPLES, generator=gen) * 60.0, dim=-1).values.to(dev)
    args = ops.RayArgs(inp["origins"][:rays].to(dev), inp["directions"][:rays].to(dev), dist)
    field = ops.field_forward(scene, args, cull=False)
    adj = torch.randn(field.shape, generator=gen) * 1e-3
    if sparse:
        keep = torch.rand(rays, generator=gen) >= sparse
        adj = (adj.reshape(adj.shape[0], rays, -1, 4) * keep[None, :, None, None]).reshape(adj.shape)
    adj = adj.to(dev)
    shipped = ops.field_backward(scene, args, adj)
    trial = ops.experimental_field_backward_tcgen05(scene, args, adj)
    for name, x, y in zip(("locations", "rotations", "half_extents"), shipped, trial):
        assert torch.isfinite(y).all(), name
        assert _rel(x, y) < 1e-4, (name, _rel(x, y))
    w, u = shipped[3], trial[3]
    assert torch.isfinite(u).all()
    assert _rel(w[:, 1600:], u[:, 1600:]) < 1e-4            # layer 4: SIMT fp32
    for lo, hi in ((0, 784), (784, 1056), (1056, 1328), (1328, 1600)):
        assert _rel(w[:, lo:hi], u[:, lo:hi]) < 8e-3, (lo, _rel(w[:, lo:hi], u[:, lo:hi]))   # bf16 operand staging
