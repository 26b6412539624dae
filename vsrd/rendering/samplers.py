"""`vsrd.rendering.samplers` (reference: vsrd/rendering/samplers.py) on the placement kernels.

Shapes follow the reference: `bins` [..., S+1] / [..., S], outputs [..., S]."""
import torch

from vsrd_b200 import ops


def _seed():
    # consume torch's CPU generator so `torch.manual_seed` keeps runs reproducible; no device sync
    return int(torch.randint(0, 2 ** 62, (1,)).item())


def quadrature_sampler(bins, deterministic=False):
    """Stratified samples, one per bin.  All rays must share the bin edges (as in the renderer)."""
    lead = bins.shape[:-1]
    flat = bins.reshape(-1, bins.shape[-1])
    num_rays = flat.shape[0]
    if num_rays > 1 and flat.stride(0) != 0 and not bool((flat == flat[:1]).all()):
        raise RuntimeError("vsrd_b200: quadrature_sampler needs the same bin edges on every ray (the renderer's "
                           "`linspace(*distance_range).expand(...)`, renderers.py:191-192)")
    jitter = torch.full((num_rays, flat.shape[1] - 1), 0.5, device=bins.device) if deterministic else None
    out = ops.place_coarse(flat[0].float(), num_rays, jitter, _seed())
    return out.reshape(*lead, -1).to(bins.dtype)


def inverse_transform_sampler(bins, weights, num_samples, deterministic=False):
    """Importance samples from the piecewise-constant pdf `weights` over `bins`.
    Returns the new samples only (the renderer merges them with `bins`)."""
    if num_samples != bins.shape[-1]:
        raise RuntimeError("vsrd_b200: inverse_transform_sampler draws exactly len(bins) samples per ray "
                           f"(got num_samples={num_samples}, bins={bins.shape[-1]})")
    lead = bins.shape[:-1]
    t = bins.reshape(-1, bins.shape[-1]).float()
    w = weights.reshape(-1, weights.shape[-1]).float()
    uniforms = None
    if deterministic:
        uniforms = torch.linspace(0.0, 1.0, num_samples, device=bins.device).expand(t.shape[0], -1).contiguous()
    merged = ops.place_fine(t, w, uniforms, _seed())
    # the kernel returns the merged, sorted list; recover the new samples as the multiset difference
    is_new = torch.ones_like(merged, dtype=torch.bool)
    pos = torch.searchsorted(merged, t.contiguous())
    is_new.scatter_(1, pos, False)
    return merged[is_new].reshape(*lead, num_samples).to(bins.dtype)
