"""TEST INFRASTRUCTURE ONLY — CPU oracle for the VSRD silhouette renderer hot path.

A plain-PyTorch (CPU, fp32 or fp64) restatement of what the reference computes on the path
named by BASELINE.json's north_star.  Every function cites the reference `file:line` it follows
(paths relative to the upstream repository root).  Gradients come from torch.autograd — including
the double backward through the spatial gradient — exactly as in the reference, so this module is
the yardstick for both values and parameter gradients of the CUDA kernels.

Parity status: pinned by `tests/golden/*.npz` (generated from the unmodified reference modules by
`tests/golden/make_golden.py`) and checked in `tests/test_oracle_golden.py`.

Differences from the reference are limited to *injection points* for randomness (stratified
jitter, importance-sampling uniforms) so both sides of a parity test consume identical draws
(SURVEY.md §5 "RNG / reproducibility").
"""
from __future__ import annotations

import dataclasses
import math
from typing import Callable, Optional, Sequence

import torch
import torch.nn.functional as F

# Residual MLP layout used by every shipped config
# (configs/kitti_360/vsrd/*/config.json:142-162): 48 -> 16 -> 16 -> 16 -> 16 -> 1.
MLP_IN = 48
MLP_HIDDEN = (16, 16, 16, 16)
NUM_FREQUENCIES = 8


def mlp_layer_sizes(in_channels: int = MLP_IN, hidden: Sequence[int] = MLP_HIDDEN):
    """(fan_in, fan_out) per layer and flat size per layer; hyper_distance_field.py:18-26."""
    fan_in = [in_channels, *hidden]
    fan_out = [*hidden, 1]
    sizes = [o * (i + 1) for i, o in zip(fan_in, fan_out)]
    return fan_in, fan_out, sizes


# --------------------------------------------------------------------------------------
# a1: ray generation — vsrd/rendering/utils.py:5-18
# --------------------------------------------------------------------------------------

def ray_casting(image_size, intrinsic_matrices, extrinsic_matrices):
    height, width = image_size
    vs, us = torch.meshgrid(torch.arange(height), torch.arange(width), indexing="ij")
    pixels = torch.stack([us, vs, torch.ones_like(us)], dim=-1)  # (u, v, 1)
    inv_k = torch.linalg.inv(intrinsic_matrices)
    inv_e = torch.linalg.inv(extrinsic_matrices)
    back = inv_e[..., :3, :3] @ inv_k
    directions = torch.einsum("...mn,hwn->...hwm", back, pixels.to(back))
    directions = F.normalize(directions, dim=-1)
    return inv_e[..., :3, 3], directions


# --------------------------------------------------------------------------------------
# a9 / a10: sample placement — vsrd/rendering/samplers.py:5-8 and :11-36
# --------------------------------------------------------------------------------------

def stratified_distances(bins, jitter=None):
    """`quadrature_sampler`; `jitter` replaces `torch.rand_like(bins[..., :-1])`."""
    if jitter is None:
        jitter = torch.rand_like(bins[..., :-1])
    return torch.lerp(bins[..., :-1], bins[..., 1:], jitter)


def importance_distances(bins, weights, num_samples, sorted_uniforms=None):
    """`inverse_transform_sampler`; `sorted_uniforms` replaces the sorted `torch.rand` draw."""
    pdf = F.normalize(weights, p=1, dim=-1)
    cdf = F.pad(torch.cumsum(pdf, dim=-1), (1, 0))
    if sorted_uniforms is None:
        sorted_uniforms = torch.rand(*cdf.shape[:-1], num_samples, device=cdf.device)
        sorted_uniforms = torch.sort(sorted_uniforms, dim=-1).values
    idx = torch.searchsorted(cdf, sorted_uniforms, right=False)
    idx = idx.clamp(min=1, max=cdf.shape[-1] - 1)
    cdf_lo, cdf_hi = cdf.gather(-1, idx - 1), cdf.gather(-1, idx)
    bin_lo, bin_hi = bins.gather(-1, idx - 1), bins.gather(-1, idx)
    frac = (sorted_uniforms - cdf_lo) / (cdf_hi - cdf_lo + 1e-6)
    return torch.lerp(bin_lo, bin_hi, frac)


# --------------------------------------------------------------------------------------
# a5 / a6 / a7 / a8: the per-instance field and its soft union
# --------------------------------------------------------------------------------------

def sinusoidal_encoding(x, num_frequencies: int = NUM_FREQUENCIES):
    """vsrd/models/encoders/sinusoidal_encoder.py:9-19; channel = coord*2F + k*2 + {cos,sin}."""
    freqs = (2.0 ** torch.arange(num_frequencies) * math.pi).to(x)
    arg = freqs * x.unsqueeze(-1)
    return torch.stack([torch.cos(arg), torch.sin(arg)], dim=-1).flatten(-3, -1)


def residual_mlp(flat_weights, features, in_channels: int = MLP_IN, hidden: Sequence[int] = MLP_HIDDEN):
    """`HyperDistanceField.distance_field`, vsrd/models/fields/hyper_distance_field.py:57-73.

    Each layer's block of `flat_weights` is `[fan_out][fan_in + 1]` row-major, bias in the last
    column; layers after the first are preceded by affine-free LayerNorm and exact (erf) GELU.
    """
    fan_in, fan_out, sizes = mlp_layer_sizes(in_channels, hidden)
    h = features
    for layer, (block, n_in, n_out) in enumerate(zip(torch.split(flat_weights, sizes, dim=-1), fan_in, fan_out)):
        if layer:
            h = F.gelu(F.layer_norm(h, [n_in]))
        mat = block.unflatten(-1, (n_out, n_in + 1))
        h = torch.einsum("...mn,...n->...m", mat, F.pad(h, (0, 1), value=1.0))
    return h


def box_sdf(p, half_extents):
    """vsrd/rendering/sdfs.py:5-19 (note the 1e-6 inside the square root)."""
    q = p.abs() - half_extents
    outside = torch.sqrt(torch.sum(F.relu(q) ** 2.0, dim=-1, keepdim=True) + 1e-6)
    inside = F.relu(-torch.max(q, dim=-1, keepdim=True).values)
    return outside - inside


def instance_sdf(x, location, rotation, half_extents, flat_weights, scale, num_frequencies=NUM_FREQUENCIES):
    """One instance: translation -> rotation -> box (+ residual).

    sdfs.py:22-37 (`positions - t`, then `positions @ R`), scripts/main.py:433-458
    (residual = sigmoid(MLP(PE((|p_x|, p_y, p_z) / scale)) - 1), summed with the box SDF).
    `flat_weights=None` is the warm-up branch (scripts/main.py:582-618).
    """
    p = (x - location) @ rotation
    d = box_sdf(p, half_extents)
    if flat_weights is not None:
        px, py, pz = torch.unbind(p, dim=-1)
        folded = torch.stack([px.abs(), py, pz], dim=-1) / scale
        out = residual_mlp(flat_weights, sinusoidal_encoding(folded, num_frequencies))
        d = d + torch.sigmoid(out - 1.0)
    return d


@dataclasses.dataclass
class Scene:
    """Decoded per-frame parameters (what scripts/main.py:530-578 closes over)."""
    locations: torch.Tensor          # [N, 3]
    rotations: torch.Tensor          # [N, 3, 3]
    half_extents: torch.Tensor       # [N, 3]
    mlp_weights: Optional[torch.Tensor]  # [N, 1617] or None during warm-up
    temperature: float               # sdf_union_temperature
    scale: float = 100.0             # max(distance_range), scripts/main.py:441
    num_frequencies: int = NUM_FREQUENCIES

    @property
    def num_instances(self):
        return self.locations.shape[0]

    def field(self) -> Callable:
        """The `soft_union` closure of scripts/main.py:477-492 over `instance_field`s (:460-475)."""

        def union(x):
            per_instance = [
                instance_sdf(
                    x, self.locations[i], self.rotations[i], self.half_extents[i],
                    None if self.mlp_weights is None else self.mlp_weights[i],
                    self.scale, self.num_frequencies,
                )
                for i in range(self.num_instances)
            ]
            d = torch.stack(per_instance, dim=0)                       # [N, ..., 1]
            w = F.softmin(d / self.temperature, dim=0)
            sdf = torch.sum(d * w, dim=0)
            # one-hot labels blended by the same weights == the weights themselves (main.py:470-488)
            labels = w.squeeze(-1).movedim(0, -1)
            return sdf, labels

        return union


# --------------------------------------------------------------------------------------
# a11: the renderer — vsrd/rendering/renderers.py:177-270
# --------------------------------------------------------------------------------------

def hierarchical_volumetric_rendering(
    distance_field,
    ray_positions,
    ray_directions,
    distance_range,
    num_samples,
    sdf_std_deviation,
    cosine_ratio=1.0,
    epsilon=1e-6,
    sampled_distances=None,
    sampled_weights=None,
    *,
    jitter=None,
    sorted_uniforms=None,
):
    """Same contract as the reference; `jitter` ([..., 1, S]) / `sorted_uniforms` ([..., 1, S])
    inject the random draws of pass 1 / pass 2."""
    if sampled_distances is None:
        bins = torch.linspace(*distance_range, num_samples + 1, device=ray_directions.device)
        bins = bins.expand(*ray_directions.shape[:-1], 1, -1)                 # renderers.py:191-192
        dist = stratified_distances(bins, jitter)
    else:
        coarse = sampled_distances.permute(*range(1, sampled_distances.ndim), 0)
        weights = sampled_weights.permute(*range(1, sampled_weights.ndim), 0)
        fine = importance_distances(coarse, weights, num_samples, sorted_uniforms)
        dist = torch.sort(torch.cat([coarse, fine], dim=-1), dim=-1).values    # renderers.py:201-210
    dist = dist.permute(-1, *range(dist.ndim - 1))                              # sample-major
    return render_pass(distance_field, ray_positions, ray_directions, dist,
                       sdf_std_deviation, cosine_ratio, epsilon)


def render_pass(distance_field, ray_positions, ray_directions, dist, sdf_std_deviation,
                cosine_ratio=1.0, epsilon=1e-6):
    """The part of the renderer after sample placement (renderers.py:212-270) for sample-major
    distances `dist` [M+1, ..., 1].  Exposed separately so parity tests can inject identical
    sample positions into the oracle and the CUDA path (SURVEY.md §7 hard part 3)."""
    intervals = dist[1:] - dist[:-1]
    midpoints = (dist[:-1] + dist[1:]) / 2.0
    positions = ray_positions + ray_directions * midpoints                      # renderers.py:216

    create_graph = torch.is_grad_enabled()
    with torch.enable_grad():
        positions.requires_grad_(True)
        sdf, *features = distance_field(positions)
        grads, = torch.autograd.grad(sdf, positions, torch.ones_like(sdf), create_graph=create_graph)
        normals = F.normalize(grads, dim=-1)

    cosines = torch.sum(ray_directions * normals, dim=-1, keepdim=True)
    cosines = -torch.lerp(F.relu(-cosines * 0.5 + 0.5), F.relu(-cosines), cosine_ratio)  # :231-236
    sdf_prev = sdf - cosines * intervals / 2.0
    sdf_next = sdf + cosines * intervals / 2.0
    cdf_prev = torch.sigmoid(sdf_prev / sdf_std_deviation)
    cdf_next = torch.sigmoid(sdf_next / sdf_std_deviation)
    alpha = F.relu((cdf_prev - cdf_next) / (cdf_prev + epsilon))                # :248
    trans = torch.cumprod(1.0 - alpha, dim=0)
    trans = torch.cat([torch.ones_like(trans[:1]), trans[:-1]], dim=0)          # exclusive, :250-256
    weights = trans * alpha
    accumulated = [torch.sum(f * weights, dim=0) for f in features]
    return (*accumulated, grads, dist, weights)


def two_pass_render(distance_field, ray_positions, ray_directions, distance_range, num_samples,
                    sdf_std_deviation, cosine_ratio, *, jitter=None, sorted_uniforms=None):
    """`hierarchical_wrapper`, scripts/main.py:511-523: no-grad coarse pass, then the fine pass.

    Returns (labels, sampled_gradients, coarse_distances, coarse_weights, fine_distances, fine_weights).
    """
    with torch.no_grad():
        *_, coarse_d, coarse_w = hierarchical_volumetric_rendering(
            distance_field, ray_positions, ray_directions, distance_range, num_samples,
            sdf_std_deviation, cosine_ratio, jitter=jitter)
    labels, grads, fine_d, fine_w = hierarchical_volumetric_rendering(
        distance_field, ray_positions, ray_directions, distance_range, num_samples,
        sdf_std_deviation, cosine_ratio, sampled_distances=coarse_d, sampled_weights=coarse_w,
        sorted_uniforms=sorted_uniforms)
    return labels, grads, coarse_d, coarse_w, fine_d, fine_w


# --------------------------------------------------------------------------------------
# a12 / a13: losses, and the annealing schedule
# --------------------------------------------------------------------------------------

def silhouette_loss(labels, targets, pd_indices=None, gt_indices=None):
    """scripts/main.py:653-671: mean BCE on clamped labels (matched instance order)."""
    if pd_indices is not None:
        labels = labels[..., pd_indices]
        targets = targets[..., gt_indices]
    return F.binary_cross_entropy(labels.clamp(1.0e-6, 1.0 - 1.0e-6), targets, reduction="none").mean()


def eikonal_loss(sampled_gradients):
    """scripts/main.py:679-687."""
    norms = torch.norm(sampled_gradients, dim=-1)
    return F.mse_loss(norms, torch.ones_like(norms), reduction="mean")


def cosine_annealing(x, hi, lo):
    """scripts/main.py:420."""
    return (math.cos(math.pi * x) + 1.0) / 2.0 * (hi - lo) + lo


def render_loss(scene: Scene, ray_positions, ray_directions, targets, *, num_samples, distance_range,
                sdf_std_deviation, cosine_ratio, eikonal_weight=0.01, jitter=None, sorted_uniforms=None):
    """One renderer step as scripts/main.py:629-687 + 855 runs it (silhouette + weighted eikonal).

    The eikonal term is only present once the residual field is on (main.py:677).
    Returns (loss, dict of intermediates).
    """
    labels, grads, cd, cw, fd, fw = two_pass_render(
        scene.field(), ray_positions, ray_directions, distance_range, num_samples,
        sdf_std_deviation, cosine_ratio, jitter=jitter, sorted_uniforms=sorted_uniforms)
    loss = silhouette_loss(labels, targets)
    parts = dict(silhouette_loss=loss)
    if scene.mlp_weights is not None:
        eik = eikonal_loss(grads)
        parts.update(eikonal_loss=eik)
        loss = loss + eikonal_weight * eik
    parts.update(labels=labels, sampled_gradients=grads, coarse_distances=cd, coarse_weights=cw,
                 fine_distances=fd, fine_weights=fw)
    return loss, parts


# --------------------------------------------------------------------------------------
# a3 / a4: parameter decoding (BoxParameters3D) and the hypernetwork (HyperDistanceField)
# --------------------------------------------------------------------------------------

DEFAULT_LOCATION_RANGE = [[-50.0, 1.55 - 1.75 / 2.0 - 5.0, 0.0], [50.0, 1.55 - 1.75 / 2.0 + 5.0, 100.0]]
DEFAULT_DIMENSION_RANGE = [[0.75, 0.75, 1.5], [1.00, 1.00, 2.5]]


def rotation_matrix_y(cos, sin):
    """vsrd/models/detectors/box_parameters.py:5-13."""
    one, zero = torch.ones_like(cos), torch.zeros_like(cos)
    return torch.stack([
        torch.stack([cos, zero, sin], dim=-1),
        torch.stack([zero, one, zero], dim=-1),
        torch.stack([-sin, zero, cos], dim=-1),
    ], dim=-2)


def decode_box_parameters(raw_locations, raw_dimensions, raw_orientations,
                          location_range=None, dimension_range=None):
    """box_parameters.py:60-71: sigmoid-lerp into the ranges; yaw from a normalised (cos, sin)."""
    lr = torch.as_tensor(DEFAULT_LOCATION_RANGE if location_range is None else location_range).to(raw_locations)
    dr = torch.as_tensor(DEFAULT_DIMENSION_RANGE if dimension_range is None else dimension_range).to(raw_dimensions)
    locations = torch.lerp(lr[0], lr[1], torch.sigmoid(raw_locations))
    half_extents = torch.lerp(dr[0], dr[1], torch.sigmoid(raw_dimensions))
    unit = F.normalize(raw_orientations, dim=-1)
    rotations = rotation_matrix_y(*torch.unbind(unit, dim=-1))
    return locations, half_extents, rotations


_CORNER_SIGNS = [
    [-1.0, -1.0, +1.0], [+1.0, -1.0, +1.0], [+1.0, -1.0, -1.0], [-1.0, -1.0, -1.0],
    [-1.0, +1.0, +1.0], [+1.0, +1.0, +1.0], [+1.0, +1.0, -1.0], [-1.0, +1.0, -1.0],
]


def box_corners(locations, half_extents, rotations):
    """`decode_box_3d`, box_parameters.py:73-91 (KITTI-360 evaluation corner order)."""
    corners = half_extents.new_tensor(_CORNER_SIGNS) * half_extents.unsqueeze(-2)
    return corners @ rotations.transpose(-2, -1) + locations.unsqueeze(-2)


class HyperNetwork(torch.nn.Module):
    """The hypernetwork half of `HyperDistanceField` (hyper_distance_field.py:30-55, 75-77):
    weight-normed Linear -> LayerNorm -> GELU blocks, then a weight-normed Linear to 1617."""

    def __init__(self, in_channels=MLP_IN, hidden=MLP_HIDDEN, hyper_in=256, hyper_hidden=(256, 256, 256, 256)):
        super().__init__()
        _, _, sizes = mlp_layer_sizes(in_channels, hidden)
        widths = [hyper_in, *hyper_hidden]
        blocks = []
        for a, b in zip(widths[:-1], widths[1:]):
            blocks.append(torch.nn.Sequential(torch.nn.Linear(a, b), torch.nn.LayerNorm(b), torch.nn.GELU()))
        blocks.append(torch.nn.Sequential(torch.nn.Linear(widths[-1], sum(sizes))))
        self.hypernetwork = torch.nn.Sequential(*blocks)
        self.apply(lambda m: torch.nn.utils.weight_norm(m) if isinstance(m, torch.nn.Linear) else m)

    def forward(self, embeddings):
        return self.hypernetwork(embeddings)
