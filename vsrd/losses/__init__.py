"""`vsrd.losses` — the loss surface of the reference package (vsrd/losses/*.py) plus the three losses of the per-frame
optimisation step, which the reference writes inline in scripts/main.py.

    reduced                              the `reduction=` decorator every reference loss carries (losses/utils.py:4-15)
    cross_entropy ... focal_tversky_loss classification losses on probabilities (losses/classification_losses.py)
    rotation/translation_consistency_loss, sampson_epipolar_distance   (losses/geometric_losses.py)
    ssim_loss, photometric_loss          (losses/photometric_losses.py)
    gradient_x/y, smoothness_loss, motion_smoothness_loss, motion_sparsity_loss   (losses/smoothness_losses.py)
    silhouette_loss, eikonal_loss        main.py:653-671, 679-687 on the renderer's outputs (any device, autograd)
    projection_losses                    main.py:339-415 for all views at once; on CUDA one fused launch
                                         (projection + Hungarian matching + DIoU / smooth-L1 and their adjoint)
    fused_silhouette_eikonal_loss        fine pass with both reductions fused into the compositing kernels (CUDA)

The probabilistic NLL / energy-score zoo (losses/probabilistic_losses.py) belongs to the DETR-style detectors, which
are out of scope, and is not provided.  tests/test_vsrd_losses_cpu.py pins the shared functions to the
reference module's outputs (tests/golden/losses.npz).
"""
import functools

import torch
import torch.nn as nn


def reduced(loss_function):
    @functools.wraps(loss_function)
    def wrapper(*args, reduction="mean", **kwargs):
        losses = loss_function(*args, **kwargs)
        if reduction == "none":
            return losses
        if reduction == "mean":
            return torch.mean(losses)
        if reduction == "sum":
            return torch.sum(losses)
        raise ValueError(f"`reduction` argument should be 'none'|'mean'|'sum', but got {reduction}.")
    return wrapper


# ---- classification losses on probabilities --------------------------------------------------------------------------

def _clamp_probability(p, epsilon):
    return torch.clamp(p, epsilon, 1.0 - epsilon)


def _two_sided(loss):
    """binary variant: loss(p, t) + loss(1 - p, 1 - t)."""
    @reduced
    def binary(inputs, targets, epsilon=1e-6):
        return (loss(inputs, targets, epsilon=epsilon, reduction="none")
                + loss(1.0 - inputs, 1.0 - targets, epsilon=epsilon, reduction="none"))
    return binary


@reduced
def cross_entropy(inputs, targets, dim=None, keepdim=False, epsilon=1e-6):
    losses = -targets * torch.log(_clamp_probability(inputs, epsilon))
    return torch.sum(losses, dim=dim, keepdim=keepdim) if dim else losses


@reduced
def kl_divergence(inputs, targets, dim=None, keepdim=False, epsilon=1e-6):
    inputs, targets = _clamp_probability(inputs, epsilon), _clamp_probability(targets, epsilon)
    losses = -targets * (torch.log(inputs) - torch.log(targets))
    return torch.sum(losses, dim=dim, keepdim=keepdim) if dim else losses


@reduced
def js_divergence(inputs, targets, dim=None, keepdim=False, epsilon=1e-6):
    means = inputs * 0.5 + targets * 0.5
    kw = dict(dim=dim, keepdim=keepdim, epsilon=epsilon, reduction="none")
    return kl_divergence(means, inputs, **kw) * 0.5 + kl_divergence(means, targets, **kw) * 0.5


binary_cross_entropy = _two_sided(cross_entropy)
binary_kl_divergence = _two_sided(kl_divergence)
binary_js_divergence = _two_sided(js_divergence)


@reduced
def focal_loss(inputs, targets, alpha=0.25, gamma=2.0):
    """alpha_t (1 - p_t)^gamma (-log p_t), arXiv:1708.02002."""
    return ((1.0 - torch.abs(targets - alpha)) * torch.abs(targets - inputs) ** gamma
            * binary_cross_entropy(inputs, targets, reduction="none"))


@reduced
def quality_focal_loss(inputs, targets, beta=2.0):
    """|y - sigma|^beta BCE(sigma, y), arXiv:2006.04388."""
    return torch.abs(targets - inputs) ** beta * binary_cross_entropy(inputs, targets, reduction="none")


@reduced
def tversky_loss(inputs, targets, alpha=0.7, beta=0.3, epsilon=1.0):
    tp = torch.sum(inputs * targets, dim=(-2, -1))
    fn = torch.sum((1.0 - inputs) * targets, dim=(-2, -1))
    fp = torch.sum(inputs * (1.0 - targets), dim=(-2, -1))
    return 1.0 - (tp + epsilon) / (tp + alpha * fn + beta * fp + epsilon)


@reduced
def focal_tversky_loss(inputs, targets, gamma=0.75, **kwargs):
    return tversky_loss(inputs, targets, **kwargs, reduction="none") ** gamma


# ---- geometric losses ------------------------------------------------------------------------------------------------

def _cycle_consistency(select, reference):
    @reduced
    def loss(source_extrinsic_matrices, target_extrinsic_matrices, epsilon=1e-6):
        def deviation(matrices):
            part = select(matrices)
            error = (part - reference(part)) ** 2
            return error.mean(dim=tuple(range(-reference(part).dim(), 0)))
        cycle = target_extrinsic_matrices @ source_extrinsic_matrices
        return deviation(cycle) / (deviation(source_extrinsic_matrices) + deviation(target_extrinsic_matrices) + epsilon)
    return loss


rotation_consistency_loss = _cycle_consistency(lambda m: m[..., :3, :3], lambda part: torch.eye(3).to(part))
translation_consistency_loss = _cycle_consistency(lambda m: m[..., :3, 3], lambda part: torch.zeros(3).to(part))


@reduced
def sampson_epipolar_distance(keypoints_1, keypoints_2, fundamental_matrices):
    x1 = nn.functional.pad(keypoints_1, (0, 1), mode="constant", value=1.0)
    x2 = nn.functional.pad(keypoints_2, (0, 1), mode="constant", value=1.0)
    lines_2 = x1 @ fundamental_matrices.transpose(-2, -1)
    lines_1 = x2 @ fundamental_matrices
    algebraic = torch.sum(x2 * lines_2, dim=-1) ** 2.0
    return algebraic / (torch.sum(lines_2[..., :2] ** 2.0, dim=-1) + torch.sum(lines_1[..., :2] ** 2.0, dim=-1))


# ---- photometric / smoothness losses ---------------------------------------------------------------------------------

@reduced
def ssim_loss(inputs, targets, C1=0.01 ** 2, C2=0.03 ** 2, kernel_size=3, stride=1, padding=1, padding_mode="reflect"):
    x = nn.functional.pad(inputs, [padding] * 4, padding_mode)
    y = nn.functional.pad(targets, [padding] * 4, padding_mode)
    pool = functools.partial(nn.functional.avg_pool2d, kernel_size=kernel_size, stride=stride)
    mu_x, mu_y = pool(x), pool(y)
    sigma_xx, sigma_yy, sigma_xy = pool(x * x) - mu_x * mu_x, pool(y * y) - mu_y * mu_y, pool(x * y) - mu_x * mu_y
    ssim = ((2.0 * mu_x * mu_y + C1) / (mu_x * mu_x + mu_y * mu_y + C1)) * ((2.0 * sigma_xy + C2) / (sigma_xx + sigma_yy + C2))
    return torch.clamp((1.0 - ssim) / 2.0, 0.0, 1.0)


@reduced
def photometric_loss(inputs, targets, alpha=0.75):
    return (ssim_loss(inputs, targets, reduction="none") * alpha
            + nn.functional.smooth_l1_loss(inputs, targets, reduction="none") * (1.0 - alpha))


def gradient_x(inputs, padding=(0, 1), padding_mode="replicate"):
    inputs = nn.functional.pad(inputs, (*padding, 0, 0), padding_mode)
    return inputs[..., :, 1:] - inputs[..., :, :-1]


def gradient_y(inputs, padding=(0, 1), padding_mode="replicate"):
    inputs = nn.functional.pad(inputs, (0, 0, *padding), padding_mode)
    return inputs[..., 1:, :] - inputs[..., :-1, :]


@reduced
def smoothness_loss(inputs, references, normalize=True, epsilon=1e-6):
    """Edge-aware first-order smoothness: |grad input| weighted by exp(-mean_c |grad reference|)."""
    if normalize:
        inputs = inputs / (torch.mean(inputs, dim=(-2, -1), keepdim=True) + epsilon)
    terms = []
    for gradient in (gradient_x, gradient_y):
        weight = torch.exp(-torch.mean(torch.abs(gradient(references)), dim=1, keepdim=True))
        terms.append(torch.abs(gradient(inputs)) * weight)
    return terms[0] + terms[1]


@reduced
def motion_smoothness_loss(inputs, epsilon=1e-6):
    return torch.sqrt(gradient_x(inputs) ** 2.0 + gradient_y(inputs) ** 2.0 + epsilon)


@reduced
def motion_sparsity_loss(inputs, epsilon=1e-6):
    with torch.no_grad():
        means = torch.mean(torch.abs(inputs), dim=(-2, -1), keepdim=True)
    return torch.sqrt(torch.abs(inputs) * means + means * means + epsilon)


# ---- the losses of the per-frame optimisation step (inline in scripts/main.py) ---------------------------------------

def silhouette_loss(labels, targets, pd_indices=None, gt_indices=None):
    """main.py:653-671: mean BCE between the rendered soft instance labels [R,N] (clamped to [1e-6, 1-1e-6]) and the
    soft masks at the same rays, instances paired by the bipartite matching."""
    if pd_indices is not None:
        labels, targets = labels[..., pd_indices], targets[..., gt_indices]
    return nn.functional.binary_cross_entropy(labels.clamp(1.0e-6, 1.0 - 1.0e-6), targets, reduction="none").mean()


def eikonal_loss(sampled_gradients):
    """main.py:679-687: mean (|grad d| - 1)^2 over the fine samples."""
    norms = torch.norm(sampled_gradients, dim=-1)
    return nn.functional.mse_loss(norms, torch.ones_like(norms), reduction="mean")


def projection_losses(world_boxes_3d, extrinsic_matrices, intrinsic_matrices, image_size, gt_boxes_2d, visible_masks,
                      target_view=None):
    """main.py:339-415 for every view of the frame in ONE call: project the predicted boxes [N,8,3] into all V views,
    match predictions to the target view's 2D boxes (DIoU cost, Hungarian), and return
    `(iou_projection_loss, l1_projection_loss, gt_indices)` — mean DIoU loss and mean smooth-L1 over the visible
    (view, instance) pairs.  CUDA only: projection, matching, both losses and their adjoint are one kernel launch
    (`projection_step_kernel`); differentiable w.r.t. `world_boxes_3d`."""
    from vsrd_b200 import frame as _frame
    from vsrd_b200 import ops
    if not world_boxes_3d.is_cuda:
        raise RuntimeError("vsrd.losses.projection_losses runs on the CUDA kernels only (no CPU fallback)")
    num_views = extrinsic_matrices.shape[0]
    views = ops.ViewArgs(extrinsic_matrices, intrinsic_matrices, image_size,
                         num_views // 2 if target_view is None else int(target_view))
    losses, gt_indices = _frame._ProjectionLosses.apply(world_boxes_3d, views, gt_boxes_2d.reshape(num_views, -1, 4),
                                                        visible_masks)
    return losses[0], losses[1], gt_indices


def fused_silhouette_eikonal_loss(*args, **kwargs):
    """`vsrd_b200.functional.fused_render_loss`: the fine pass with silhouette BCE + weighted eikonal reduced inside the
    compositing kernels; returns (loss, labels, [silhouette, eikonal] parts)."""
    from vsrd_b200 import functional
    return functional.fused_render_loss(*args, **kwargs)
