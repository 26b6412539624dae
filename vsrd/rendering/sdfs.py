"""`vsrd.rendering.sdfs` — tagged SDF leaves.

Same call protocol as the reference (vsrd/rendering/sdfs.py:5-37): `box(dimension)`,
`translation(sdf, t)`, `rotation(sdf, R)` return callables `positions -> distances`.  Here they are
small objects instead of bare closures so the renderer can read their parameters back and run the
whole composed field in the fused kernels; called directly they evaluate in plain PyTorch (used for
ad-hoc queries, any device).
"""
import torch


class BoxSDF:
    """`box(dimension)`: `sqrt(sum relu(|p| - dim)^2 + 1e-6) - relu(-max(|p| - dim))`."""

    def __init__(self, dimension):
        self.dimension = dimension

    def __call__(self, positions):
        q = positions.abs() - self.dimension
        outside = torch.sqrt(torch.sum(torch.relu(q) ** 2.0, dim=-1, keepdim=True) + 1e-6)
        inside = torch.relu(-q.max(dim=-1, keepdim=True).values)
        return outside - inside


class TranslatedSDF:

    def __init__(self, sdf, translation_vector):
        self.sdf = sdf
        self.translation_vector = translation_vector

    def __call__(self, positions):
        return self.sdf(positions - self.translation_vector)


class RotatedSDF:

    def __init__(self, sdf, rotation_matrix):
        self.sdf = sdf
        self.rotation_matrix = rotation_matrix

    def __call__(self, positions):
        return self.sdf(positions @ self.rotation_matrix)


def box(dimension):
    return BoxSDF(dimension)


def translation(sdf, translation_vector):
    return TranslatedSDF(sdf, translation_vector)


def rotation(sdf, rotation_matrix):
    return RotatedSDF(sdf, rotation_matrix)
