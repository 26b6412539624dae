"""`HyperDistanceField` — API of vsrd/models/fields/hyper_distance_field.py:7-77.

The hypernetwork (embeddings -> per-instance MLP weights) stays in PyTorch/cuBLAS: it is ~4 MFLOP
per instance per step.  The functional per-instance MLP (`distance_field`) is what the fused
kernels evaluate per sample; the method here is its plain-PyTorch form for direct calls, and
`check_fused_layout` verifies the architecture matches what the kernels are compiled for.
State-dict keys match the reference (weight-norm `weight_g` / `weight_v`)."""
import torch
import torch.nn as nn

from vsrd_b200 import ops


class HyperDistanceField(nn.Module):

    def __init__(self, in_channels, out_channels_list, hyper_in_channels, hyper_out_channels_list):
        super().__init__()
        fan_in = [in_channels, *out_channels_list]
        fan_out = [*out_channels_list, 1]
        self.in_channels_list = fan_in
        self.out_channels_list = fan_out
        self.num_neurons_list = [o * (i + 1) for i, o in zip(fan_in, fan_out)]

        widths = [hyper_in_channels, *hyper_out_channels_list]
        blocks = [
            nn.Sequential(nn.Linear(a, b), nn.LayerNorm(b), nn.GELU())
            for a, b in zip(widths[:-1], widths[1:])
        ]
        blocks.append(nn.Sequential(nn.Linear(widths[-1], sum(self.num_neurons_list))))
        self.hypernetwork = nn.Sequential(*blocks)
        for module in list(self.modules()):
            if isinstance(module, nn.Linear):
                nn.utils.weight_norm(module)

    def check_fused_layout(self, positional_encoder=None):
        if self.in_channels_list != [48, 16, 16, 16, 16] or sum(self.num_neurons_list) != ops.MLP_WEIGHTS:
            raise RuntimeError(
                "vsrd_b200: the fused kernels are compiled for the 48-16-16-16-16-1 residual field of "
                f"configs/kitti_360 (got fan-in {self.in_channels_list}); rebuild csrc for other sizes")
        if positional_encoder is not None and positional_encoder.frequencies.numel() != 8:
            raise RuntimeError("vsrd_b200: the fused kernels are compiled for SinusoidalEncoder(num_frequencies=8)")

    def distance_field(self, weights, positions):
        h = positions
        blocks = torch.split(weights, self.num_neurons_list, dim=-1)
        for layer, (block, n_in, n_out) in enumerate(zip(blocks, self.in_channels_list, self.out_channels_list)):
            if layer:
                h = nn.functional.gelu(nn.functional.layer_norm(h, [n_in]))
            mat = block.unflatten(-1, (n_out, n_in + 1))
            h = (mat[..., :-1] @ h.unsqueeze(-1)).squeeze(-1) + mat[..., -1]
        return h

    def forward(self, embeddings):
        if embeddings.is_cuda:
            # the configs/kitti_360 architecture runs as 5 launches (6 backward) of csrc/vsrd_model.cu under autograd;
            # any other architecture / batch shape is plain nn.Sequential on the GPU
            from vsrd_b200 import functional
            weights = functional.hypernetwork(self, embeddings) if functional.fused_modules() else None
            if weights is not None:
                return weights
        return self.hypernetwork(embeddings)
