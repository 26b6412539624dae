"""`SinusoidalEncoder` — API of vsrd/models/encoders/sinusoidal_encoder.py:6-19.

Inside the renderer the encoding is evaluated in-kernel; this module exists for the constructor /
state-dict contract (`frequencies` buffer) and for direct calls."""
import math

import torch
import torch.nn as nn


class SinusoidalEncoder(nn.Module):

    def __init__(self, num_frequencies):
        super().__init__()
        self.register_buffer("frequencies", torch.exp2(torch.arange(num_frequencies).float()) * math.pi)

    def forward(self, inputs):
        phase = inputs.unsqueeze(-1) * self.frequencies
        return torch.stack([phase.cos(), phase.sin()], dim=-1).flatten(-3, -1)
