"""ResnetBlockFC with the reference's parameter names (reference src/layers.py:8-50).

Inside LocalPoolPointnet / LocalDecoder the block is executed by the fused CUDA
kernels, which read its parameters directly; `forward` below exists so the block
stays usable stand-alone (plain torch ops — it is not on the hot path)."""
import torch.nn as nn
import torch.nn.functional as F


class ResnetBlockFC(nn.Module):
    def __init__(self, size_in, size_out=None, size_h=None):
        super().__init__()
        if size_out is None:
            size_out = size_in
        if size_h is None:
            size_h = min(size_in, size_out)
        self.size_in, self.size_h, self.size_out = size_in, size_h, size_out
        self.fc_0 = nn.Linear(size_in, size_h)
        self.fc_1 = nn.Linear(size_h, size_out)
        self.actvn = nn.ReLU()
        self.shortcut = None if size_in == size_out else nn.Linear(size_in, size_out, bias=False)
        nn.init.zeros_(self.fc_1.weight)  # reference src/layers.py:39

    def forward(self, x):
        net = self.fc_0(F.relu(x))
        dx = self.fc_1(F.relu(net))
        x_s = self.shortcut(x) if self.shortcut is not None else x
        return x_s + dx
