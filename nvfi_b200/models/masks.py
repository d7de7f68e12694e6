"""MaskField parameter container (reference: models/mask_field.py:34-83).

On the render path the field is evaluated inside the appearance kernel (softmax mask of
the advected sample position, composited with the ray weights).  ``forward`` below is the
stand-alone module call used by the segmentation trainer's autograd, which is outside
the hot path; it is plain torch on purpose.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F


class MaskField(nn.Module):
    def __init__(self, n_layer=8, n_dim=256, input_dim=3, skips=[4], mask_dim=2, mask_act="softmax",
                 point_embed=False):
        super().__init__()
        if point_embed:
            raise NotImplementedError("nvfi_b200: MaskField(point_embed=True) is not supported")
        self.skips = list(skips)
        self.mask_dim = mask_dim
        self.mask_act_name = mask_act
        self.point_embed = None
        self.point_fc = nn.ModuleList([nn.Linear(input_dim, n_dim)])
        for l in range(n_layer - 1):
            self.point_fc.append(nn.Linear(n_dim + input_dim if l in self.skips else n_dim, n_dim))
        self.mask_fc = nn.Linear(n_dim, mask_dim)

    def forward(self, point):
        h = point
        for l, fc in enumerate(self.point_fc):
            h = F.relu(fc(h))
            if l in self.skips:
                h = torch.cat([point, h], 1)
        m = self.mask_fc(h)
        if self.mask_act_name == "softmax":
            return F.softmax(m, dim=1)
        if self.mask_act_name == "sigmoid":
            return torch.sigmoid(m)
        return m
