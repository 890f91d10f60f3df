"""Functional ops as modules so that they can be swapped for quantised versions / hooked for statistics
(mirror of mobilellm/model/ops.py:6-60; same class names so act_dict.json keys line up)."""
import torch
import torch.nn as nn


class L2Norm(nn.Module):
    def __init__(self, p=2, dim=-1, eps=1e-12):
        super().__init__()
        self.p, self.dim, self.eps = p, dim, eps

    def forward(self, x):
        return torch.nn.functional.normalize(x, p=self.p, dim=self.dim, eps=self.eps)


class ElementwiseAdd(nn.Module):
    def forward(self, x, y):
        return x + y


class ElementwiseMul(nn.Module):
    def forward(self, x, y):
        return x * y


class FMatMul(nn.Module):
    def forward(self, a, b):
        return torch.matmul(a, b)


class FCat(nn.Module):
    def __init__(self, axis=0):
        super().__init__()
        self._axis = axis

    def forward(self, *x):
        return torch.cat(x, dim=self._axis)
