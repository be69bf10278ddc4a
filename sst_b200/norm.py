"""naiveSyncBN1d/2d/3d (mmdet3d/ops/norm.py:28-199): BatchNorm whose training statistics are averaged over ranks.
Eval mode == plain BatchNorm on running stats.  The reference does one all_gather per layer
(ops/norm.py:9-24); here the [mean ‖ meansqr] vector is all-reduced once (NCCL sum) - same numbers."""
import torch
import torch.distributed as dist
import torch.nn as nn

from .registry import NORM_LAYERS


class _AllReduceSum(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        x = x.clone()
        dist.all_reduce(x)
        return x

    @staticmethod
    def backward(ctx, g):
        g = g.clone()
        dist.all_reduce(g)
        return g


def _sync_bn_forward(self, input, dims):
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1 or not self.training:
        return None
    assert input.shape[0] > 0, "SyncBN does not support empty inputs"
    C = input.shape[1]
    mean = torch.mean(input, dim=dims)
    meansqr = torch.mean(input * input, dim=dims)
    vec = _AllReduceSum.apply(torch.cat([mean, meansqr], dim=0)) * (1.0 / dist.get_world_size())
    mean, meansqr = torch.split(vec, C)
    var = meansqr - mean * mean
    self.running_mean += self.momentum * (mean.detach() - self.running_mean)
    self.running_var += self.momentum * (var.detach() - self.running_var)
    invstd = torch.rsqrt(var + self.eps)
    scale = self.weight * invstd
    bias = self.bias - mean * scale
    return scale, bias


@NORM_LAYERS.register_module("naiveSyncBN1d")
class NaiveSyncBatchNorm1d(nn.BatchNorm1d):
    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.fp16_enabled = False

    def forward(self, input):
        input = input.float()
        if input.dim() == 2:
            r = _sync_bn_forward(self, input, [0])
            if r is None:
                return super().forward(input)
            return input * r[0][None, :] + r[1][None, :]
        r = _sync_bn_forward(self, input, [0, 2])
        if r is None:
            return super().forward(input)
        return input * r[0].reshape(1, -1, 1) + r[1].reshape(1, -1, 1)


@NORM_LAYERS.register_module("naiveSyncBN2d")
class NaiveSyncBatchNorm2d(nn.BatchNorm2d):
    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.fp16_enabled = False

    def forward(self, input):
        input = input.float()
        r = _sync_bn_forward(self, input, [0, 2, 3])
        if r is None:
            return super().forward(input)
        return input * r[0].reshape(1, -1, 1, 1) + r[1].reshape(1, -1, 1, 1)


@NORM_LAYERS.register_module("naiveSyncBN3d")
class NaiveSyncBatchNorm3d(nn.BatchNorm3d):
    """mmdet3d/ops/norm.py:145-199."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.fp16_enabled = False

    def forward(self, input):
        input = input.float()
        r = _sync_bn_forward(self, input, [0, 2, 3, 4])
        if r is None:
            return super().forward(input)
        return input * r[0].reshape(1, -1, 1, 1, 1) + r[1].reshape(1, -1, 1, 1, 1)
