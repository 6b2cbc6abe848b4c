"""Multi-task BCE criterion of the Tox21-style configs, as a parameter holder for the fused head kernel.

`fused_head_loss(mlp, inp, target, MultitaskLoss(ignore_index, n_tasks))` evaluates the criterion inside the MLP head's last
kernel (csrc/ocgcn_head.cuh, `head_fwd_kernel` + `head_loss_kernel`); that is the product path. Calling the object directly on
predictions is supported for code that follows the reference's `criterion(output, target)` convention and evaluates the same
number in closed form.

What the reference's criterion computes AS EXECUTED (openchem/criterion/multitask_loss.py:40-49): its
`F.binary_cross_entropy` call uses the default 'mean' reduction (:46 - the constructor's reduction='none' is never passed
on), so the "per-element" loss it then masks and averages per task is already ONE scalar, the mean BCE over all B*T entries
with ignored labels counted as class 0. Multiplying by the mask, summing a column and dividing by that column's valid count
n_j gives the same scalar back (or nan when n_j == 0), and so does the mean over tasks:

    loss = mean_{b,j} BCE(p_bj, [t_bj valid] * t_bj)  *  mean_j(n_j / n_j)

Both this class and the kernel implement that expression (parity: tests/test_head_host.py against fixtures written from the
reference's class). Unlike the reference there is no hard-coded `.cuda()` (:43-44): tensors stay on their own device.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch import nn


class MultitaskLoss(nn.Module):
    weight = None          # the reference's _WeightedLoss weight (always None there); fused_head_loss requires None
    reduction = "none"     # what the reference's constructor records (and never uses)

    def __init__(self, ignore_index, n_tasks):
        super().__init__()
        self.ignore_index, self.n_tasks = ignore_index, int(n_tasks)

    def forward(self, input, target):
        if input.dim() != 2 or target.shape != input.shape or input.shape[1] != self.n_tasks:
            raise AssertionError("MultitaskLoss: input and target must both be (B, %d); got %s and %s" %
                                 (self.n_tasks, tuple(input.shape), tuple(target.shape)))
        valid = target != self.ignore_index
        mean_bce = F.binary_cross_entropy(input, target.masked_fill(~valid, 0.0))
        n_valid = valid.sum(dim=0).to(input.dtype)
        return mean_bce * (n_valid / n_valid).mean()      # nan when a task has no valid label, as in the reference

    def extra_repr(self):
        return "ignore_index=%r, n_tasks=%d" % (self.ignore_index, self.n_tasks)
