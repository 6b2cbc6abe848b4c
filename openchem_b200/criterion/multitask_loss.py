"""MultitaskLoss with the reference's surface and as-executed arithmetic (openchem/criterion/multitask_loss.py:40-49),
without its hard-coded `.cuda()` calls (:43-44), so it also works on whatever device the tensors live on.

As executed, the reference calls `F.binary_cross_entropy(input, mask * target, weight)` with the default 'mean' reduction
(:46; the constructor's reduction='none' is never passed on), so the criterion is the mean BCE over ALL B*T entries with
ignored labels counted as class 0, times mean_j(n_j / n_j) (nan when a task has no valid sample). This class keeps that
behaviour bit for bit in structure; `openchem_b200.modules.mlp.openchem_mlp.fused_head_loss` recognises it (and the
reference's own class) and runs it inside the head's last kernel.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch.nn.modules.loss import _WeightedLoss


class MultitaskLoss(_WeightedLoss):
    def __init__(self, ignore_index, n_tasks):
        super().__init__(reduction="none")
        self.n_tasks = n_tasks
        self.ignore_index = ignore_index

    def forward(self, input, target):
        assert target.size()[1] == self.n_tasks
        assert input.size()[1] == self.n_tasks
        mask = torch.where(target == self.ignore_index, torch.zeros_like(target), torch.ones_like(target))
        loss = F.binary_cross_entropy(input, mask * target, weight=self.weight)
        loss = loss * mask
        n_samples = mask.sum(dim=0)
        return (loss.sum(dim=0) / n_samples).mean()
