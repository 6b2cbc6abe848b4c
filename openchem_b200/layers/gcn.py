"""GraphConvolution with the reference's nn.Module surface (openchem/layers/gcn.py:11-47), computed by
libocgcn.so: y = BatchNorm1d_{over B*N rows}((adj @ x) @ W + b).

Same constructor, attributes (in_features, out_features, weight (F_in,F_out), bias, bn), initialisation
(RNG-compatible with gcn.py:27-31) and state_dict keys, so reference checkpoints load unchanged.
Differences, all value-preserving: the output is returned contiguous (B,N,F_out) instead of a transposed
view, and CPU tensors raise instead of silently running on the host.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
from torch.nn.parameter import Parameter

from ..functional import LayerFn, LayerSpec


class GraphConvolution(nn.Module):
    def __init__(self, in_features, out_features, bias=True, precision="fp32"):
        super().__init__()
        self.in_features = in_features
        self.out_features = out_features
        self.precision = precision
        self.weight = Parameter(torch.empty(in_features, out_features))
        if bias:
            self.bias = Parameter(torch.empty(out_features))
        else:
            self.register_parameter("bias", None)
        self.reset_parameters()
        self.bn = nn.BatchNorm1d(out_features)

    def reset_parameters(self):
        # same draws, same order as gcn.py:27-31 -> identical initial weights under the same seed
        bound = 1.0 / math.sqrt(self.weight.size(1))
        self.weight.data.uniform_(-bound, bound)
        if self.bias is not None:
            self.bias.data.uniform_(-bound, bound)

    def _spec(self):
        bn = self.bn
        if bn.momentum is None or not bn.track_running_stats or not bn.affine:
            raise RuntimeError("openchem_b200: only the default BatchNorm1d configuration of the reference is supported")
        return LayerSpec([self.in_features, self.out_features], 0, self.precision, bn.training, self.bias is not None,
                         bn.eps, bn.momentum, [bn.running_mean], [bn.running_var], [bn.num_batches_tracked])

    def forward(self, x, adj):
        return LayerFn.apply(x, adj, self._spec(), self.weight, self.bias, self.bn.weight, self.bn.bias)

    def __repr__(self):
        return "{} ({} -> {})".format(self.__class__.__name__, self.in_features, self.out_features)
