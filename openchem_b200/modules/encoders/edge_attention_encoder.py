"""GraphEdgeAttentionEncoder with the reference's surface (openchem/modules/encoders/edge_attention_encoder.py:9-67) on the
B200 kernels.

forward(inp) with inp = (x (B,N,F0), edge_attr (B,N,N,A)) returns (B, encoder_dim). Per layer i and edge channel j (A =
sum(edge_attr_sizes) channels): a GraphConvolution with the SOFT adjacency softmax_j'(edge_attr[:, :, :, j]) (:52-53, libocgcn's
general float-adjacency layer kernels, which also return dX for the chained layers), tanh and the neighbour max-pool over the
HARD adjacency edge_attr[..., 0] > 0 (:49,54-61, libocgcn's stand-alone pool kernels, no (B,N,N,D) temporaries); the A pooled
tensors are summed (:63). Readout tanh(dense) -> unmasked sum over nodes -> tanh (:65-66), as in GraphCNNEncoder.

Same constructor, attributes (`graph_conv` ModuleList of n_layers*A GraphConvolution, `dense`, `dropout_layer` - constructed and
never applied, as in the reference), state_dict keys and parameter-initialisation RNG order (dense first, :27, then the layers),
so reference checkpoints load. Unlike the reference there are no hard-coded `.cuda()` calls (:47-48): tensors stay on their device.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from ...functional import tanh_neighbour_max_pool
from ...layers.gcn import GraphConvolution
from ...utils import check_params
from .openchem_encoder import OpenChemEncoder


class GraphEdgeAttentionEncoder(OpenChemEncoder):
    def __init__(self, params, use_cuda):
        super().__init__(params, use_cuda)
        check_params(params, self.get_required_params(), self.get_optional_params())
        self.n_layers = params["n_layers"]
        self.attr_size = params["edge_attr_sizes"]
        self.attr_dim = sum(self.attr_size)
        self.dropout = params["dropout"] if "dropout" in params.keys() else 0
        self.precision = params.get("precision", "fp32")
        if len(params["hidden_size"]) != self.n_layers:
            raise AssertionError("hidden_size must list one width per layer")
        self.hidden_size = [self.input_size] + list(params["hidden_size"])
        self.graph_conv = nn.ModuleList()
        self.dropout_layer = nn.Dropout(p=self.dropout)
        self.dense = nn.Linear(in_features=self.hidden_size[-1], out_features=self.encoder_dim)     # before the layers: RNG order (:27-30)
        for i in range(1, self.n_layers + 1):
            for _ in range(self.attr_dim):
                self.graph_conv.append(GraphConvolution(self.hidden_size[i - 1], self.hidden_size[i], precision=self.precision))

    @staticmethod
    def get_optional_params():
        return {"dropout": float, "precision": ["fp32", "tf32", "bf16"]}

    @staticmethod
    def get_required_params():
        return {"n_layers": int, "hidden_size": list, "edge_attr_sizes": list}

    def forward(self, inp):
        x, edge_attr = inp[0], inp[1]
        if edge_attr.dim() != 4 or edge_attr.shape[3] != self.attr_dim:
            raise RuntimeError("openchem_b200: edge_attr must be (B,N,N,%d); got %s" % (self.attr_dim, tuple(edge_attr.shape)))
        hard = (edge_attr[:, :, :, 0] > 0).to(torch.float32)                       # :47-49
        soft = torch.softmax(edge_attr.to(torch.float32), dim=2)                   # :52, all channels at once; (B,N,N,A), strided views below
        cur = x
        for i in range(self.n_layers):
            acc = None
            for j in range(self.attr_dim):
                y = self.graph_conv[i * self.attr_dim + j](cur, soft[:, :, :, j])   # :53
                p = tanh_neighbour_max_pool(y, hard)                                # :54-61
                acc = p if acc is None else acc + p                                 # :62-63
            cur = acc
        d = torch.tanh(self.dense(cur))                                             # :65
        return torch.tanh(d.sum(dim=1))                                             # :66
