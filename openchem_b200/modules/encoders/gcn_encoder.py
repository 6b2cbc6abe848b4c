"""GraphCNNEncoder with the reference's surface (openchem/modules/encoders/gcn_encoder.py:9-53), executed as
one fused autograd node by libocgcn.so.

forward(inp) with inp = (x (B,N,F0), adj (B,N,N)) — or a bit-packed `openchem_b200.compact.CompactGraphBatch` — returns
(B, encoder_dim):
per layer  tanh(BN((adj x) W + b)) -> neighbour max-pool  (gcn_encoder.py:41-50), then
tanh(dense(.)) -> sum over ALL N rows (unmasked) -> tanh  (:51-53). dropout_layer is constructed but, as in the
reference (:22, :38-53), never applied.

Optional extra config key (unknown keys are tolerated by the reference's check_params):
    'precision': 'fp32' (default; exact-fp32 FFMA products) | 'tf32' | 'bf16' (tensor-core classes).
"""
from __future__ import annotations

import torch.nn as nn

from ...compact import CompactGraphBatch
from ...functional import EncoderFn, LayerSpec, compact_direct
from ...layers.gcn import GraphConvolution
from ...utils import check_params
from .openchem_encoder import OpenChemEncoder


class GraphCNNEncoder(OpenChemEncoder):
    def __init__(self, params, use_cuda):
        super().__init__(params, use_cuda)
        check_params(params, self.get_required_params(), self.get_optional_params())
        self.n_layers = params["n_layers"]
        self.hidden_size = params["hidden_size"]
        self.dropout = params["dropout"] if "dropout" in params.keys() else 0
        self.precision = params.get("precision", "fp32")
        assert len(self.hidden_size) == self.n_layers
        self.hidden_size = [self.input_size] + list(self.hidden_size)
        self.graph_convolutions = nn.ModuleList()
        self.dropout_layer = nn.Dropout(p=self.dropout)
        # creation order matters for RNG parity with the reference: dense first (gcn_encoder.py:23), then the layers
        self.dense = nn.Linear(in_features=self.hidden_size[-1], out_features=self.encoder_dim)
        for i in range(1, self.n_layers + 1):
            self.graph_convolutions.append(GraphConvolution(self.hidden_size[i - 1], self.hidden_size[i], precision=self.precision))

    @staticmethod
    def get_optional_params():
        return {"dropout": float, "precision": ["fp32", "tf32", "bf16"]}

    @staticmethod
    def get_required_params():
        return {"n_layers": int, "hidden_size": list}

    def forward(self, inp):
        if isinstance(inp, CompactGraphBatch):
            # bit-packed batch: the kernels read the words themselves (ocgcn_encoder_*_compact); `compact_direct["on"] = False`
            # rebuilds the dense fp32 tensors on the device first (bit-exact, same results)
            x, adj = (inp.x_bits.contiguous(), inp.adj_bits.contiguous()) if compact_direct["on"] else inp.expand()
        else:
            x, adj = inp[0], inp[1]           # dense fp32 tensors, or a (x_bits, adj_bits) pair of int32 bit words
        gcs = list(self.graph_convolutions)
        bn0 = gcs[0].bn
        spec = LayerSpec(self.hidden_size, self.encoder_dim, self.precision, self.training, gcs[0].bias is not None,
                         bn0.eps, bn0.momentum, [g.bn.running_mean for g in gcs], [g.bn.running_var for g in gcs],
                         [g.bn.num_batches_tracked for g in gcs])
        flat = []
        for g in gcs:
            flat += [g.weight, g.bias, g.bn.weight, g.bn.bias]
        return EncoderFn.apply(x, adj, spec, *flat, self.dense.weight, self.dense.bias)
