"""OpenChemMLP with the reference's surface (openchem/modules/mlp/openchem_mlp.py:9-61), executed by libocgcn.so as one
kernel per layer and direction (ocgcn_head_forward / ocgcn_head_backward).

forward(inp): for every layer but the last  Dropout -> Linear -> BatchNorm1d -> activation[i];  last layer
Dropout -> Linear -> activation[-1]  (:51-61). Same constructor, attributes (`hidden_size`, `input_size`, `n_layers`,
`activation`, `dropout`, `layers`, `bn`, `dropouts`), `state_dict` keys and parameter-initialisation RNG order (:29-36), so
reference checkpoints load unchanged.

Supported by the kernels: 1-4 layers, widths <= 256, activations relu / identity / sigmoid / tanh (the callables the
reference's configs use, recognised by name), inactive dropout (p = 0 or eval mode — every GraphCNN config of the
reference). Anything else raises at construction / call time: there is no silent fallback.

`fused_head_loss(mlp, inp, target, criterion)` additionally runs nn.MSELoss() or MultitaskLoss inside the last kernel.
`OpenChemMLPSimple` (:64-111: Linear -> activation per layer, no BatchNorm) is mirrored on the same kernels.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from ...functional import HeadFn, HeadSpec
from ...utils import check_params

_ACT_NAMES = {"relu": "relu", "identity": "identity", "sigmoid": "sigmoid", "tanh": "tanh",
              "ReLU": "relu", "Identity": "identity", "Sigmoid": "sigmoid", "Tanh": "tanh"}


def _act_name(fn):
    if isinstance(fn, str):
        name = fn
    elif isinstance(fn, nn.Module):
        name = type(fn).__name__
    else:
        name = getattr(fn, "__name__", None)
    if name not in _ACT_NAMES:
        raise NotImplementedError("openchem_b200.OpenChemMLP: activation %r is not one of relu / identity / sigmoid / tanh; "
                                  "use the reference's OpenChemMLP for it" % (fn,))
    return _ACT_NAMES[name]


class OpenChemMLP(nn.Module):
    """Base class for MLP module"""

    def __init__(self, params):
        super().__init__()
        check_params(params, self.get_required_params(), self.get_optional_params())
        self.params = params
        self.hidden_size = self.params["hidden_size"]
        self.input_size = [self.params["input_size"]] + self.hidden_size[:-1]
        self.n_layers = self.params["n_layers"]
        self.activation = self.params["activation"]
        if type(self.activation) is list:
            assert len(self.activation) == self.n_layers
        else:
            self.activation = [self.activation] * self.n_layers
        self.dropout = self.params["dropout"] if "dropout" in self.params.keys() else 0
        self.layers = nn.ModuleList([])
        self.bn = nn.ModuleList([])
        self.dropouts = nn.ModuleList([])
        for i in range(self.n_layers - 1):            # creation order = the reference's (:29-33): same RNG consumption
            self.dropouts.append(nn.Dropout(self.dropout))
            self.bn.append(nn.BatchNorm1d(self.hidden_size[i]))
            self.layers.append(nn.Linear(in_features=self.input_size[i], out_features=self.hidden_size[i]))
        i = self.n_layers - 1
        self.dropouts.append(nn.Dropout(self.dropout))
        self.layers.append(nn.Linear(in_features=self.input_size[i], out_features=self.hidden_size[i]))
        self._acts = [_act_name(a) for a in self.activation]

    @staticmethod
    def get_required_params():
        return {"input_size": int, "n_layers": int, "hidden_size": list, "activation": None}

    @staticmethod
    def get_optional_params():
        return {"dropout": float}

    def _spec(self, loss="none", ignore_index=0.0):
        if self.training and self.dropout > 0:
            raise NotImplementedError("openchem_b200.OpenChemMLP: training-mode dropout (p = %g) is not implemented by the fused "
                                      "head; use the reference's OpenChemMLP for it" % self.dropout)
        bn0 = self.bn[0] if len(self.bn) else None
        for b in self.bn:
            if b.momentum is None or not b.track_running_stats or not b.affine:
                raise NotImplementedError("openchem_b200.OpenChemMLP: BatchNorm1d must be the default (affine, running stats, momentum)")
        return HeadSpec([self.input_size[0]] + list(self.hidden_size), self._acts, self.training,
                        bn0.eps if bn0 is not None else 1e-5, bn0.momentum if bn0 is not None else 0.1,
                        [b.running_mean for b in self.bn], [b.running_var for b in self.bn], [b.num_batches_tracked for b in self.bn],
                        loss, ignore_index)

    def _tensors(self):
        return ([l.weight for l in self.layers] + [l.bias for l in self.layers] + [b.weight for b in self.bn] + [b.bias for b in self.bn])

    def forward(self, inp):
        return HeadFn.apply(inp, None, self._spec(), *self._tensors())


class OpenChemMLPSimple(nn.Module):
    """OpenChemMLPSimple with the reference's surface (openchem/modules/mlp/openchem_mlp.py:64-111): every layer is
    Linear -> activation[i], no BatchNorm, no dropout; optional 'init': 'xavier_uniform' (gain of relu) applied to every Linear in
    module order (:86-94). Same attributes and `state_dict` keys (`layers.{i}.weight/bias`). Each layer runs as one head kernel per
    direction (a one-layer ocgcn_head_forward / ocgcn_head_backward call whose epilogue applies the activation)."""

    def __init__(self, params):
        super().__init__()
        check_params(params, self.get_required_params(), self.get_optional_params())
        self.params = params
        self.hidden_size = self.params["hidden_size"]
        self.input_size = [self.params["input_size"]] + self.hidden_size[:-1]
        self.n_layers = self.params["n_layers"]
        self.activation = self.params["activation"]
        if type(self.activation) is list:
            assert len(self.activation) == self.n_layers
        else:
            self.activation = [self.activation] * self.n_layers
        self.layers = nn.ModuleList([])
        for i in range(self.n_layers):
            self.layers.append(nn.Linear(in_features=self.input_size[i], out_features=self.hidden_size[i]))
        if "init" in self.params.keys():
            if self.params["init"] == "xavier_uniform":
                for m in self.modules():
                    if isinstance(m, nn.Linear):
                        nn.init.xavier_uniform_(m.weight, gain=nn.init.calculate_gain("relu"))
            else:
                raise NotImplementedError("Only xavier_uniform initialization is supported now in OpenChemMLPSimple")
        self._acts = [_act_name(a) for a in self.activation]

    @staticmethod
    def get_required_params():
        return {"input_size": int, "n_layers": int, "hidden_size": list, "activation": None}

    @staticmethod
    def get_optional_params():
        return {"init": str}

    def forward(self, inp):
        output = inp
        for i, layer in enumerate(self.layers):
            spec = HeadSpec([self.input_size[i], self.hidden_size[i]], [self._acts[i]], self.training, 1e-5, 0.1, [], [], [])
            output = HeadFn.apply(output, None, spec, layer.weight, layer.bias)
        return output


def fused_head_loss(mlp, inp, target, criterion):
    """-> (loss, predictions): `criterion(mlp(inp), target)` with the criterion evaluated inside the head's last kernel.
    criterion: nn.MSELoss() (mean reduction) or a MultitaskLoss (the reference's or openchem_b200.criterion's; weight=None)."""
    name = type(criterion).__name__
    if isinstance(criterion, nn.MSELoss) and criterion.reduction == "mean":
        spec = mlp._spec("mse")
    elif name == "MultitaskLoss" and getattr(criterion, "weight", None) is None:
        if criterion.n_tasks != mlp.hidden_size[-1]:
            raise ValueError("MultitaskLoss.n_tasks does not match the head's output width")
        spec = mlp._spec("multitask", float(criterion.ignore_index))
    else:
        raise NotImplementedError("openchem_b200.fused_head_loss: criterion %r is not fused; call it on mlp(inp) instead" % (criterion,))
    return HeadFn.apply(inp, target, spec, *mlp._tensors())
