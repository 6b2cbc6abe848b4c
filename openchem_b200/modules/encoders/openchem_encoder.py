"""Common base of the encoders on this path, behaviour-compatible with the reference's
openchem/modules/encoders/openchem_encoder.py:7-28 (what subclasses and `Graph2Label` rely on):

* `params` is kept as given; `input_size` / `encoder_dim` are read from it and exposed as attributes;
* `use_cuda=None` means "whatever torch reports";
* the dict returned by the (possibly overridden) `get_required_params()` is validated as BOTH the required and the optional
  specification — the reference passes it twice (:11); subclasses then validate their own optional keys themselves.
"""
from __future__ import annotations

import torch
from torch import nn

from ...utils import check_params

_SIZE_KEYS = ("input_size", "encoder_dim")


class OpenChemEncoder(nn.Module):
    def __init__(self, params, use_cuda=None):
        nn.Module.__init__(self)
        spec = self.get_required_params()
        check_params(params, spec, spec)
        self.params = params
        self.use_cuda = use_cuda if use_cuda is not None else torch.cuda.is_available()
        for key in _SIZE_KEYS:
            setattr(self, key, params[key])

    @staticmethod
    def get_required_params():
        return dict.fromkeys(_SIZE_KEYS, int)

    @staticmethod
    def get_optional_params():
        return dict()

    def forward(self, inp):
        raise NotImplementedError("OpenChemEncoder is abstract")
