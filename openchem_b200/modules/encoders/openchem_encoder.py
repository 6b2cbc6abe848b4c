"""Encoder base class with the reference's surface (openchem/modules/encoders/openchem_encoder.py:7-28)."""
from __future__ import annotations

import torch
from torch import nn

from ...utils import check_params


class OpenChemEncoder(nn.Module):
    def __init__(self, params, use_cuda=None):
        super().__init__()
        check_params(params, self.get_required_params(), self.get_required_params())
        self.params = params
        self.use_cuda = torch.cuda.is_available() if use_cuda is None else use_cuda
        self.input_size = self.params["input_size"]
        self.encoder_dim = self.params["encoder_dim"]

    @staticmethod
    def get_required_params():
        return {"input_size": int, "encoder_dim": int}

    @staticmethod
    def get_optional_params():
        return {}

    def forward(self, inp):
        raise NotImplementedError
