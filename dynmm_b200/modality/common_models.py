"""The MultiBench building blocks ModalityDynMM is assembled from.

The reference imports these from an external MultiBench checkout
(``unimodals.common_models`` / ``fusions.common_fusions``; imdb_dyn.py:10-13,
affect_dyn.py:12-15) that is neither vendored nor pinned.  They are provided
here with the attribute names MultiBench uses (``fc``/``fc2``, ``lin``,
``op0..op4``/``hid2val``, ``conv``/``transformer``) so ``state_dict`` keys and
whole-module pickles line up; :func:`dynmm_b200.modality.compat.install_aliases`
registers them under MultiBench's module paths.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F


class MLP(nn.Module):
    """fc -> ReLU -> (dropout) -> fc2 -> (dropout)."""

    def __init__(self, indim, hiddim, outdim, dropout=False, dropoutp=0.1, output_each_layer=False):
        super().__init__()
        self.fc = nn.Linear(indim, hiddim)
        self.fc2 = nn.Linear(hiddim, outdim)
        self.dropout_layer = nn.Dropout(dropoutp)
        self.dropout = dropout
        self.output_each_layer = output_each_layer
        self.lklu = nn.LeakyReLU(0.2)

    def forward(self, x):
        h = F.relu(self.fc(x))
        if self.dropout:
            h = self.dropout_layer(h)
        y = self.fc2(h)
        if self.dropout:
            y = self.dropout_layer(y)
        if self.output_each_layer:
            return [0, x, h, self.lklu(y)]
        return y


class Linear(nn.Module):
    def __init__(self, indim, outdim, xavier_init=False):
        super().__init__()
        self.fc = nn.Linear(indim, outdim)
        if xavier_init:
            nn.init.xavier_normal_(self.fc.weight)
            self.fc.bias.data.fill_(0.0)

    def forward(self, x):
        return self.fc(x)


class Maxout(nn.Module):
    """Linear(d, m*k) followed by a max over the k pieces."""

    def __init__(self, d, m, k):
        super().__init__()
        self.d_in, self.d_out, self.pool_size = d, m, k
        self.lin = nn.Linear(d, m * k)

    def forward(self, inputs):
        y = self.lin(inputs)
        return y.view(*y.shape[:-1], self.d_out, self.pool_size).max(-1)[0]


class MaxOut_MLP(nn.Module):
    def __init__(self, num_outputs, first_hidden=64, number_input_feats=300, second_hidden=None, linear_layer=True):
        super().__init__()
        second_hidden = first_hidden if second_hidden is None else second_hidden
        self.op0 = nn.BatchNorm1d(number_input_feats, 1e-4)
        self.op1 = Maxout(number_input_feats, first_hidden, 2)
        self.op2 = nn.Sequential(nn.BatchNorm1d(first_hidden), nn.Dropout(0.3))
        self.op3 = Maxout(first_hidden, second_hidden, 2)
        self.op4 = nn.Sequential(nn.BatchNorm1d(second_hidden), nn.Dropout(0.3))
        self.hid2val = nn.Linear(second_hidden, num_outputs) if linear_layer else None

    def forward(self, x):
        y = self.op4(self.op3(self.op2(self.op1(self.op0(x)))))
        return y if self.hid2val is None else self.hid2val(y)


class Transformer(nn.Module):
    """Conv1d(k=1, no bias) embedding + 5 post-norm encoder layers (nhead 5, ffn 2048); last time step."""

    def __init__(self, n_features, dim):
        super().__init__()
        self.embed_dim = dim
        self.conv = nn.Conv1d(n_features, dim, kernel_size=1, padding=0, bias=False)
        layer = nn.TransformerEncoderLayer(d_model=dim, nhead=5)
        self.transformer = nn.TransformerEncoder(layer, num_layers=5, enable_nested_tensor=False)

    def forward(self, x):
        if type(x) is list:
            x = x[0]
        x = self.conv(x.permute(0, 2, 1)).permute(2, 0, 1)
        return self.transformer(x)[-1]


class Identity(nn.Module):
    def forward(self, x):
        return x


class Sequential(nn.Sequential):
    """MultiBench's variadic Sequential (forwards extra args to the first layer only)."""

    def forward(self, *args, **kwargs):
        it = iter(self)
        x = next(it)(*args, **kwargs)
        for m in it:
            x = m(x)
        return x


class Concat(nn.Module):
    """fusions.common_fusions.Concat: flatten each modality and concatenate on dim 1."""

    def forward(self, modalities):
        return torch.cat([torch.flatten(m, start_dim=1) for m in modalities], dim=1)
