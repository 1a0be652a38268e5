"""CMU-MOSEI / MOSI modality-level DynMM (ModalityDynMM/affect/affect_dyn.py:31-175)."""
from __future__ import annotations

import torch
import torch.nn as nn

from .common_models import MLP, Concat, Transformer
from .gating import DiffSoftmax, can_route, mix, routed_mix
from .supervised import MMDL


def _load(path):
    return torch.load(path, weights_only=False)


def build_mosei_experts():
    """The architectures the reference trains and pickles (affect_uni.py:69-73, affect_mm.py:61-66)."""
    text_encoder, text_head = Transformer(300, 120), MLP(120, 64, 1)
    branch2 = MMDL([Transformer(35, 60), Transformer(74, 120), Transformer(300, 120)], Concat(), MLP(300, 128, 1),
                   has_padding=True)
    return text_encoder, text_head, branch2


class DynMMNetV2(nn.Module):
    """Expert 1 = text Transformer, expert 2 = 3-modality late fusion, gate = Transformer(409,10)+Linear."""

    def __init__(self, temp, hard_gate, freeze, model_name_list):
        super().__init__()
        self.branch_num = 2
        if model_name_list is None:
            self.text_encoder, self.text_head, self.branch2 = build_mosei_experts()
        else:
            self.text_encoder = _load(model_name_list[0])
            self.text_head = _load(model_name_list[0].replace('encoder', 'head'))
            self.branch2 = _load(model_name_list[1])
        if freeze:
            for m in (self.text_encoder, self.text_head, self.branch2):
                self.freeze_branch(m)
        self.gate = nn.Sequential(Transformer(409, 10), nn.Linear(10, self.branch_num))
        self.temp = temp
        self.hard_gate = hard_gate
        self.weight_list = torch.Tensor()
        self.store_weight = False
        self.infer_mode = 0
        self.flop = torch.Tensor([135.13226, 320.03205])
        self.last_route_counts = None

    def freeze_branch(self, m):
        for param in m.parameters():
            param.requires_grad = False

    def reset_weight(self):
        self.weight_list = torch.Tensor()
        self.store_weight = True

    def weight_stat(self):
        print(self.weight_list)
        tmp = torch.mean(self.weight_list, dim=0)
        print(f'mean branch weight {tmp[0].item():.4f}, {tmp[1].item():.4f}')
        self.store_weight = False
        return tmp[1].item()

    def cal_flop(self):
        tmp = torch.mean(self.weight_list, dim=0)
        total_flop = (self.flop * tmp).sum()
        print(f'Total Flops {total_flop.item():.2f}M')
        return total_flop.item()

    def _expert1(self, feats, lens):
        return self.text_head(self.text_encoder([feats[2], lens[2]]))

    def forward(self, inputs):
        feats, lens = inputs[0], inputs[1]
        x = torch.cat(feats, dim=2)
        weight = DiffSoftmax(self.gate([x, lens[0]]), tau=self.temp, hard=self.hard_gate)
        if self.store_weight:
            self.weight_list = torch.cat((self.weight_list, weight.detach().cpu()))
        if self.infer_mode > 0:
            pred = self._expert1(feats, lens) if self.infer_mode == 1 else self.branch2(inputs)
            return pred, 0
        if self.infer_mode == -1:
            weight = torch.ones_like(weight) / self.branch_num
        if self.infer_mode == 0 and can_route(weight, self.hard_gate, self.training):
            def sub(rows):
                pick = lambda ln: ln[rows.to(ln.device)] if torch.is_tensor(ln) else ln   # lengths may live on the host
                return [[f[rows] for f in feats], [pick(ln) for ln in lens]]
            experts = [lambda rows: self._expert1(*sub(rows)), lambda rows: self.branch2(sub(rows))]
            output, self.last_route_counts = routed_mix(weight, experts, 1)
        else:
            output = mix(weight, [self._expert1(feats, lens), self.branch2(inputs)])
        return output, weight[:, 1].mean()

    def forward_separate_branch(self, inputs, path, weight_enable):
        if weight_enable:
            x = torch.cat(inputs[0], dim=2)
            DiffSoftmax(self.gate([x, inputs[1][0]]), tau=self.temp, hard=self.hard_gate)
        if path == 1:
            return self._expert1(inputs[0], inputs[1])
        return self.branch2(inputs)


class DynMMNet(nn.Module):
    """Three unimodal branches with a 3-way gate (affect_dyn.py:31-104)."""

    def __init__(self, temp, hard_gate, freeze=True, model_name_list=None, encoders=None, heads=None):
        super().__init__()
        self.branch_num = 3
        if model_name_list is not None:
            encoders = [_load(n) for n in model_name_list]
            heads = [_load(n.replace('encoder', 'head')) for n in model_name_list]
        self.encoders, self.heads = nn.ModuleList(encoders), nn.ModuleList(heads)
        if freeze:
            self.freeze_model()
        self.gate = nn.Sequential(Transformer(409, 10), nn.Linear(10, self.branch_num))
        self.temp = temp
        self.hard_gate = hard_gate
        self.weight_list = torch.Tensor()
        self.store_weight = False
        self.infer_mode = 0

    def freeze_model(self):
        for m in (self.encoders, self.heads):
            for param in m.parameters():
                param.requires_grad = False

    def reset_weight(self):
        self.weight_list = torch.Tensor()
        self.store_weight = True

    def weight_stat(self):
        print(self.weight_list)
        tmp = torch.mean(self.weight_list, dim=0)
        print(f'mean branch weight {tmp[0].item():.4f}, {tmp[1].item():.4f}, {tmp[2].item():.4f}')
        self.store_weight = False

    def forward2(self, inputs):
        x = torch.cat(inputs[0], dim=2)
        weight = DiffSoftmax(self.gate([x, inputs[1][0]]), tau=self.temp, hard=self.hard_gate)
        if self.store_weight:
            self.weight_list = torch.cat((self.weight_list, weight.detach().cpu()))
        preds = [self.heads[i](self.encoders[i]([inputs[0][i], inputs[1][i]])) for i in range(len(inputs[0]))]
        if self.infer_mode > 0:
            return preds[self.infer_mode - 1]
        if self.infer_mode == -1:
            weight = torch.ones_like(weight) / self.branch_num
        return mix(weight, preds), weight[:, 2].mean()

    def forward(self, inputs, path, weight_enable):
        if weight_enable:
            x = torch.cat(inputs[0], dim=2)
            DiffSoftmax(self.gate([x, inputs[1][0]]), tau=self.temp, hard=self.hard_gate)
        return self.heads[path](self.encoders[path]([inputs[0][path], inputs[1][path]]))
