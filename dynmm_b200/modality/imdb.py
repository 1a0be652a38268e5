"""MM-IMDB modality-level DynMM (ModalityDynMM/multimedia/imdb_dyn.py:29-114)."""
from __future__ import annotations

import torch
import torch.nn as nn

from .common_models import MLP, Concat, Linear, MaxOut_MLP
from .gating import DiffSoftmax, can_route, mix, routed_mix
from .supervised import MMDL


def _load(path):
    return torch.load(path, weights_only=False)      # whole-module pickles (torch >= 2.6 needs the flag)


class DynMMNet(nn.Module):
    """Expert 1 = text MLP, expert 2 = text+image late fusion, gate = MLP(4396,128,2)."""

    def __init__(self, branch_num=2, pretrain=True, freeze=True):
        super().__init__()
        self.branch_num = branch_num
        self.text_encoder = _load('./log/imdb/encoder_text.pt') if pretrain else MLP(300, 512, 512)
        self.text_head = _load('./log/imdb/head_text.pt') if pretrain else MLP(512, 512, 23)
        # image-only branch: kept for state_dict / pickle compatibility, unused in forward (imdb_dyn.py:38-41)
        self.image_encoder = _load('./log/imdb/encoder_image.pt') if pretrain else MLP(4096, 1024, 512)
        self.image_head = _load('./log/imdb/head_image.pt') if pretrain else MLP(512, 512, 23)
        if pretrain:
            self.branch3 = _load('./log/imdb/best_lf.pt')
        else:
            encoders = [MaxOut_MLP(512, 512, 300, linear_layer=False), MaxOut_MLP(512, 1024, 4096, 512, False)]
            self.branch3 = MMDL(encoders, Concat(), Linear(1024, 23), has_padding=False)
        if freeze:
            for m in (self.text_encoder, self.text_head, self.image_encoder, self.image_head, self.branch3):
                self.freeze_branch(m)
        self.gate = MLP(4396, 128, branch_num)
        self.temp = 1
        self.hard_gate = True
        self.weight_list = torch.Tensor()
        self.store_weight = False
        self.infer_mode = 0
        self.flop = torch.Tensor([1.25261, 10.86908])
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

    def _expert1(self, text):
        return self.text_head(self.text_encoder(text))

    def forward(self, inputs):
        x = torch.cat(inputs, dim=1)
        weight = DiffSoftmax(self.gate(x), tau=self.temp, hard=self.hard_gate)
        if self.store_weight:
            self.weight_list = torch.cat((self.weight_list, weight.detach().cpu()))
        if self.infer_mode > 0:
            pred = self._expert1(inputs[0]) if self.infer_mode == 1 else self.branch3(inputs)
            return pred, 0
        if can_route(weight, self.hard_gate, self.training):
            experts = [lambda rows: self._expert1(inputs[0][rows]),
                       lambda rows: self.branch3([t[rows] for t in inputs])]
            output, self.last_route_counts = routed_mix(weight, experts, 23)
        else:
            output = mix(weight, [self._expert1(inputs[0]), self.branch3(inputs)])
        return output, weight[:, 1].mean()

    def forward_separate_branch(self, inputs, path, weight_enable):
        if weight_enable:
            x = torch.cat(inputs, dim=1)
            DiffSoftmax(self.gate(x), tau=self.temp, hard=self.hard_gate)
        if path == 1:
            return self._expert1(inputs[0])
        if path == 2:
            return self.image_head(self.image_encoder(inputs[1]))
        return self.branch3(inputs)
