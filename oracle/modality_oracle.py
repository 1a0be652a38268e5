"""CPU oracle for modality-level DynMM (MM-IMDB / CMU-MOSEI).

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.

PARITY UNPINNED (numerically): the experts and gates of ModalityDynMM are built
from **pliang279/MultiBench** (``unimodals/common_models.py``,
``fusions/common_fusions.py``), which the reference neither vendors nor pins
(README.md:13,16-18 "clone the MultiBench repository ... copy folders"; no
commit hash).  It is absent from /root/reference, so this file restates the
published MultiBench definitions the reference's call sites rely on:

  imdb_dyn.py:34-50,60   MLP(300,512,512), MLP(512,512,23), gate MLP(4396,128,2),
                         MMDL([MaxOut_MLP(512,512,300,linear_layer=False),
                               MaxOut_MLP(512,1024,4096,512,False)], Concat(), Linear(1024,23))
  affect_dyn.py:120      gate Sequential(Transformer(409,10), nn.Linear(10,2))
  affect_uni.py:69-73    Transformer(300,120) + MLP(120,64,1)            (expert 1)
  affect_mm.py:61-66     Transformer(35,60), Transformer(74,120), Transformer(300,120),
                         Concat(), MLP(300,128,1), has_padding=True       (expert 2)

What IS pinned: the architectures.  The reference's own FLOP constants
(imdb_dyn.py:66 ``[1.25261, 10.86908]``, affect_dyn.py:126 ``[135.13226,
320.03205]`` MMACs, measured with thop) are re-derived analytically from these
definitions by :func:`imdb_mmacs` / :func:`mosei_mmacs` and asserted in
tests/test_modality_cpu.py.  The DynMM-specific arithmetic (gate, DiffSoftmax,
mixing, returned regulariser) follows the reference files line by line.
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Dict[str, Tensor]


def diff_softmax(logits: Tensor, tau: float = 1.0, hard: bool = False, dim: int = -1) -> Tensor:
    """imdb_dyn.py:16-26 / affect_dyn.py:18-28."""
    y_soft = torch.softmax(logits / tau, dim)
    if not hard:
        return y_soft
    idx = y_soft.max(dim, keepdim=True)[1]
    return torch.zeros_like(logits).scatter_(dim, idx, 1.0) - y_soft.detach() + y_soft


# ------------------------------------------------------------------ MultiBench building blocks (functional)

def linear(sd: SD, key: str, x: Tensor) -> Tensor:
    return F.linear(x, sd[key + ".weight"], sd.get(key + ".bias"))


def mlp(sd: SD, key: str, x: Tensor) -> Tensor:
    """MultiBench MLP: fc -> ReLU -> fc2 (dropout off)."""
    return linear(sd, key + ".fc2", F.relu(linear(sd, key + ".fc", x)))


def maxout(sd: SD, key: str, x: Tensor, d_out: int, k: int = 2) -> Tensor:
    """MultiBench Maxout(d, m, k): Linear(d, m*k) viewed [.., m, k], max over k."""
    y = linear(sd, key + ".lin", x)
    return y.view(*y.shape[:-1], d_out, k).max(-1)[0]


def bn1d(sd: SD, key: str, x: Tensor, eps: float) -> Tensor:
    return F.batch_norm(x, sd[key + ".running_mean"], sd[key + ".running_var"], sd[key + ".weight"],
                        sd[key + ".bias"], False, 0.1, eps)


def maxout_mlp(sd: SD, key: str, x: Tensor, first_hidden: int, second_hidden: int) -> Tensor:
    """MultiBench MaxOut_MLP(..., linear_layer=False): BN(eps 1e-4) -> Maxout -> BN -> Maxout -> BN
    (Dropout(0.3) is an identity in eval mode)."""
    x = bn1d(sd, key + ".op0", x, 1e-4)
    x = maxout(sd, key + ".op1", x, first_hidden)
    x = bn1d(sd, key + ".op2.0", x, 1e-5)
    x = maxout(sd, key + ".op3", x, second_hidden)
    return bn1d(sd, key + ".op4.0", x, 1e-5)


def transformer_encoder_layer(sd: SD, key: str, x: Tensor, nhead: int) -> Tensor:
    """nn.TransformerEncoderLayer defaults (post-norm, ReLU, ffn 2048), eval mode; x [T,B,E]."""
    t, b, e = x.shape
    hd = e // nhead
    qkv = F.linear(x, sd[key + ".self_attn.in_proj_weight"], sd[key + ".self_attn.in_proj_bias"])
    q, k, v = qkv.chunk(3, -1)
    sh = lambda z: z.reshape(t, b * nhead, hd).transpose(0, 1)          # [B*h, T, hd]
    q, k, v = sh(q), sh(k), sh(v)
    att = torch.softmax(torch.bmm(q, k.transpose(1, 2)) / math.sqrt(hd), -1)
    o = torch.bmm(att, v).transpose(0, 1).reshape(t, b, e)
    o = F.linear(o, sd[key + ".self_attn.out_proj.weight"], sd[key + ".self_attn.out_proj.bias"])
    x = F.layer_norm(x + o, (e,), sd[key + ".norm1.weight"], sd[key + ".norm1.bias"])
    f = linear(sd, key + ".linear2", F.relu(linear(sd, key + ".linear1", x)))
    return F.layer_norm(x + f, (e,), sd[key + ".norm2.weight"], sd[key + ".norm2.bias"])


def transformer(sd: SD, key: str, x: Tensor, layers: int = 5, nhead: int = 5) -> Tensor:
    """MultiBench Transformer(n_features, dim): Conv1d(k=1, no bias) -> 5 encoder layers -> last step.
    x [B,T,F] -> [B,dim]."""
    y = F.conv1d(x.permute(0, 2, 1), sd[key + ".conv.weight"])          # [B,dim,T]
    y = y.permute(2, 0, 1)                                              # [T,B,dim]
    for i in range(layers):
        y = transformer_encoder_layer(sd, f"{key}.transformer.layers.{i}", y, nhead)
    return y[-1]


# ------------------------------------------------------------------ DynMM forward passes

def imdb_forward(sd: SD, inputs: Sequence[Tensor], temp: float = 1.0, hard_gate: bool = True, infer_mode: int = 0):
    """DynMMNet.forward, imdb_dyn.py:89-101.  inputs = [text [B,300], image [B,4096]]."""
    x = torch.cat(list(inputs), 1)
    weight = diff_softmax(mlp(sd, "gate", x), temp, hard_gate)
    p0 = mlp(sd, "text_head", mlp(sd, "text_encoder", inputs[0]))
    e0 = maxout_mlp(sd, "branch3.encoders.0", inputs[0], 512, 512)
    e1 = maxout_mlp(sd, "branch3.encoders.1", inputs[1], 1024, 512)
    p1 = linear(sd, "branch3.head.fc", torch.cat([e0.flatten(1), e1.flatten(1)], 1))
    if infer_mode > 0:
        return (p0, p1)[infer_mode - 1], 0
    out = weight[:, 0:1] * p0 + weight[:, 1:2] * p1
    return out, weight[:, 1].mean(), weight


def mosei_forward(sd: SD, inputs, temp: float = 1.0, hard_gate: bool = False, infer_mode: int = 0):
    """DynMMNetV2.forward, affect_dyn.py:152-165.  inputs = [[vis [B,T,35], aud [B,T,74], txt [B,T,300]], lens]."""
    feats = inputs[0]
    x = torch.cat(list(feats), 2)
    weight = diff_softmax(linear(sd, "gate.1", transformer(sd, "gate.0", x)), temp, hard_gate)
    p0 = mlp(sd, "text_head", transformer(sd, "text_encoder", feats[2]))
    reps = [transformer(sd, f"branch2.encoders.{i}", feats[i]) for i in range(3)]
    p1 = mlp(sd, "branch2.head", torch.cat([r.flatten(1) for r in reps], 1))
    if infer_mode > 0:
        return (p0, p1)[infer_mode - 1], 0
    if infer_mode == -1:
        weight = torch.ones_like(weight) / 2
    out = weight[:, 0:1] * p0 + weight[:, 1:2] * p1
    return out, weight[:, 1].mean(), weight


# ------------------------------------------------------------------ architecture pins: thop-style MAC counts

def _transformer_macs(n_feat: int, dim: int, t: int, layers: int = 5, ffn: int = 2048) -> int:
    """thop counts Conv1d and the two FFN Linears; it has no rule for nn.MultiheadAttention
    internals or LayerNorm, which is how the reference's constants come out."""
    return n_feat * dim * t + layers * (2 * dim * ffn * t)


def imdb_mmacs():
    """-> (E1 + gate, E2 + gate) in MMACs per sample; reference: imdb_dyn.py:66."""
    gate = 4396 * 128 + 128 * 2
    e1 = (300 * 512 + 512 * 512) + (512 * 512 + 512 * 23)
    bn = 2 * (300 + 512 + 512 + 4096 + 1024 + 512)      # thop: two ops per BatchNorm1d feature
    e2 = 300 * 1024 + 512 * 1024 + 4096 * 2048 + 1024 * 1024 + 1024 * 23 + bn
    return (e1 + gate) / 1e6, (e2 + gate) / 1e6


def mosei_mmacs(t: int = 50):
    """-> (E1 + gate, E2 + gate) in MMACs per sample at sequence length t; reference: affect_dyn.py:126."""
    gate = _transformer_macs(409, 10, t) + 10 * 2
    e1 = _transformer_macs(300, 120, t) + 120 * 64 + 64
    e2 = _transformer_macs(35, 60, t) + _transformer_macs(74, 120, t) + _transformer_macs(300, 120, t) + 300 * 128 + 128
    return (e1 + gate) / 1e6, (e2 + gate) / 1e6
