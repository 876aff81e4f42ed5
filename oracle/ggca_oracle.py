"""CPU oracle for the `cvit_GGCA_ADD_DEConv_RepBn8` CViT variant — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Restates /root/reference/CViT-main/model/cvit_GGCA_ADD_DEConv_RepBn8.py in fp32 torch functional ops:

* ``deconv_weight``   DEConv (:329-351) folded to ONE 3x3 kernel, exactly as its forward does: central-difference
                      (:218-235, centre tap minus the tap sum), horizontal / vertical difference from Conv1d weights
                      (:290-326), angular difference (:238-255, theta = 1, permutation [3,0,1,6,4,2,7,8,5]) and a
                      plain 3x3 conv; biases add.
* ``features``        features1 + features2 (:361-423): the CViT conv plan with DEConv in place of most 3x3 convs and a
                      BN-less, activation-less ``Conv2d(128,128)`` followed by ``DEConv(128)`` + ReLU in stage 3 (:385-388).
* ``ggca``            GGCA(512, 7, 7) (:143-213) and the extra ``x = x * ggca(x)`` of forward (:447-448).
* ``transformer``     attention branch pre-normed by nn.LayerNorm (eps 1e-5, PreNorm2 :62-71); the MLP branch by
                      LinearNorm, which in eval() is its ``norm1`` = LayerNorm(eps 1e-6) (:22-47, :50-60).
* head / slots        as cvit_oracle (same modules).

Pinned by tests/golden/ggca_*.npz: outputs of the reference class, which needs three shims to import and run on a
CPU-only container (its DEConv allocates with ``torch.cuda.FloatTensor`` and calls ``.cuda()`` in a constructor, and
it imports ``torchsummary``) — see oracle/make_golden.py:main_ggca.
"""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn.functional as F

from . import cvit_oracle as C

BN_EPS = 1e-5
# (sequential, conv index, kind, bn index or None, relu, pool)
PLAN = (
    ("features1", 0, "conv", 1, True, False), ("features1", 3, "de", 4, True, False), ("features1", 6, "de", 7, True, True),
    ("features1", 10, "conv", 11, True, False), ("features1", 13, "de", 14, True, False), ("features1", 16, "de", 17, True, True),
    ("features1", 20, "conv", 21, True, False), ("features1", 23, "de", 24, True, False),
    ("features1", 26, "conv", None, False, False), ("features1", 27, "de", None, True, True),
    ("features1", 30, "conv", 31, True, False), ("features1", 33, "de", 34, True, False), ("features1", 36, "de", 37, True, False),
    ("features1", 39, "de", 40, True, True),
    ("features2", 0, "conv", 1, True, False), ("features2", 3, "de", 4, True, False), ("features2", 6, "de", 7, True, False),
    ("features2", 9, "de", 10, True, True),
)
AD_PERM = [3, 0, 1, 6, 4, 2, 7, 8, 5]


def deconv_weight(sd: Dict[str, torch.Tensor], p: str):
    """Folded 3x3 kernel and bias of the DEConv at key prefix ``p`` (DEConv.forward :337-351)."""
    w1 = sd[p + ".conv1_1.conv.weight"]
    o, i = w1.shape[:2]
    f1 = w1.reshape(o, i, 9).clone()
    f1[:, :, 4] = f1[:, :, 4] - w1.reshape(o, i, 9).sum(2)
    w2 = sd[p + ".conv1_2.conv.weight"]                    # Conv1d [o, i, 3]
    f2 = torch.zeros(o, i, 9)
    f2[:, :, [0, 3, 6]] = w2
    f2[:, :, [2, 5, 8]] = -w2
    w3 = sd[p + ".conv1_3.conv.weight"]
    f3 = torch.zeros(o, i, 9)
    f3[:, :, [0, 1, 2]] = w3
    f3[:, :, [6, 7, 8]] = -w3
    w4 = sd[p + ".conv1_4.conv.weight"].reshape(o, i, 9)
    f4 = w4 - 1.0 * w4[:, :, AD_PERM]
    f5 = sd[p + ".conv1_5.weight"].reshape(o, i, 9)
    w = (f1 + f2 + f3 + f4 + f5).reshape(o, i, 3, 3)
    b = (sd[p + ".conv1_1.conv.bias"] + sd[p + ".conv1_2.conv.bias"] + sd[p + ".conv1_3.conv.bias"]
         + sd[p + ".conv1_4.conv.bias"] + sd[p + ".conv1_5.bias"])
    return w, b


def feature_layer(x, sd, layer: int):
    seq, ci, kind, bi, relu, pool = PLAN[layer]
    p = f"{seq}.{ci}"
    if kind == "de":
        w, b = deconv_weight(sd, p)
    else:
        w, b = sd[p + ".weight"], sd[p + ".bias"]
    y = F.conv2d(x, w, b, stride=1, padding=1)
    if bi is not None:
        q = f"{seq}.{bi}"
        y = F.batch_norm(y, sd[q + ".running_mean"], sd[q + ".running_var"], sd[q + ".weight"], sd[q + ".bias"], False, 0.0, BN_EPS)
    if relu:
        y = F.relu(y)
    if pool:
        y = F.max_pool2d(y, kernel_size=2, stride=2)
    return y


def features(x, sd, upto: int = len(PLAN)):
    for layer in range(upto):
        x = feature_layer(x, sd, layer)
    return x


def ggca(x, sd, groups: int = 4):
    """GGCA.forward (:172-213) on [b,512,7,7]; returns x * att_h * att_w."""
    b, c, hh, ww = x.shape
    gc = c // groups
    xg = x.reshape(b * groups, gc, hh, ww)

    def shared(v):
        y = F.conv2d(v, sd["ggca.shared_conv.0.weight"], sd["ggca.shared_conv.0.bias"])
        y = F.batch_norm(y, sd["ggca.shared_conv.1.running_mean"], sd["ggca.shared_conv.1.running_var"],
                         sd["ggca.shared_conv.1.weight"], sd["ggca.shared_conv.1.bias"], False, 0.0, BN_EPS)
        return F.conv2d(F.relu(y), sd["ggca.shared_conv.3.weight"], sd["ggca.shared_conv.3.bias"])

    h_avg, h_max = xg.mean(3, keepdim=True), xg.amax(3, keepdim=True)       # adaptive pools to (7,1) on a 7x7 map
    w_avg, w_max = xg.mean(2, keepdim=True), xg.amax(2, keepdim=True)
    att_h = torch.sigmoid(shared(h_avg) + shared(h_max))
    att_w = torch.sigmoid(shared(w_avg) + shared(w_max))
    return (xg * att_h * att_w).reshape(b, c, hh, ww)


def gated_features(x, sd):
    f = features(x, sd)
    return f * ggca(f, sd)                                   # forward :447-448


def transformer(x, sd):
    dim = x.shape[-1]
    for layer in range(C.DEPTH):
        p = f"transformer.layers.{layer}"
        y = F.layer_norm(x, (dim,), sd[f"{p}.0.fn.norm.weight"], sd[f"{p}.0.fn.norm.bias"], 1e-5)
        x = C.attention(y, sd, f"{p}.0.fn.fn") + x
        y = F.layer_norm(x, (dim,), sd[f"{p}.1.fn.norm.norm1.weight"], sd[f"{p}.1.fn.norm.norm1.bias"], 1e-6)
        y = F.linear(y, sd[f"{p}.1.fn.fn.net.0.weight"], sd[f"{p}.1.fn.fn.net.0.bias"])
        y = F.linear(F.gelu(y), sd[f"{p}.1.fn.fn.net.2.weight"], sd[f"{p}.1.fn.fn.net.2.bias"])
        x = y + x
    return x


def forward_slots(x, sd, slots):
    with torch.no_grad():
        t = C.embed_tokens(gated_features(x, sd), sd, slots)
        return C.head(transformer(t, sd), sd)


def forward(x, sd):
    b = x.shape[0]
    if b > 32:
        raise RuntimeError("CViT.forward: batch > 32 cannot broadcast against pos_embedding[0:32]")
    return forward_slots(x, sd, torch.arange(b))
