"""CPU oracle for the S3D clip classifier — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Restates /root/reference/sx_exp_deepfakedetect-master/S3D/model.py in fp32 torch functional ops (SRM_net == 'no', and
SRM_net == 'yes' = ``hpf`` below in front of ``features``):

* ``hpf``                         SRM/HPF.py:11-37  Conv3d(3, 30, (1,5,5), padding (0,2,2), bias=False) with `SRM.hpf.weight`
* ``basic_conv`` / ``sep_conv``   :50-82   Conv3d(bias=False) + BatchNorm3d(eps 1e-3, eval) + ReLU; separable = (1,k,k) then (k,1,1)
* ``mixed``                       :84-342  four branches (1x1x1 | 1x1x1+sep3 | 1x1x1+sep3 | MaxPool3d(3,1,1)+1x1x1), channel concat
* ``features``                    :17-34   stem sep7/2, pools, 9 Mixed blocks
* ``forward``                     :37-48   avg_pool3d((2,H,W), stride 1), 1x1x1 conv fc (+bias), mean over time -> logits [b, classes]
* ``video_score``                 S3D-test.py:269-279: sigmoid of each clip logit, mean

Input is the raw 0..255 BGR float clip [b,3,T,224,224] (S3D-test.py:94-96: no normalisation).
Pinned by tests/golden/s3d_*.npz (outputs of the reference class, oracle/make_golden.py:main_s3d).
"""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn.functional as F

BN_EPS = 1e-3
# name -> (cin, branch0, (branch1 mid, out), (branch2 mid, out), branch3)   model.py:84-342
MIXED = {
    "3b": (192, 64, (96, 128), (16, 32), 32), "3c": (256, 128, (128, 192), (32, 96), 64),
    "4b": (480, 192, (96, 208), (16, 48), 64), "4c": (512, 160, (112, 224), (24, 64), 64),
    "4d": (512, 128, (128, 256), (24, 64), 64), "4e": (512, 112, (144, 288), (32, 64), 64),
    "4f": (528, 256, (160, 320), (32, 128), 128), "5b": (832, 256, (160, 320), (32, 128), 128),
    "5c": (832, 384, (192, 384), (48, 128), 128),
}
# base.N -> what it is (model.py:17-34)
BASE = ((0, "sep", (3, 64, 7, 2, 3)), (1, "pool", ((1, 3, 3), (1, 2, 2), (0, 1, 1))), (2, "basic", (64, 64)),
        (3, "sep", (64, 192, 3, 1, 1)), (4, "pool", ((1, 3, 3), (1, 2, 2), (0, 1, 1))), (5, "mixed", "3b"), (6, "mixed", "3c"),
        (7, "pool", ((3, 3, 3), (2, 2, 2), (1, 1, 1))), (8, "mixed", "4b"), (9, "mixed", "4c"), (10, "mixed", "4d"),
        (11, "mixed", "4e"), (12, "mixed", "4f"), (13, "pool", ((2, 2, 2), (2, 2, 2), (0, 0, 0))), (14, "mixed", "5b"),
        (15, "mixed", "5c"))


def _bn_relu(x, sd, p):
    x = F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"], False, 0.0, BN_EPS)
    return F.relu(x)


def basic_conv(x, sd: Dict[str, torch.Tensor], p: str):
    return _bn_relu(F.conv3d(x, sd[p + ".conv.weight"]), sd, p + ".bn")


def sep_conv(x, sd, p: str, stride: int, pad: int):
    x = _bn_relu(F.conv3d(x, sd[p + ".conv_s.weight"], stride=(1, stride, stride), padding=(0, pad, pad)), sd, p + ".bn_s")
    return _bn_relu(F.conv3d(x, sd[p + ".conv_t.weight"], stride=(stride, 1, 1), padding=(pad, 0, 0)), sd, p + ".bn_t")


def mixed(x, sd, p: str):
    x0 = basic_conv(x, sd, p + ".branch0.0")
    x1 = sep_conv(basic_conv(x, sd, p + ".branch1.0"), sd, p + ".branch1.1", 1, 1)
    x2 = sep_conv(basic_conv(x, sd, p + ".branch2.0"), sd, p + ".branch2.1", 1, 1)
    x3 = basic_conv(F.max_pool3d(x, kernel_size=3, stride=1, padding=1), sd, p + ".branch3.1")
    return torch.cat((x0, x1, x2, x3), 1)


def features(x, sd, taps=None, upto: int = 16):
    for idx, kind, arg in BASE:
        if idx >= upto:
            break
        p = f"base.{idx}"
        if kind == "sep":
            x = sep_conv(x, sd, p, arg[3], arg[4])
        elif kind == "basic":
            x = basic_conv(x, sd, p)
        elif kind == "pool":
            x = F.max_pool3d(x, kernel_size=arg[0], stride=arg[1], padding=arg[2])
        else:
            x = mixed(x, sd, p)
        if taps is not None:
            taps[idx] = x
    return x


def hpf(x, sd):
    """SRM/HPF.py:30-35 (model.py:38-39): the 30 high-pass residual maps of every frame."""
    return F.conv3d(x, sd["SRM.hpf.weight"], padding=(0, 2, 2))


def forward(x, sd, taps=None, srm: bool = False):
    """Raw clip [b,3,T,224,224] (0..255 BGR floats) -> logits [b, classes].  srm=True: S3D(num_class, 'yes')."""
    with torch.no_grad():
        if srm:
            x = hpf(x, sd)
            if taps is not None:
                taps["hpf"] = x
        y = features(x, sd, taps)
        y = F.avg_pool3d(y, (2, y.size(3), y.size(4)), stride=1)
        y = F.conv3d(y, sd["fc.0.weight"], sd["fc.0.bias"])
        y = y.view(y.size(0), y.size(1), y.size(2))
        return y.mean(2)


def video_score(clip_logits: torch.Tensor) -> float:
    """S3D-test.py:269-279: mean over a video's clips of sigmoid(logit)."""
    return torch.sigmoid(clip_logits.double().flatten()).mean().item()
