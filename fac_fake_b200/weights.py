"""Deterministic synthetic CViT parameter sets (no reference import needed).

The key names and shapes are those of the reference ``CViT.state_dict()``
(/root/reference/CViT-main/model/cvit.py:80-165).  The values
follow the distributions of the default torch initialisers the reference relies
on (kaiming-uniform for conv / linear, ``randn`` for ``pos_embedding`` /
``cls_token``) but are drawn from our own seeded ``torch.Generator`` so that the
same tensors can be regenerated bit-for-bit on the GPU box, where
/root/reference does not exist.

``variant="bn"`` additionally randomises the BatchNorm / LayerNorm affine parameters
and BN running statistics (default init makes eval-mode BN ≈ identity and would hide
BN-folding bugs) and uses a variance-preserving (He) gain for
the conv weights so that the conv features are O(1) at the patch embedding.
"""
from __future__ import annotations

import math
from collections import OrderedDict

import torch

# (conv index in features, bn index, cin, cout); pool follows convs 3,6,9,13,17
CONV_PLAN = [
    (0, 1, 3, 32), (3, 4, 32, 32), (6, 7, 32, 32),
    (10, 11, 32, 64), (13, 14, 64, 64), (16, 17, 64, 64),
    (20, 21, 64, 128), (23, 24, 128, 128), (26, 27, 128, 128),
    (30, 31, 128, 256), (33, 34, 256, 256), (36, 37, 256, 256), (39, 40, 256, 256),
    (43, 44, 256, 512), (46, 47, 512, 512), (49, 50, 512, 512), (52, 53, 512, 512),
]
POOL_AFTER = (2, 5, 8, 12, 16)          # 0-based conv layer indices followed by MaxPool2d(2)
DIM, DEPTH, HEADS, MLP_DIM, PATCH_DIM, NUM_CLASSES, MAX_SLOTS = 1024, 6, 8, 2048, 25088, 2, 32
BN_EPS = 1e-5
LN_EPS = 1e-5


def _uniform(gen, shape, bound):
    return (torch.rand(shape, generator=gen, dtype=torch.float32) * 2.0 - 1.0) * bound


def _linear(gen, sd, name, out_f, in_f, bias=True):
    bound = 1.0 / math.sqrt(in_f)           # kaiming_uniform(a=sqrt(5)) == U(-1/sqrt(fan_in), +)
    sd[name + ".weight"] = _uniform(gen, (out_f, in_f), bound)
    if bias:
        sd[name + ".bias"] = _uniform(gen, (out_f,), bound)


def state_dict_keys():
    return list(make_state_dict(0, variant="shape_only").keys())


def make_state_dict(seed: int = 0, variant: str = "default") -> "OrderedDict[str, torch.Tensor]":
    """Return a CViT state_dict. variant: "default" | "bn" | "shape_only"."""
    gen = torch.Generator(device="cpu")
    gen.manual_seed(1000003 * seed + 17)
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    sd["pos_embedding"] = torch.randn((MAX_SLOTS, 1, DIM), generator=gen)
    sd["cls_token"] = torch.randn((1, 1, DIM), generator=gen)
    for ci, bi, cin, cout in CONV_PLAN:
        bound = 1.0 / math.sqrt(cin * 9)
        # "bn": He-uniform gain so activations stay O(1) through the 17 layers and the
        # conv features actually reach the logits (default init shrinks them to ~1e-7).
        wbound = math.sqrt(6.0 / (cin * 9)) if variant == "bn" else bound
        sd[f"features.{ci}.weight"] = _uniform(gen, (cout, cin, 3, 3), wbound)
        sd[f"features.{ci}.bias"] = _uniform(gen, (cout,), bound)
        if variant == "bn":
            sd[f"features.{bi}.weight"] = torch.rand((cout,), generator=gen) + 0.5
            sd[f"features.{bi}.bias"] = torch.randn((cout,), generator=gen) * 0.1
            sd[f"features.{bi}.running_mean"] = torch.randn((cout,), generator=gen) * 0.1
            sd[f"features.{bi}.running_var"] = torch.rand((cout,), generator=gen) + 0.5
        else:
            sd[f"features.{bi}.weight"] = torch.ones(cout)
            sd[f"features.{bi}.bias"] = torch.zeros(cout)
            sd[f"features.{bi}.running_mean"] = torch.zeros(cout)
            sd[f"features.{bi}.running_var"] = torch.ones(cout)
        sd[f"features.{bi}.num_batches_tracked"] = torch.zeros((), dtype=torch.long)
    _linear(gen, sd, "patch_to_embedding", DIM, PATCH_DIM)
    for layer in range(DEPTH):
        p = f"transformer.layers.{layer}"
        for blk in (0, 1):
            if variant == "bn":
                sd[f"{p}.{blk}.fn.norm.weight"] = torch.rand((DIM,), generator=gen) + 0.5
                sd[f"{p}.{blk}.fn.norm.bias"] = torch.randn((DIM,), generator=gen) * 0.1
            else:
                sd[f"{p}.{blk}.fn.norm.weight"] = torch.ones(DIM)
                sd[f"{p}.{blk}.fn.norm.bias"] = torch.zeros(DIM)
            if blk == 0:
                _linear(gen, sd, f"{p}.0.fn.fn.to_qkv", 3 * DIM, DIM, bias=False)
                _linear(gen, sd, f"{p}.0.fn.fn.to_out", DIM, DIM)
            else:
                _linear(gen, sd, f"{p}.1.fn.fn.net.0", MLP_DIM, DIM)
                _linear(gen, sd, f"{p}.1.fn.fn.net.2", DIM, MLP_DIM)
    _linear(gen, sd, "mlp_head.0", MLP_DIM, DIM)
    _linear(gen, sd, "mlp_head.2", NUM_CLASSES, MLP_DIM)
    # reorder to the reference's state_dict order is not required: load_state_dict is key based.
    return sd


def synthetic_crops(n: int, seed: int = 0) -> torch.Tensor:
    """uint8 [n,224,224,3] uniform crops."""
    gen = torch.Generator(device="cpu")
    gen.manual_seed(7919 * seed + 3)
    return torch.randint(0, 256, (n, 224, 224, 3), generator=gen, dtype=torch.uint8)


# ----------------------------------------------------------------------------------------------------------------
# ResVitKan (/root/reference/CViT-main/ResVitKan/ResVitKan.py:284-329, kan.py:18-206)
RESNET_LAYERS = ((64, 3, 1), (128, 4, 2), (256, 6, 2), (512, 3, 2))     # (planes, blocks, stride of the first block)


def _bn(gen, sd, name, c, variant):
    if variant == "bn":
        sd[name + ".weight"] = torch.rand((c,), generator=gen) * 0.5 + 0.5
        sd[name + ".bias"] = torch.randn((c,), generator=gen) * 0.1
        sd[name + ".running_mean"] = torch.randn((c,), generator=gen) * 0.1
        sd[name + ".running_var"] = torch.rand((c,), generator=gen) + 0.5
    else:
        sd[name + ".weight"] = torch.ones(c)
        sd[name + ".bias"] = torch.zeros(c)
        sd[name + ".running_mean"] = torch.zeros(c)
        sd[name + ".running_var"] = torch.ones(c)
    sd[name + ".num_batches_tracked"] = torch.zeros((), dtype=torch.long)


def _conv_normal(gen, shape):
    # ResNet.__init__ (ResVitKan.py:202-209): normal(0, sqrt(2 / (k*k*out_channels)))
    out_c, _, kh, kw = shape
    return torch.randn(shape, generator=gen) * math.sqrt(2.0 / (kh * kw * out_c))


def make_resvitkan_state_dict(seed: int = 0, variant: str = "default") -> "OrderedDict[str, torch.Tensor]":
    """state_dict of the reference ResVitKan `CViT` (ResNet-50 features + ViT + KAN head), key names as in the reference."""
    gen = torch.Generator(device="cpu")
    gen.manual_seed(7000003 * seed + 29)
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    sd["pos_embedding"] = torch.randn((MAX_SLOTS, 1, DIM), generator=gen)
    sd["cls_token"] = torch.randn((1, 1, DIM), generator=gen)
    sd["features.conv1.weight"] = _conv_normal(gen, (64, 3, 7, 7))
    _bn(gen, sd, "features.bn1", 64, variant)
    inplanes = 64
    for li, (planes, blocks, stride) in enumerate(RESNET_LAYERS, start=1):
        for b in range(blocks):
            p = f"features.layer{li}.{b}"
            sd[p + ".conv1.weight"] = _conv_normal(gen, (planes, inplanes, 1, 1))
            _bn(gen, sd, p + ".bn1", planes, variant)
            sd[p + ".conv2.weight"] = _conv_normal(gen, (planes, planes, 3, 3))
            _bn(gen, sd, p + ".bn2", planes, variant)
            sd[p + ".conv3.weight"] = _conv_normal(gen, (planes * 4, planes, 1, 1))
            _bn(gen, sd, p + ".bn3", planes * 4, variant)
            if b == 0:
                sd[p + ".downsample.0.weight"] = _conv_normal(gen, (planes * 4, inplanes, 1, 1))
                _bn(gen, sd, p + ".downsample.1", planes * 4, variant)
            inplanes = planes * 4
    sd["features.channel.weight"] = _conv_normal(gen, (512, 2048, 1, 1))
    _bn(gen, sd, "features.bn2", 512, variant)
    _linear(gen, sd, "patch_to_embedding", DIM, PATCH_DIM)
    for layer in range(DEPTH):
        p = f"transformer.layers.{layer}"
        for blk in (0, 1):
            if variant == "bn":
                sd[f"{p}.{blk}.fn.norm.weight"] = torch.rand((DIM,), generator=gen) + 0.5
                sd[f"{p}.{blk}.fn.norm.bias"] = torch.randn((DIM,), generator=gen) * 0.1
            else:
                sd[f"{p}.{blk}.fn.norm.weight"] = torch.ones(DIM)
                sd[f"{p}.{blk}.fn.norm.bias"] = torch.zeros(DIM)
            if blk == 0:
                _linear(gen, sd, f"{p}.0.fn.fn.to_qkv", 3 * DIM, DIM, bias=False)
                _linear(gen, sd, f"{p}.0.fn.fn.to_out", DIM, DIM)
            else:
                _linear(gen, sd, f"{p}.1.fn.fn.net.0", MLP_DIM, DIM)
                _linear(gen, sd, f"{p}.1.fn.fn.net.2", DIM, MLP_DIM)
    _linear(gen, sd, "kan_head.0", MLP_DIM, DIM)
    for i, (fin, fout) in enumerate(((MLP_DIM, 64), (64, NUM_CLASSES))):
        q = f"kan_head.3.layers.{i}"
        b = 1.0 / math.sqrt(fin)
        sd[q + ".base_weight"] = _uniform(gen, (fout, fin), b)
        sd[q + ".spline_weight"] = torch.randn((fout, fin, 8), generator=gen) * 0.1
        sd[q + ".spline_scaler"] = _uniform(gen, (fout, fin), b)
        h = 2.0 / 5
        sd[q + ".grid"] = (torch.arange(-3, 5 + 3 + 1) * h - 1.0).expand(fin, -1).contiguous()
    # mlp_head exists in the reference module but is not used by forward(); keep the keys so load_state_dict(strict) works
    _linear(gen, sd, "mlp_head.0", MLP_DIM, DIM)
    _linear(gen, sd, "mlp_head.3", NUM_CLASSES, MLP_DIM)
    return sd
