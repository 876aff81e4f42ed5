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
    """Return a CViT state_dict. variant: "default" | "bn" | "decisive" | "shape_only".

    "decisive" = "bn" re-balanced so that REAL/FAKE decisions mean something on synthetic inputs.  With random weights
    the logits are dominated by pos_embedding[slot] (sigma = 1) while the conv features barely move them, so every
    30-frame video scores 0.498 +- 0.002 and its decision is a coin flip of rounding noise.  Here pos_embedding is
    scaled by 0.02 (content decides), the last head layer by 60 and its bias re-centred on the mean logit of
    `synthetic_video_crops` inputs: per-video scores then spread over 0.15 .. 0.75 on both sides of the threshold."""
    if variant == "decisive":
        sd = make_state_dict(seed, "bn")
        sd["pos_embedding"] = sd["pos_embedding"] * 0.02
        centre = torch.tensor([-0.2253, 0.0540])          # mean logit of the bn variant with the scaled pos_embedding
        sd["mlp_head.2.weight"] = sd["mlp_head.2.weight"] * 60.0
        sd["mlp_head.2.bias"] = (sd["mlp_head.2.bias"] - centre) * 60.0
        return sd
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


def synthetic_video_crops(video_id: int, frames: int = 30) -> torch.Tensor:
    """uint8 [frames,224,224,3]: uniform noise (per-video seed = video id) with a per-video
    contrast and brightness, so that different videos have different content statistics (and different scores)."""
    gen = torch.Generator(device="cpu")
    gen.manual_seed(5000 + video_id)
    a = float(torch.rand((), generator=gen)) * 0.9 + 0.1
    b = float(torch.rand((), generator=gen)) * (255.0 * (1.0 - a))
    c = synthetic_crops(frames, seed=1000 + video_id).float() * a + b
    return c.round().clamp(0, 255).to(torch.uint8)


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


# ----------------------------------------------------------------------------------------------------------------
# cvit_GGCA_ADD_DEConv_RepBn8 (/root/reference/CViT-main/model/cvit_GGCA_ADD_DEConv_RepBn8.py:353-455)
# (sequential, conv index, kind, bn index or None, cin, cout)
GGCA_PLAN = (
    ("features1", 0, "conv", 1, 3, 32), ("features1", 3, "de", 4, 32, 32), ("features1", 6, "de", 7, 32, 32),
    ("features1", 10, "conv", 11, 32, 64), ("features1", 13, "de", 14, 64, 64), ("features1", 16, "de", 17, 64, 64),
    ("features1", 20, "conv", 21, 64, 128), ("features1", 23, "de", 24, 128, 128),
    ("features1", 26, "conv", None, 128, 128), ("features1", 27, "de", None, 128, 128),
    ("features1", 30, "conv", 31, 128, 256), ("features1", 33, "de", 34, 256, 256), ("features1", 36, "de", 37, 256, 256),
    ("features1", 39, "de", 40, 256, 256),
    ("features2", 0, "conv", 1, 256, 512), ("features2", 3, "de", 4, 512, 512), ("features2", 6, "de", 7, 512, 512),
    ("features2", 9, "de", 10, 512, 512),
)


def _deconv(gen, sd, p, dim, std):
    """The five branches of a DEConv (:329-335): 2-D convs for cd / ad / plain, Conv1d weights for hd / vd."""
    b = 1.0 / math.sqrt(dim * 9)
    for name, shape in (("conv1_1.conv", (dim, dim, 3, 3)), ("conv1_2.conv", (dim, dim, 3)), ("conv1_3.conv", (dim, dim, 3)),
                        ("conv1_4.conv", (dim, dim, 3, 3)), ("conv1_5", (dim, dim, 3, 3))):
        sd[f"{p}.{name}.weight"] = torch.randn(shape, generator=gen) * std
        sd[f"{p}.{name}.bias"] = _uniform(gen, (dim,), b * 0.2)


def make_ggca_state_dict(seed: int = 0, variant: str = "default") -> "OrderedDict[str, torch.Tensor]":
    """state_dict of the reference `cvit_GGCA_ADD_DEConv_RepBn8.CViT`, key names as in the reference."""
    gen = torch.Generator(device="cpu")
    gen.manual_seed(9000011 * seed + 41)
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    sd["pos_embedding"] = torch.randn((MAX_SLOTS, 1, DIM), generator=gen)
    sd["cls_token"] = torch.randn((1, 1, DIM), generator=gen)
    gain = 1.3 if variant == "bn" else 1.15      # "bn": the random BN scales (mean 0.75) shrink every layer
    for seq, ci, kind, bi, cin, cout in GGCA_PLAN:
        p = f"{seq}.{ci}"
        if kind == "de":
            # the folded kernel is a sum of five terms (6.0x the variance of one, measured); He gain x `gain` keeps the
            # activations O(1) through the 18 layers so that the conv features reach the logits
            _deconv(gen, sd, p, cout, gain * math.sqrt(2.0 / (9 * cin * 6.0)))
        else:
            sd[p + ".weight"] = torch.randn((cout, cin, 3, 3), generator=gen) * gain * math.sqrt((1.0 if bi is None else 2.0) / (9 * cin))
            sd[p + ".bias"] = _uniform(gen, (cout,), 0.2 / math.sqrt(cin * 9))
        if bi is not None:
            _bn(gen, sd, f"{seq}.{bi}", cout, variant)
    # GGCA(512, 7, 7): shared 1x1 convs 128 -> 8 -> 128 with a BN in between (:159-166)
    sd["ggca.shared_conv.0.weight"] = _uniform(gen, (8, 128, 1, 1), 1.0 / math.sqrt(128))
    sd["ggca.shared_conv.0.bias"] = _uniform(gen, (8,), 1.0 / math.sqrt(128))
    _bn(gen, sd, "ggca.shared_conv.1", 8, variant)
    sd["ggca.shared_conv.3.weight"] = _uniform(gen, (128, 8, 1, 1), 1.0 / math.sqrt(8))
    sd["ggca.shared_conv.3.bias"] = _uniform(gen, (128,), 1.0 / math.sqrt(8))
    _deconv(gen, sd, "Deconv", 256, 0.01)                    # constructed by the reference (:431) but unused in forward
    _linear(gen, sd, "patch_to_embedding", DIM, PATCH_DIM)
    for layer in range(DEPTH):
        p = f"transformer.layers.{layer}"
        rand = variant == "bn"
        sd[f"{p}.0.fn.norm.weight"] = torch.rand((DIM,), generator=gen) + 0.5 if rand else torch.ones(DIM)
        sd[f"{p}.0.fn.norm.bias"] = torch.randn((DIM,), generator=gen) * 0.1 if rand else torch.zeros(DIM)
        _linear(gen, sd, f"{p}.0.fn.fn.to_qkv", 3 * DIM, DIM, bias=False)
        _linear(gen, sd, f"{p}.0.fn.fn.to_out", DIM, DIM)
        q = f"{p}.1.fn.norm"                                 # LinearNorm (:22-47): eval() uses norm1 only
        sd[q + ".warm"] = torch.tensor(0)
        sd[q + ".iter"] = torch.tensor(300000)
        sd[q + ".total_step"] = torch.tensor(300000)
        sd[q + ".norm1.weight"] = torch.rand((DIM,), generator=gen) + 0.5 if rand else torch.ones(DIM)
        sd[q + ".norm1.bias"] = torch.randn((DIM,), generator=gen) * 0.1 if rand else torch.zeros(DIM)
        sd[q + ".norm2.alpha"] = torch.ones(1)
        _bn(gen, sd, q + ".norm2.bn", DIM, variant)
        _linear(gen, sd, f"{p}.1.fn.fn.net.0", MLP_DIM, DIM)
        _linear(gen, sd, f"{p}.1.fn.fn.net.2", DIM, MLP_DIM)
    _linear(gen, sd, "mlp_head.0", MLP_DIM, DIM)
    _linear(gen, sd, "mlp_head.2", NUM_CLASSES, MLP_DIM)
    return sd


# ----------------------------------------------------------------------------------------------------------------
# S3D (/root/reference/sx_exp_deepfakedetect-master/S3D/model.py)
S3D_MIXED = {
    "3b": (192, 64, (96, 128), (16, 32), 32), "3c": (256, 128, (128, 192), (32, 96), 64),
    "4b": (480, 192, (96, 208), (16, 48), 64), "4c": (512, 160, (112, 224), (24, 64), 64),
    "4d": (512, 128, (128, 256), (24, 64), 64), "4e": (512, 112, (144, 288), (32, 64), 64),
    "4f": (528, 256, (160, 320), (32, 128), 128), "5b": (832, 256, (160, 320), (32, 128), 128),
    "5c": (832, 384, (192, 384), (48, 128), 128),
}
S3D_BASE_MIXED = {5: "3b", 6: "3c", 8: "4b", 9: "4c", 10: "4d", 11: "4e", 12: "4f", 14: "5b", 15: "5c"}


def _bn3(gen, sd, name, c, variant):
    if variant == "bn":
        sd[name + ".weight"] = torch.rand((c,), generator=gen) * 0.5 + 0.75
        sd[name + ".bias"] = torch.randn((c,), generator=gen) * 0.1
        sd[name + ".running_mean"] = torch.randn((c,), generator=gen) * 0.1
        sd[name + ".running_var"] = torch.rand((c,), generator=gen) * 0.5 + 0.75
    else:
        sd[name + ".weight"] = torch.ones(c)
        sd[name + ".bias"] = torch.zeros(c)
        sd[name + ".running_mean"] = torch.zeros(c)
        sd[name + ".running_var"] = torch.ones(c)
    sd[name + ".num_batches_tracked"] = torch.zeros((), dtype=torch.long)


def _he3(gen, shape, gain=1.0):
    fan_in = shape[1] * shape[2] * shape[3] * shape[4]
    return torch.randn(shape, generator=gen) * gain * math.sqrt(2.0 / fan_in)


def _s3d_basic(gen, sd, p, cin, cout, variant, gain=1.0):
    sd[p + ".conv.weight"] = _he3(gen, (cout, cin, 1, 1, 1), gain)
    _bn3(gen, sd, p + ".bn", cout, variant)


def _s3d_sep(gen, sd, p, cin, cout, k, variant, gain=1.0):
    sd[p + ".conv_s.weight"] = _he3(gen, (cout, cin, 1, k, k), gain)
    _bn3(gen, sd, p + ".bn_s", cout, variant)
    sd[p + ".conv_t.weight"] = _he3(gen, (cout, cout, k, 1, 1))
    _bn3(gen, sd, p + ".bn_t", cout, variant)


def make_s3d_state_dict(seed: int = 0, variant: str = "default", num_class: int = 1, srm: bool = False) -> "OrderedDict[str, torch.Tensor]":
    """state_dict of the reference `S3D(num_class, 'no')` (model.py:6-48), key names as in the reference (incl. the
    always-constructed, unused-without-SRM `SRM.hpf.weight`).  ``srm=True``: `S3D(num_class, 'yes')` — the first
    convolution takes the 30 high-pass channels; `SRM.hpf.weight` is a set of zero-sum (high-pass) 5x5 kernels replicated
    over the three input channels / 3, the structure of the reference's SRM bank (SRM/HPF.py:17-28)."""
    gen = torch.Generator(device="cpu")
    gen.manual_seed(5000011 * seed + 53)
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    if srm:
        k = torch.randn((30, 1, 1, 5, 5), generator=gen)
        k = (k - k.mean(dim=(3, 4), keepdim=True)) * 0.25           # zero-sum residual filters
        sd["SRM.hpf.weight"] = (k / 3.0).repeat(1, 3, 1, 1, 1).contiguous()
        _s3d_sep(gen, sd, "base.0", 30, 64, 7, variant, gain=1.0 / 64.0)
    else:
        sd["SRM.hpf.weight"] = torch.randn((30, 3, 1, 5, 5), generator=gen) * 0.1
        # the input is raw 0..255 pixels: scale the first kernel so that activations are O(1) after the stem
        _s3d_sep(gen, sd, "base.0", 3, 64, 7, variant, gain=1.0 / 128.0)
    _s3d_basic(gen, sd, "base.2", 64, 64, variant)
    _s3d_sep(gen, sd, "base.3", 64, 192, 3, variant)
    for idx, name in S3D_BASE_MIXED.items():
        cin, b0, (m1, o1), (m2, o2), b3 = S3D_MIXED[name]
        p = f"base.{idx}"
        _s3d_basic(gen, sd, p + ".branch0.0", cin, b0, variant)
        _s3d_basic(gen, sd, p + ".branch1.0", cin, m1, variant)
        _s3d_sep(gen, sd, p + ".branch1.1", m1, o1, 3, variant)
        _s3d_basic(gen, sd, p + ".branch2.0", cin, m2, variant)
        _s3d_sep(gen, sd, p + ".branch2.1", m2, o2, 3, variant)
        _s3d_basic(gen, sd, p + ".branch3.1", cin, b3, variant)
    sd["fc.0.weight"] = torch.randn((num_class, 1024, 1, 1, 1), generator=gen) * (1.0 / 32.0)
    sd["fc.0.bias"] = torch.randn((num_class,), generator=gen) * 0.1
    return sd


def synthetic_clips(b: int, t: int, seed: int = 0, hw: int = 224) -> torch.Tensor:
    """uint8 [b,T,hw,hw,3] uniform clips (BGR, frame-major NHWC: the layout frames come out of the decoder in)."""
    gen = torch.Generator(device="cpu")
    gen.manual_seed(104729 * seed + 7)
    return torch.randint(0, 256, (b, t, hw, hw, 3), generator=gen, dtype=torch.uint8)
