"""CPU oracle for ResVitKan inference — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Restates /root/reference/CViT-main/ResVitKan/ResVitKan.py:284-329 (`CViT.forward`: ResNet-50 features,
patch embedding, 6-layer ViT, `kan_head`) and kan.py:90-206 (`KANLinear.b_splines` / `forward`) with torch
functional ops in fp32.  Non-standard details kept on purpose: the Bottleneck applies ReLU after bn3 AND after the
residual add (ResVitKan.py:169-176); `features.channel` + `bn2` has no ReLU (:238-239); the head is
Linear -> Dropout(eval: identity) -> ReLU -> KAN([2048, 64, 2]) (:302-307,329); B-spline intervals are half-open
`x >= g[i] & x < g[i+1]` (kan.py:115).  Pinned by tests/golden/resvitkan_*.npz (outputs of the reference class).
"""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn.functional as F

from . import cvit_oracle as C

BN_EPS = 1e-5
LAYERS = ((64, 3, 1), (128, 4, 2), (256, 6, 2), (512, 3, 2))


def _bn(x, sd, name):
    return F.batch_norm(x, sd[name + ".running_mean"], sd[name + ".running_var"], sd[name + ".weight"], sd[name + ".bias"],
                        False, 0.0, BN_EPS)


def bottleneck(x, sd, p, stride, has_down):
    """ResVitKan.py:150-177 (note the extra ReLU before the residual add)."""
    out = F.relu(_bn(F.conv2d(x, sd[p + ".conv1.weight"]), sd, p + ".bn1"))
    out = F.relu(_bn(F.conv2d(out, sd[p + ".conv2.weight"], stride=stride, padding=1), sd, p + ".bn2"))
    out = F.relu(_bn(F.conv2d(out, sd[p + ".conv3.weight"]), sd, p + ".bn3"))
    res = x
    if has_down:
        res = _bn(F.conv2d(x, sd[p + ".downsample.0.weight"], stride=stride), sd, p + ".downsample.1")
    return F.relu(out + res)


def stem(x, sd):
    x = F.relu(_bn(F.conv2d(x, sd["features.conv1.weight"], stride=2, padding=3), sd, "features.bn1"))
    return F.max_pool2d(x, kernel_size=3, stride=2, padding=1)


def features(x, sd, upto_layer: int = 4, taps: Dict[str, torch.Tensor] = None):
    """ResNet.forward (ResVitKan.py:228-240); `taps` (optional dict) receives intermediate activations."""
    x = stem(x, sd)
    if taps is not None:
        taps["stem"] = x
    for li, (planes, blocks, stride) in enumerate(LAYERS, start=1):
        if li > upto_layer:
            return x
        for b in range(blocks):
            x = bottleneck(x, sd, f"features.layer{li}.{b}", stride if b == 0 else 1, b == 0)
        if taps is not None:
            taps[f"layer{li}"] = x
    x = _bn(F.conv2d(x, sd["features.channel.weight"]), sd, "features.bn2")
    if taps is not None:
        taps["channel"] = x
    return x


def b_splines(x, grid, spline_order: int = 3):
    """kan.py:90-132: x [n, in], grid [in, 12] -> bases [n, in, 8]."""
    x = x.unsqueeze(-1)
    bases = ((x >= grid[:, :-1]) & (x < grid[:, 1:])).to(x.dtype)
    for k in range(1, spline_order + 1):
        bases = ((x - grid[:, : -(k + 1)]) / (grid[:, k:-1] - grid[:, : -(k + 1)]) * bases[:, :, :-1]) + (
            (grid[:, k + 1:] - x) / (grid[:, k + 1:] - grid[:, 1:(-k)]) * bases[:, :, 1:])
    return bases.contiguous()


def kan_linear(x, sd, q):
    """kan.py:189-206: silu(x) W_base^T + Bspline(x) (W_spline * scaler)^T."""
    base = F.linear(F.silu(x), sd[q + ".base_weight"])
    w = sd[q + ".spline_weight"] * sd[q + ".spline_scaler"].unsqueeze(-1)
    spline = F.linear(b_splines(x, sd[q + ".grid"]).view(x.size(0), -1), w.view(w.size(0), -1))
    return base + spline


def kan_head(c, sd):
    h = F.relu(F.linear(c, sd["kan_head.0.weight"], sd["kan_head.0.bias"]))      # Dropout is identity in eval()
    h = kan_linear(h, sd, "kan_head.3.layers.0")
    return kan_linear(h, sd, "kan_head.3.layers.1")


def forward_slots(x, sd, slots):
    """fp32 NCHW [n,3,224,224] + slot[n] -> logits [n,2]."""
    with torch.no_grad():
        f = features(x, sd)
        t = C.embed_tokens(f, sd, slots)
        t = C.transformer(t, sd)
        return kan_head(t[:, 0], sd)


def forward(x, sd):
    b = x.shape[0]
    if b > 32:
        raise RuntimeError("ResVitKan CViT.forward: batch > 32 cannot broadcast against pos_embedding[0:32]")
    return forward_slots(x, sd, torch.arange(b))
