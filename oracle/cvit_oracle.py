"""CPU oracle for the CViT hot path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module; the product path
(``fac_fake_b200``) never does and fails loudly when its CUDA library is missing.

It restates, op by op in fp32 on the CPU, what the reference computes on the path
(all line numbers are /root/reference/CViT-main/...):

* ``normalize_crops``         cvit_prediction.py:41-45,209-215  (to-tensor, permute, /255, Normalize)
* ``features`` / ``feature_layer``  model/cvit.py:86-148        (17 x conv3x3+BN(eval)+ReLU, 5 x MaxPool2d(2))
* ``embed_tokens``            model/cvit.py:170-175              (rearrange, patch_to_embedding, cls cat, += pos[0:b])
* ``transformer``             model/cvit.py:5-78                  (6 x pre-LN attention + GELU MLP, residual)
* ``head``                    model/cvit.py:177-179
* ``forward``                 model/cvit.py:167-179               (b <= 32, like the reference)
* ``forward_slots``           same, with an explicit batch-slot index per crop
* ``pred_sig`` / ``pre_process_prediction`` / ``video_score``   cvit_prediction.py:258-281
* ``predict_from_crops``      cvit_prediction.py:209-242          (chunks [0:32],[32:64],[64:90])
* ``real_or_fake``            cvit_prediction.py:284-292

Pinning: the reference ships no golden vectors usable without its (absent) trained
weights, so this restatement is pinned against OUTPUTS OF THE
REFERENCE ITSELF: ``oracle/make_golden.py`` imports ``cvit.CViT`` from
/root/reference in the build container, loads the same synthetic ``state_dict``
and stores its logits under ``tests/golden/``; ``tests/test_oracle.py`` checks this
module against those fixtures (and against the live class when /root/reference
exists).

``bf16_sim=True`` emulates the engine's precision contract (bf16 GEMM/conv operands,
fp32 accumulation, fp32 epilogues / residual stream) — used only to budget the
2e-2 tolerance, never as an expected value.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F

MEAN = (0.485, 0.456, 0.406)     # cvit_prediction.py:41
STD = (0.229, 0.224, 0.225)      # cvit_prediction.py:42

# features.N indices (model/cvit.py:86-148)
CONV_IDX = (0, 3, 6, 10, 13, 16, 20, 23, 26, 30, 33, 36, 39, 43, 46, 49, 52)
POOL_AFTER = (2, 5, 8, 12, 16)
HEADS = 8
DEPTH = 6
BN_EPS = 1e-5
LN_EPS = 1e-5


def _q(t: torch.Tensor, on: bool) -> torch.Tensor:
    return t.to(torch.bfloat16).to(torch.float32) if on else t


# --------------------------------------------------------------------------- preprocessing
def normalize_crops(crops_u8: torch.Tensor) -> torch.Tensor:
    """uint8 [n,224,224,3] -> fp32 [n,3,224,224]; cvit_prediction.py:209-215."""
    x = crops_u8.to(torch.float32).permute(0, 3, 1, 2)
    mean = torch.tensor(MEAN, dtype=torch.float32).view(1, 3, 1, 1)
    std = torch.tensor(STD, dtype=torch.float32).view(1, 3, 1, 1)
    return ((x / 255.0) - mean) / std


# --------------------------------------------------------------------------- model pieces
def feature_layer(x: torch.Tensor, sd: Dict[str, torch.Tensor], layer: int, bf16_sim: bool = False) -> torch.Tensor:
    """One conv3x3(pad 1)+BN(eval)+ReLU (+MaxPool2d(2) if the reference has one here)."""
    ci = CONV_IDX[layer]
    bi = ci + 1
    w = _q(sd[f"features.{ci}.weight"], bf16_sim and layer > 0)
    xin = _q(x, bf16_sim and layer > 0)
    y = F.conv2d(xin, w, sd[f"features.{ci}.bias"], stride=1, padding=1)
    y = F.batch_norm(y, sd[f"features.{bi}.running_mean"], sd[f"features.{bi}.running_var"],
                     sd[f"features.{bi}.weight"], sd[f"features.{bi}.bias"], False, 0.0, BN_EPS)
    y = F.relu(y)
    if layer in POOL_AFTER:
        y = F.max_pool2d(y, kernel_size=2, stride=2)
    return y


def features(x: torch.Tensor, sd, bf16_sim: bool = False, upto: int = 17) -> torch.Tensor:
    for layer in range(upto):
        x = feature_layer(x, sd, layer, bf16_sim)
    return x


def embed_tokens(feat: torch.Tensor, sd, slots: torch.Tensor, bf16_sim: bool = False) -> torch.Tensor:
    """[n,512,7,7] -> tokens [n,2,1024]; model/cvit.py:170-175.

    rearrange 'b c (h p1) (w p2) -> b (h w) (p1 p2 c)' with h=w=1, p=7 is the NHWC flatten.
    ``x += pos_embedding[0:b]`` broadcasts pos[s] over BOTH tokens of batch slot s.
    """
    n = feat.shape[0]
    y = feat.permute(0, 2, 3, 1).reshape(n, 1, -1)
    y = F.linear(_q(y, bf16_sim), _q(sd["patch_to_embedding.weight"], bf16_sim), sd["patch_to_embedding.bias"])
    cls = sd["cls_token"].expand(n, -1, -1)
    x = torch.cat((cls, y), 1)
    return x + sd["pos_embedding"][slots.long()]           # [n,1,1024] broadcast over tokens


def attention(x: torch.Tensor, sd, p: str, bf16_sim: bool = False) -> torch.Tensor:
    b, n, dim = x.shape
    h = HEADS
    d = dim // h
    qkv = F.linear(_q(x, bf16_sim), _q(sd[p + ".to_qkv.weight"], bf16_sim))
    qkv = qkv.view(b, n, 3, h, d)                           # '(qkv h d)'  model/cvit.py:46
    q, k, v = (qkv[:, :, i].permute(0, 2, 1, 3) for i in range(3))   # b h n d
    dots = torch.einsum("bhid,bhjd->bhij", q, k) * (dim ** -0.5)     # scale = dim**-0.5  (:38)
    attn = dots.softmax(dim=-1)
    out = torch.einsum("bhij,bhjd->bhid", attn, v)
    out = out.permute(0, 2, 1, 3).reshape(b, n, dim)        # 'b h n d -> b n (h d)'
    return F.linear(_q(out, bf16_sim), _q(sd[p + ".to_out.weight"], bf16_sim), sd[p + ".to_out.bias"])


def transformer(x: torch.Tensor, sd, bf16_sim: bool = False, depth: int = DEPTH) -> torch.Tensor:
    dim = x.shape[-1]
    for layer in range(depth):
        p = f"transformer.layers.{layer}"
        y = F.layer_norm(x, (dim,), sd[f"{p}.0.fn.norm.weight"], sd[f"{p}.0.fn.norm.bias"], LN_EPS)
        x = attention(y, sd, f"{p}.0.fn.fn", bf16_sim) + x
        y = F.layer_norm(x, (dim,), sd[f"{p}.1.fn.norm.weight"], sd[f"{p}.1.fn.norm.bias"], LN_EPS)
        y = F.linear(_q(y, bf16_sim), _q(sd[f"{p}.1.fn.fn.net.0.weight"], bf16_sim), sd[f"{p}.1.fn.fn.net.0.bias"])
        y = F.gelu(y)                                        # exact erf GELU (nn.GELU default)
        y = F.linear(_q(y, bf16_sim), _q(sd[f"{p}.1.fn.fn.net.2.weight"], bf16_sim), sd[f"{p}.1.fn.fn.net.2.bias"])
        x = y + x
    return x


def head(x: torch.Tensor, sd, bf16_sim: bool = False) -> torch.Tensor:
    c = x[:, 0]
    y = F.relu(F.linear(_q(c, bf16_sim), _q(sd["mlp_head.0.weight"], bf16_sim), sd["mlp_head.0.bias"]))
    return F.linear(y, sd["mlp_head.2.weight"], sd["mlp_head.2.bias"])


def forward_slots(x: torch.Tensor, sd, slots: torch.Tensor, bf16_sim: bool = False) -> torch.Tensor:
    """fp32 NCHW [n,3,224,224] + slot[n] in [0,32) -> logits [n,2]."""
    with torch.no_grad():
        f = features(x, sd, bf16_sim)
        t = embed_tokens(f, sd, slots, bf16_sim)
        t = transformer(t, sd, bf16_sim)
        return head(t, sd, bf16_sim)


def forward(x: torch.Tensor, sd, bf16_sim: bool = False) -> torch.Tensor:
    """Reference-compatible call: slot = batch index; b > 32 raises like the reference."""
    b = x.shape[0]
    if b > 32:
        raise RuntimeError("CViT.forward: batch > 32 cannot broadcast against pos_embedding[0:32] (cvit.py:175)")
    return forward_slots(x, sd, torch.arange(b), bf16_sim)


def forward_chunked(x: torch.Tensor, sd, slots: Optional[torch.Tensor] = None, chunk: int = 32,
                    bf16_sim: bool = False) -> torch.Tensor:
    """Any n: evaluated ``chunk`` crops at a time; default slot = i % 32."""
    n = x.shape[0]
    if slots is None:
        slots = torch.arange(n) % 32
    outs = [forward_slots(x[a:a + chunk], sd, slots[a:a + chunk], bf16_sim) for a in range(0, n, chunk)]
    return torch.cat(outs, 0) if outs else torch.zeros((0, 2))


# --------------------------------------------------------------------------- per-video reduction
def pred_sig(logits: torch.Tensor) -> torch.Tensor:
    """cvit_prediction.py:258-259 — independent sigmoid per logit (not softmax), after squeeze()."""
    return torch.sigmoid(logits.squeeze())


def pre_process_prediction(y_pred: torch.Tensor) -> torch.Tensor:
    """cvit_prediction.py:266-281 — len() of the squeezed tensor decides; sequential fp32 sums."""
    f: List[torch.Tensor] = []
    r: List[torch.Tensor] = []
    if len(y_pred) > 2:
        for row in y_pred:
            i, j = row
            f.append(i)
            r.append(j)
        f_c = sum(f) / len(f)
        r_c = sum(r) / len(r)
        if f_c > r_c:
            return f_c
        return abs(1 - r_c)
    return torch.tensor(0.5)


def video_score(logits: torch.Tensor) -> float:
    """logits [n,2] of ONE video -> score; n == 0 -> 0.5 (cvit_prediction.py:218-219)."""
    if logits.shape[0] == 0:
        return 0.5
    return float(pre_process_prediction(pred_sig(logits)).item())


def video_scores(logits: torch.Tensor, offsets: Sequence[int]) -> List[float]:
    return [video_score(logits[offsets[v]:offsets[v + 1]]) for v in range(len(offsets) - 1)]


def predict_from_crops(crops_u8: torch.Tensor, sd, bf16_sim: bool = False):
    """The model half of ``predict`` (cvit_prediction.py:209-242) for one video's crops.

    Returns (score, logits). Chunks are [0:32], [32:64], [64:90]; crops >= 90 are dropped.
    """
    n = crops_u8.shape[0]
    if n == 0:
        return 0.5, torch.zeros((0, 2))
    x = normalize_crops(crops_u8)
    outs = [forward(x[0:min(n, 32)], sd, bf16_sim)]
    if n > 32:
        outs.append(forward(x[32:min(n, 64)], sd, bf16_sim))
    if n > 64:
        outs.append(forward(x[64:min(n, 90)], sd, bf16_sim))
    logits = torch.cat(outs, 0)
    return video_score(logits), logits


def real_or_fake(score: float) -> str:
    """cvit_prediction.py:284-292 / README: < 0.5 REAL, >= 0.5 FAKE."""
    return "REAL" if score < 0.5 else "FAKE"


# --------------------------------------------------------------------------- work accounting
FLOPS_PER_CROP = 13_291_528_192          # BASELINE.md §2 (2*MAC, convs + linears + attention einsums)


def count_flops_per_crop() -> int:
    """Recompute BASELINE.md's 13.2915 GFLOP/crop from the layer plan (sanity check for bench.py)."""
    plan = [(3, 32, 224), (32, 32, 224), (32, 32, 224), (32, 64, 112), (64, 64, 112), (64, 64, 112),
            (64, 128, 56), (128, 128, 56), (128, 128, 56), (128, 256, 28), (256, 256, 28), (256, 256, 28),
            (256, 256, 28), (256, 512, 14), (512, 512, 14), (512, 512, 14), (512, 512, 14)]
    fl = sum(2 * 9 * cin * cout * hw * hw for cin, cout, hw in plan)
    fl += 2 * 25088 * 1024
    per_tok = 2 * (1024 * 3072 + 1024 * 1024 + 1024 * 2048 + 2048 * 1024)
    fl += DEPTH * 2 * per_tok
    fl += DEPTH * 2 * (2 * HEADS * 2 * 2 * 128)       # two einsums, 8 heads, 2x2 scores, d=128
    fl += 2 * (1024 * 2048 + 2048 * 2)
    return fl
