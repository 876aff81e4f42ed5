"""CPU oracle for the BlazeFace face detector — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Restates /root/reference/CViT-main/helpers/blazeface.py in fp32 torch functional ops:

* ``preprocess``              :162-164   x / 127.5 - 1
* ``forward``                 :8-43 (BlazeBlock), :80-148 (backbones, heads, TFLite-style asymmetric padding,
                              anchor-major reshapes) -> raw boxes [b,896,16], raw scores [b,896,1]
* ``decode_boxes``            :277-303
* ``tensors_to_detections``   :236-275   clamp +-100, sigmoid, score >= 0.75 mask
* ``weighted_nms`` / ``nms``  :225-234, :305-358 (blending NMS, IoU > 0.3), ``overlap_similarity`` :363-405
* ``predict_on_batch``        :182-223

Pinned by tests/golden/blazeface_golden.npz: outputs of the reference class with the reference's own shipped weights
(helpers/blazeface.pth, helpers/anchors.npy) on tiles cut from the reference's sample videos
(oracle/make_golden.py:main_blazeface).
"""
from __future__ import annotations

from typing import Dict, List

import torch
import torch.nn.functional as F

# (cin, cout, stride) of backbone1.2..12 and backbone2.0..4 (blazeface.py:86-107)
BLOCKS1 = ((24, 24, 1), (24, 28, 1), (28, 32, 2), (32, 36, 1), (36, 42, 1), (42, 48, 2), (48, 56, 1), (56, 64, 1),
           (64, 72, 1), (72, 80, 1), (80, 88, 1))
BLOCKS2 = ((88, 96, 2), (96, 96, 1), (96, 96, 1), (96, 96, 1), (96, 96, 1))
NUM_ANCHORS = 896
SCORE_CLIP = 100.0
MIN_SCORE = 0.75
MIN_SUPPRESSION = 0.3
SCALE = 128.0


def preprocess(x_u8_nchw: torch.Tensor) -> torch.Tensor:
    return x_u8_nchw.float() / 127.5 - 1.0


def blaze_block(x, sd: Dict[str, torch.Tensor], p: str, cin: int, cout: int, stride: int):
    if stride == 2:
        h = F.pad(x, (0, 2, 0, 2), "constant", 0)
        x = F.max_pool2d(x, kernel_size=2, stride=2)
        pad = 0
    else:
        h = x
        pad = 1
    h = F.conv2d(h, sd[p + ".convs.0.weight"], sd[p + ".convs.0.bias"], stride=stride, padding=pad, groups=cin)
    h = F.conv2d(h, sd[p + ".convs.1.weight"], sd[p + ".convs.1.bias"])
    if cout > cin:
        x = F.pad(x, (0, 0, 0, 0, 0, cout - cin), "constant", 0)
    return F.relu(h + x)


def backbone(x, sd, taps=None):
    x = F.pad(x, (1, 2, 1, 2), "constant", 0)
    x = F.relu(F.conv2d(x, sd["backbone1.0.weight"], sd["backbone1.0.bias"], stride=2))
    if taps is not None:
        taps.append(x)
    for i, (cin, cout, stride) in enumerate(BLOCKS1):
        x = blaze_block(x, sd, f"backbone1.{i + 2}", cin, cout, stride)
        if taps is not None:
            taps.append(x)
    h = x
    for i, (cin, cout, stride) in enumerate(BLOCKS2):
        h = blaze_block(h, sd, f"backbone2.{i}", cin, cout, stride)
        if taps is not None:
            taps.append(h)
    return x, h


def forward(x, sd, taps=None):
    """Preprocessed [b,3,128,128] -> (raw boxes [b,896,16], raw scores [b,896,1])."""
    b = x.shape[0]
    x, h = backbone(x, sd, taps)
    c1 = F.conv2d(x, sd["classifier_8.weight"], sd["classifier_8.bias"]).permute(0, 2, 3, 1).reshape(b, -1, 1)
    c2 = F.conv2d(h, sd["classifier_16.weight"], sd["classifier_16.bias"]).permute(0, 2, 3, 1).reshape(b, -1, 1)
    r1 = F.conv2d(x, sd["regressor_8.weight"], sd["regressor_8.bias"]).permute(0, 2, 3, 1).reshape(b, -1, 16)
    r2 = F.conv2d(h, sd["regressor_16.weight"], sd["regressor_16.bias"]).permute(0, 2, 3, 1).reshape(b, -1, 16)
    return torch.cat((r1, r2), 1), torch.cat((c1, c2), 1)


def decode_boxes(raw, anchors):
    boxes = torch.zeros_like(raw)
    xc = raw[..., 0] / SCALE * anchors[:, 2] + anchors[:, 0]
    yc = raw[..., 1] / SCALE * anchors[:, 3] + anchors[:, 1]
    w = raw[..., 2] / SCALE * anchors[:, 2]
    h = raw[..., 3] / SCALE * anchors[:, 3]
    boxes[..., 0] = yc - h / 2.0
    boxes[..., 1] = xc - w / 2.0
    boxes[..., 2] = yc + h / 2.0
    boxes[..., 3] = xc + w / 2.0
    for k in range(6):
        o = 4 + 2 * k
        boxes[..., o] = raw[..., o] / SCALE * anchors[:, 2] + anchors[:, 0]
        boxes[..., o + 1] = raw[..., o + 1] / SCALE * anchors[:, 3] + anchors[:, 1]
    return boxes


def dense_detections(raw_boxes, raw_scores, anchors):
    """[b,896,17]: decoded boxes + sigmoid(clamped score), before the score mask (what the GPU path returns)."""
    scores = raw_scores.clamp(-SCORE_CLIP, SCORE_CLIP).sigmoid()
    return torch.cat((decode_boxes(raw_boxes, anchors), scores), -1)


def tensors_to_detections(raw_boxes, raw_scores, anchors) -> List[torch.Tensor]:
    dense = dense_detections(raw_boxes, raw_scores, anchors)
    return [d[d[:, 16] >= MIN_SCORE] for d in dense]


def overlap_similarity(box, others):
    """IoU of one box [4] (ymin, xmin, ymax, xmax) against [k,4]."""
    mx = torch.min(box[2:].unsqueeze(0), others[:, 2:])
    mn = torch.max(box[:2].unsqueeze(0), others[:, :2])
    inter = torch.clamp(mx - mn, min=0)
    inter = inter[:, 0] * inter[:, 1]
    area_a = (box[2] - box[0]) * (box[3] - box[1])
    area_b = (others[:, 2] - others[:, 0]) * (others[:, 3] - others[:, 1])
    return inter / (area_a + area_b - inter)


def weighted_nms(det: torch.Tensor) -> List[torch.Tensor]:
    if len(det) == 0:
        return []
    out = []
    remaining = torch.argsort(det[:, 16], descending=True)
    while len(remaining) > 0:
        d = det[remaining[0]]
        ious = overlap_similarity(d[:4], det[remaining, :4])
        mask = ious > MIN_SUPPRESSION
        overlapping = remaining[mask]
        remaining = remaining[~mask]
        wd = d.clone()
        if len(overlapping) > 1:
            coords = det[overlapping, :16]
            scores = det[overlapping, 16:17]
            total = scores.sum()
            wd[:16] = (coords * scores).sum(0) / total
            wd[16] = total / len(overlapping)
        out.append(wd)
    return out


def nms(detections: List[torch.Tensor]) -> List[torch.Tensor]:
    res = []
    for d in detections:
        faces = weighted_nms(d)
        res.append(torch.stack(faces) if faces else torch.zeros((0, 17)))
    return res


def predict_on_batch(x_u8_nhwc, sd, anchors, apply_nms: bool = True):
    """uint8 [b,128,128,3] (numpy or tensor) -> list of [k,17] detections, like BlazeFace.predict_on_batch."""
    x = torch.as_tensor(x_u8_nhwc).permute(0, 3, 1, 2)
    with torch.no_grad():
        r, c = forward(preprocess(x), sd)
    det = tensors_to_detections(r, c, anchors)
    return nms(det) if apply_nms else det
