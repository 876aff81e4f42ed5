"""``BlazeFaceEngine`` — host-side mirror of the reference's face detector object.

Used exactly where the reference uses ``BlazeFace`` (/root/reference/CViT-main/cvit_prediction.py:27-33,
helpers/helpers_face_extract_1.py:11-21,87,110):

    facedet = BlazeFaceEngine().to(device)
    facedet.load_weights("helpers/blazeface.pth"); facedet.load_anchors("helpers/anchors.npy")
    face_extractor = FaceExtractor(video_read_fn, facedet)      # the reference class, unmodified

The network and the box decoding run in ``libfacfake.so`` (``ff_blazeface_*``, fp32 CUDA kernels); the score mask and
the blending NMS are data dependent and run here on the host, as in the reference (blazeface.py:225-234,305-358).
Detections are returned as CPU tensors (the reference returns them on the model's device and every consumer calls
``.cpu()`` on them).
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional

import numpy as np
import torch

from . import _lib as L
from .engine import EngineError, _stream_ptr


class BlazeFaceEngine:
    input_size = (128, 128)                      # blazeface.py:62

    def __init__(self, *, max_tiles: int = 512):
        self.num_classes = 1                     # blazeface.py:69-79
        self.num_anchors = 896
        self.num_coords = 16
        self.score_clipping_thresh = 100.0
        self.x_scale = self.y_scale = self.h_scale = self.w_scale = 128.0
        self.min_score_thresh = 0.75
        self.min_suppression_threshold = 0.3
        self._lib = L.load()
        self._h: Optional[C.c_void_p] = None
        self._device: Optional[torch.device] = None
        self._max_tiles = int(max_tiles)
        self._state = None
        self.anchors: Optional[torch.Tensor] = None
        self._ready = False

    # ------------------------------------------------------------------ nn.Module-like surface
    def to(self, device):
        device = torch.device(device)
        if device.type != "cuda":
            raise EngineError("BlazeFaceEngine runs on CUDA only; there is no CPU fallback")
        if device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        self._device = device
        return self

    def eval(self):
        return self

    def _device_(self):
        return self._device

    def load_state_dict(self, state_dict):
        self._state = {k: v.detach().to("cpu", torch.float32).contiguous() for k, v in state_dict.items()}
        self._ready = False
        return self

    def load_weights(self, path):
        """blazeface.py:152-154 (``path`` may also be a state_dict)."""
        sd = torch.load(path, map_location="cpu") if isinstance(path, (str, bytes)) or hasattr(path, "__fspath__") else path
        return self.load_state_dict(sd)

    def load_anchors(self, path):
        """blazeface.py:156-160 (``path`` may also be an array)."""
        a = np.load(path) if isinstance(path, (str, bytes)) or hasattr(path, "__fspath__") else np.asarray(path)
        self.anchors = torch.tensor(a, dtype=torch.float32)
        assert self.anchors.ndimension() == 2 and self.anchors.shape == (self.num_anchors, 4)
        self._ready = False
        return self

    def _err(self) -> str:
        msg = self._lib.ff_blazeface_last_error(self._h)
        return msg.decode("utf-8", "replace") if msg else ""

    def _check(self, rc: int, what: str):
        if rc != L.FF_OK:
            raise (ValueError if rc in (L.FF_ERR_BAD_ARG, L.FF_ERR_SHAPE) else EngineError)(f"{what}: {self._err()} (code {rc})")

    def _ensure_ready(self):
        if self._ready:
            return
        if self._state is None or self.anchors is None:
            raise EngineError("load_weights() and load_anchors() must be called before predicting")
        if self._device is None:
            self.to("cuda")
        if self._h is not None:
            self._lib.ff_blazeface_destroy(self._h)
            self._h = None
        h = C.c_void_p()
        rc = self._lib.ff_blazeface_create(C.byref(h), self._device.index, self._max_tiles)
        if rc != L.FF_OK:
            msg = self._lib.ff_blazeface_last_error(None)
            raise EngineError(f"ff_blazeface_create failed: {msg.decode() if msg else ''} (code {rc})")
        self._h = h
        for key, t in list(self._state.items()) + [("anchors", self.anchors.contiguous())]:
            shape = (C.c_int64 * max(t.dim(), 1))(*t.shape) if t.dim() else (C.c_int64 * 1)(1)
            self._check(self._lib.ff_blazeface_load_weight(self._h, key.encode(), C.c_void_p(t.data_ptr()), shape, t.dim()),
                        f"load_weight({key})")
        self._check(self._lib.ff_blazeface_finalize(self._h), "ff_blazeface_finalize")
        self._ready = True

    def __del__(self):
        try:
            if getattr(self, "_h", None) is not None:
                self._lib.ff_blazeface_destroy(self._h)
                self._h = None
        except Exception:
            pass

    # ------------------------------------------------------------------ prediction
    def predict_dense(self, x, return_raw: bool = False):
        """uint8 tiles -> DEVICE tensor [b,896,17] (decoded box, keypoints, score for EVERY anchor)."""
        self._ensure_ready()
        if isinstance(x, np.ndarray):
            x = torch.from_numpy(x)                                 # (b, H, W, 3)
        elif x.dim() == 4 and x.shape[1] == 3 and x.shape[-1] != 3:
            x = x.permute(0, 2, 3, 1)                               # reference accepts (b, 3, H, W) tensors
        if x.dtype != torch.uint8:
            raise ValueError("BlazeFaceEngine expects uint8 pixels (the reference's _preprocess is applied on the GPU)")
        assert x.shape[1:] == (128, 128, 3), "tiles must be 128x128x3"
        x = x.to(self._device).contiguous()
        b = x.shape[0]
        det = torch.empty((b, self.num_anchors, 17), dtype=torch.float32, device=self._device)
        rb = torch.empty((b, self.num_anchors, 16), dtype=torch.float32, device=self._device) if return_raw else None
        rs = torch.empty((b, self.num_anchors), dtype=torch.float32, device=self._device) if return_raw else None
        with torch.cuda.device(self._device):
            rc = self._lib.ff_blazeface_predict(self._h, C.c_void_p(x.data_ptr()), b, C.c_void_p(det.data_ptr()),
                                                C.c_void_p(rb.data_ptr()) if return_raw else None,
                                                C.c_void_p(rs.data_ptr()) if return_raw else None,
                                                C.c_void_p(_stream_ptr(self._device)))
        self._check(rc, "ff_blazeface_predict")
        return (det, rb, rs) if return_raw else det

    def predict_on_image(self, img):
        """blazeface.py:166-180."""
        if isinstance(img, np.ndarray):
            img = torch.from_numpy(img).permute((2, 0, 1))
        return self.predict_on_batch(img.unsqueeze(0))[0]

    def predict_on_batch(self, x, apply_nms: bool = True) -> List[torch.Tensor]:
        """blazeface.py:182-223: list of (num_detections, 17) tensors, one per image."""
        dense = self.predict_dense(x)
        if apply_nms:
            return self._nms_on_device(dense)
        # score mask on the device (plumbing, like the reference's boolean indexing): only the survivors cross PCIe
        img, anchor = torch.nonzero(dense[..., 16] >= self.min_score_thresh, as_tuple=True)
        kept = dense[img, anchor].cpu()
        counts = torch.bincount(img, minlength=dense.shape[0]).cpu().tolist()
        detections = list(torch.split(kept, counts))
        return self.nms(detections) if apply_nms else detections

    def _nms_on_device(self, dense: torch.Tensor) -> List[torch.Tensor]:
        """Mask + blending NMS of every tile in one kernel (``ff_blazeface_nms``); tiles it cannot hold (> 64
        candidates or > 16 faces) fall back to the host loop."""
        b = dense.shape[0]
        faces = torch.empty((b, 16, 17), dtype=torch.float32, device=self._device)
        counts = torch.empty((b,), dtype=torch.int32, device=self._device)
        with torch.cuda.device(self._device):
            rc = self._lib.ff_blazeface_nms(self._h, C.c_void_p(dense.data_ptr()), b, C.c_float(self.min_score_thresh),
                                            C.c_float(self.min_suppression_threshold), C.c_void_p(faces.data_ptr()),
                                            C.c_void_p(counts.data_ptr()), C.c_void_p(_stream_ptr(self._device)))
        self._check(rc, "ff_blazeface_nms")
        counts_h = counts.cpu().tolist()
        faces_h = faces.cpu()
        out = []
        for i, k in enumerate(counts_h):
            if k >= 0:
                out.append(faces_h[i, :k].clone())
            else:
                d = dense[i].cpu()
                f = self._weighted_non_max_suppression(d[d[:, 16] >= self.min_score_thresh])
                out.append(torch.stack(f) if f else torch.zeros((0, 17)))
        return out

    def nms(self, detections: List[torch.Tensor]) -> List[torch.Tensor]:
        """blazeface.py:225-234: blending NMS of every list in one kernel (``ff_blazeface_nms_lists``); a list it cannot
        hold (> 64 detections or > 16 faces) is merged by the same algorithm on the host."""
        n = len(detections)
        if n == 0:
            return []
        sizes = [int(d.shape[0]) for d in detections]
        offs = torch.tensor([0] + list(np.cumsum(sizes)), dtype=torch.int32)
        out: List[Optional[torch.Tensor]] = [None] * n
        if offs[-1] > 0:
            self._ensure_ready()
            flat = torch.cat([d.reshape(-1, 17).to(torch.float32) for d in detections]).to(self._device).contiguous()
            offs_d = offs.to(self._device)
            faces = torch.empty((n, 16, 17), dtype=torch.float32, device=self._device)
            counts = torch.empty((n,), dtype=torch.int32, device=self._device)
            with torch.cuda.device(self._device):
                rc = self._lib.ff_blazeface_nms_lists(self._h, C.c_void_p(flat.data_ptr()), C.c_void_p(offs_d.data_ptr()), n,
                                                      C.c_float(self.min_suppression_threshold), C.c_void_p(faces.data_ptr()),
                                                      C.c_void_p(counts.data_ptr()), C.c_void_p(_stream_ptr(self._device)))
            self._check(rc, "ff_blazeface_nms_lists")
            faces_h, counts_h = faces.cpu(), counts.cpu().tolist()
            for i, k in enumerate(counts_h):
                if k >= 0:
                    out[i] = faces_h[i, :k].clone()
        for i in range(n):
            if out[i] is None:
                f = self._weighted_non_max_suppression(detections[i].cpu()) if sizes[i] else []
                out[i] = torch.stack(f) if f else torch.zeros((0, 17))
        return out

    def _weighted_non_max_suppression(self, detections: torch.Tensor) -> List[torch.Tensor]:
        """Blending NMS (blazeface.py:305-358): detections overlapping the most confident one by IoU > 0.3 are merged
        into their score-weighted mean; the merged score is the mean score."""
        if len(detections) == 0:
            return []
        det = detections.to(torch.float32)
        ymin, xmin, ymax, xmax = det[:, 0], det[:, 1], det[:, 2], det[:, 3]
        area = (ymax - ymin) * (xmax - xmin)
        order = torch.argsort(det[:, 16], descending=True)
        alive = torch.ones(len(det), dtype=torch.bool)
        out = []
        for i in order.tolist():
            if not alive[i]:
                continue
            idx = order[alive[order]]                               # remaining detections, most confident first
            ih = (torch.minimum(ymax[i], ymax[idx]) - torch.maximum(ymin[i], ymin[idx])).clamp(min=0)
            iw = (torch.minimum(xmax[i], xmax[idx]) - torch.maximum(xmin[i], xmin[idx])).clamp(min=0)
            inter = ih * iw
            iou = inter / (area[i] + area[idx] - inter)
            overlapping = idx[iou > self.min_suppression_threshold]
            alive[overlapping] = False
            alive[i] = False                                        # a degenerate (zero-area) box must still be consumed
            merged = det[i].clone()
            if len(overlapping) > 1:
                scores = det[overlapping, 16:17]
                total = scores.sum()
                merged[:16] = (det[overlapping, :16] * scores).sum(dim=0) / total
                merged[16] = total / len(overlapping)
            out.append(merged)
        return out

    def launch_count(self) -> int:
        return int(self._lib.ff_blazeface_launch_count(self._h)) if self._h is not None else 0
