"""``S3DEngine`` — host-side mirror of the reference's S3D clip classifier.

Used where the reference builds and calls its module
(/root/reference/sx_exp_deepfakedetect-master/S3D/S3D-test.py:199-212,267-279):

    model = S3DEngine(num_class=1, SRM_net='no', frames_per_clip=64)      # or SRM_net='yes'
    model.load_state_dict(state_dict)          # 'module.' prefixes of DataParallel checkpoints are stripped (:199-205)
    logits = model(video_faces)                # fp32 [b,3,T,224,224], raw 0..255 BGR  ->  [b, num_class]

Every kernel runs in ``libfacfake.so`` (``ff_s3d_*``, bf16 tcgen05 path); there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import torch

from . import _lib as L
from .engine import EngineError, _stream_ptr


class S3DEngine:
    def __init__(self, num_class: int = 1, SRM_net: str = "no", *, frames_per_clip: int = 64, max_clips: int = 8):
        if SRM_net not in ("no", "yes"):
            raise ValueError("SRM_net must be 'no' or 'yes' (model.py:10-14)")
        self.SRM_net = SRM_net                      # 'yes': 30 SRM high-pass filters in front of `base` (model.py:38-39)
        self.num_class = int(num_class)
        self.frames_per_clip = int(frames_per_clip)
        self._max_clips = int(max_clips)
        self._lib = L.load()
        self._h: Optional[C.c_void_p] = None
        self._device: Optional[torch.device] = None
        self.training = False

    def to(self, device):
        device = torch.device(device)
        if device.type != "cuda":
            raise EngineError("S3DEngine runs on CUDA (B200, sm_100a) only; there is no CPU fallback")
        if device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        self._device = device
        return self

    def cuda(self, device=None):
        return self.to("cuda" if device is None else device)

    def eval(self):
        return self

    def _err(self) -> str:
        msg = self._lib.ff_s3d_last_error(self._h)
        return msg.decode("utf-8", "replace") if msg else ""

    def _check(self, rc: int, what: str):
        if rc != L.FF_OK:
            raise (ValueError if rc in (L.FF_ERR_BAD_ARG, L.FF_ERR_SHAPE) else EngineError)(f"{what}: {self._err()} (code {rc})")

    def load_state_dict(self, state_dict: Dict[str, torch.Tensor], strict: bool = True):
        if self._device is None:
            self.to("cuda")
        if self._h is not None:
            self._lib.ff_s3d_destroy(self._h)
            self._h = None
        h = C.c_void_p()
        rc = self._lib.ff_s3d_create(C.byref(h), self._device.index, self._max_clips, self.frames_per_clip, self.num_class,
                                     1 if self.SRM_net == "yes" else 0)
        if rc != L.FF_OK:
            msg = self._lib.ff_s3d_last_error(None)
            raise (ValueError if rc == L.FF_ERR_BAD_ARG else EngineError)(f"ff_s3d_create failed: {msg.decode() if msg else ''} (code {rc})")
        self._h = h
        for key, t in state_dict.items():
            if key.startswith("module."):
                key = key[len("module."):]
            t = t.detach().to("cpu", torch.float32).contiguous()
            shape = (C.c_int64 * max(t.dim(), 1))(*t.shape) if t.dim() else (C.c_int64 * 1)(1)
            self._check(self._lib.ff_s3d_load_weight(self._h, key.encode(), C.c_void_p(t.data_ptr()), shape, t.dim()), f"load_weight({key})")
        self._check(self._lib.ff_s3d_finalize(self._h), "ff_s3d_finalize")
        return self

    def __del__(self):
        try:
            if getattr(self, "_h", None) is not None:
                self._lib.ff_s3d_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def _prep(self, x: torch.Tensor):
        if self._h is None:
            raise EngineError("load_state_dict() must be called before the forward pass")
        T = self.frames_per_clip
        if x.dtype == torch.uint8:
            if x.dim() != 5 or tuple(x.shape[1:]) != (T, 224, 224, 3):
                raise ValueError(f"uint8 clips must be [b,{T},224,224,3], got {tuple(x.shape)}")
            layout = L.FF_X_NHWC_U8
        else:
            if x.dim() != 5 or tuple(x.shape[1:]) != (3, T, 224, 224):
                raise ValueError(f"float clips must be [b,3,{T},224,224], got {tuple(x.shape)}")
            x = x.to(torch.float32)
            layout = L.FF_X_NCHW_F32
        return x.to(self._device).contiguous(), layout

    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        return self.forward(x)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """fp32 [b,3,T,224,224] (raw 0..255 BGR, S3D-test.py:94-96) or uint8 [b,T,224,224,3] -> logits [b, num_class].

        A pinned HOST tensor is streamed: chunk i+1 is copied on a side stream while chunk i is computed (a 64-frame
        uint8 clip is 9.6 MB, so the copy is as long as the compute and must not be serialised with it)."""
        if x.device.type == "cpu" and x.is_pinned() and x.shape[0] > 1 and self._h is not None:
            return self._forward_streamed(x)
        x, layout = self._prep(x)
        b = x.shape[0]
        out = torch.empty((b, self.num_class), dtype=torch.float32, device=self._device)
        with torch.cuda.device(self._device):
            rc = self._lib.ff_s3d_forward(self._h, C.c_void_p(x.data_ptr()), layout, b, C.c_void_p(out.data_ptr()),
                                          C.c_void_p(_stream_ptr(self._device)))
        self._check(rc, "ff_s3d_forward")
        return out

    def _forward_streamed(self, x: torch.Tensor, chunks: int = 4) -> torch.Tensor:
        b = x.shape[0]
        step = max(1, (b + chunks - 1) // chunks)
        bounds = [(i, min(b, i + step)) for i in range(0, b, step)]
        if x.dtype != torch.uint8:
            x = x.to(torch.float32)
        layout = L.FF_X_NHWC_U8 if x.dtype == torch.uint8 else L.FF_X_NCHW_F32
        expect = (self.frames_per_clip, 224, 224, 3) if x.dtype == torch.uint8 else (3, self.frames_per_clip, 224, 224)
        if tuple(x.shape[1:]) != expect:
            raise ValueError(f"clips must be [b,{','.join(map(str, expect))}], got {tuple(x.shape)}")
        out = torch.empty((b, self.num_class), dtype=torch.float32, device=self._device)
        with torch.cuda.device(self._device):
            main = torch.cuda.current_stream(self._device)
            if getattr(self, "_copy_stream", None) is None:
                self._copy_stream = torch.cuda.Stream(self._device)
                self._stage = [None, None]
            cs = self._copy_stream
            cs.wait_stream(main)
            done = [None, None]                       # compute-finished events guarding the two staging buffers
            for i, (lo, hi) in enumerate(bounds):
                slot = i & 1
                if self._stage[slot] is None or self._stage[slot].shape[0] < hi - lo or self._stage[slot].dtype != x.dtype:
                    self._stage[slot] = torch.empty((step,) + tuple(x.shape[1:]), dtype=x.dtype, device=self._device)
                with torch.cuda.stream(cs):
                    if done[slot] is not None:
                        cs.wait_event(done[slot])
                    self._stage[slot][: hi - lo].copy_(x[lo:hi], non_blocking=True)
                    ready = torch.cuda.Event()
                    ready.record(cs)
                main.wait_event(ready)
                rc = self._lib.ff_s3d_forward(self._h, C.c_void_p(self._stage[slot].data_ptr()), layout, hi - lo,
                                              C.c_void_p(out[lo:hi].data_ptr()), C.c_void_p(main.cuda_stream))
                self._check(rc, "ff_s3d_forward")
                done[slot] = torch.cuda.Event()
                done[slot].record(main)
        return out

    def video_score(self, clips: torch.Tensor) -> float:
        """S3D-test.py:267-279: mean over a video's clips of sigmoid(logit)."""
        return torch.sigmoid(self.forward(clips).double().flatten()).mean().item()

    def debug_activation(self, x: torch.Tensor, base_index: int) -> torch.Tensor:
        """Activation after ``base[base_index]`` (model.py:17-34) as a flat fp32 CPU tensor in [b,T',H',W',C] order."""
        x, layout = self._prep(x)
        b = x.shape[0]
        cap = b * self.frames_per_clip * 112 * 112 * 64
        out = torch.empty((cap,), dtype=torch.float32)
        with torch.cuda.device(self._device):
            cnt = self._lib.ff_s3d_debug_activation(self._h, C.c_void_p(x.data_ptr()), layout, b, int(base_index),
                                                    C.c_void_p(out.data_ptr()), cap, C.c_void_p(_stream_ptr(self._device)))
        if cnt < 0:
            self._check(int(cnt), "ff_s3d_debug_activation")
        return out[:cnt]

    def launch_count(self) -> int:
        return int(self._lib.ff_s3d_launch_count(self._h)) if self._h is not None else 0
