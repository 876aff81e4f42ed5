"""``CViTEngine`` — host-side mirror of the reference's model seam for the hot path.

It is used exactly where the reference uses its ``nn.Module``
(/root/reference/CViT-main/cvit_prediction.py:62-70,229-238):

    model = CViTEngine(image_size=224, patch_size=7, num_classes=2, channels=512,
                       dim=1024, depth=6, heads=8, mlp_dim=2048)
    model.to(device); model.load_state_dict(checkpoint); model.eval()
    y = model(x[0:32])                     # fp32 NCHW in, logits [n,2] out

PyTorch tensors appear only at this boundary (device memory + streams); every kernel runs in
``libfacfake.so`` through the C-ABI of ``include/facfake.h``.
"""
from __future__ import annotations

import ctypes as C
import fnmatch
from typing import Dict, Optional, Sequence

import torch

from . import _lib as L


class EngineError(RuntimeError):
    pass


_FIXED = dict(image_size=224, patch_size=7, num_classes=2, channels=512, dim=1024, depth=6, heads=8, mlp_dim=2048)


def _stream_ptr(device: torch.device) -> int:
    return int(torch.cuda.current_stream(device).cuda_stream)


class CViTEngine:
    """Drop-in for ``cvit.CViT`` at inference time (model/cvit.py:80-179)."""

    # state_dict keys that exist in the reference module but are not on its inference path (glob patterns);
    # everything else that finalize did not consume is an "unexpected key" under strict loading
    _IGNORED_KEYS: tuple = ()

    def __init__(self, image_size=224, patch_size=7, num_classes=2, channels=512, dim=1024, depth=6, heads=8,
                 mlp_dim=2048, *, max_crops: int = 512, compute_dtype: str = "bf16"):
        given = dict(image_size=image_size, patch_size=patch_size, num_classes=num_classes, channels=channels,
                     dim=dim, depth=depth, heads=heads, mlp_dim=mlp_dim)
        if given != _FIXED:
            raise ValueError(f"CViTEngine is specialised for {_FIXED}; got {given}")
        if compute_dtype not in ("bf16", "fp32"):
            raise ValueError("compute_dtype must be 'bf16' or 'fp32'")
        self._lib = L.load()
        self._h: Optional[C.c_void_p] = None
        self._device: Optional[torch.device] = None
        self._max_crops = int(max_crops)
        self._compute = L.FF_COMPUTE_BF16 if compute_dtype == "bf16" else L.FF_COMPUTE_FP32
        self.compute_dtype = compute_dtype
        self.training = False

    # ------------------------------------------------------------------ nn.Module-like surface
    def to(self, device):
        device = torch.device(device)
        if device.type != "cuda":
            raise EngineError("CViTEngine runs on CUDA (B200, sm_100a) only; there is no CPU fallback")
        if device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        if self._h is not None and device != self._device:
            raise EngineError("engine already created on another device")
        self._device = device
        return self

    def cuda(self, device=None):
        return self.to("cuda" if device is None else device)

    def eval(self):
        self.training = False
        return self

    def train(self, mode: bool = True):
        if mode:
            raise EngineError("CViTEngine is inference-only")
        return self

    def _err(self) -> str:
        msg = self._lib.ff_last_error(self._h)
        return msg.decode("utf-8", "replace") if msg else ""

    def _check(self, rc: int, what: str):
        if rc == L.FF_OK:
            return
        msg = f"{what}: {self._err()} (code {rc})"
        if rc in (L.FF_ERR_BAD_ARG, L.FF_ERR_SHAPE):
            raise ValueError(msg)
        raise EngineError(msg)

    def _ensure_handle(self):
        if self._h is not None:
            return
        if self._device is None:
            self.to("cuda")
        h = C.c_void_p()
        rc = self._create(h)
        if rc != L.FF_OK:
            msg = self._lib.ff_last_error(None)
            raise EngineError(f"engine create failed: {msg.decode() if msg else ''} (code {rc})")
        self._h = h

    def _create(self, h) -> int:
        return self._lib.ff_cvit_create(C.byref(h), self._device.index, self._max_crops, self._compute)

    def load_state_dict(self, state_dict: Dict[str, torch.Tensor], strict: bool = True):
        """Accepts a bare CViT state_dict or ``{'state_dict': ...}`` (cvit_prediction.py:66-69).

        Missing keys always fail (``EngineError``), wrong shapes raise ``ValueError``.  With ``strict`` (the
        ``nn.Module.load_state_dict`` default) keys the model does not have raise too; ``strict=False`` ignores them
        like the reference's training script does (cvit_train.py:70-71)."""
        if "state_dict" in state_dict and isinstance(state_dict["state_dict"], dict):
            state_dict = state_dict["state_dict"]
        if self._h is not None:
            self._lib.ff_cvit_destroy(self._h)
            self._h = None
        self._ensure_handle()
        for key, t in state_dict.items():
            if key.startswith("module."):
                key = key[len("module."):]
            t = t.detach().to("cpu", torch.float32).contiguous()
            shape = (C.c_int64 * max(t.dim(), 1))(*t.shape) if t.dim() else (C.c_int64 * 1)(1)
            rc = self._lib.ff_cvit_load_weight(self._h, key.encode(), C.c_void_p(t.data_ptr()), shape, t.dim())
            self._check(rc, f"load_weight({key})")
        self._check(self._lib.ff_cvit_finalize_weights(self._h), "finalize_weights")
        if strict:
            raw = self._lib.ff_cvit_unused_keys(self._h)
            unused = [k for k in (raw.decode() if raw else "").split(",") if k]
            unexpected = [k for k in unused if not any(fnmatch.fnmatchcase(k, pat) for pat in self._IGNORED_KEYS)]
            if unexpected:
                self._lib.ff_cvit_destroy(self._h)
                self._h = None
                raise EngineError("Unexpected key(s) in state_dict: " + ", ".join(sorted(unexpected)[:8])
                                  + (" ..." if len(unexpected) > 8 else ""))
        return self

    def __del__(self):
        try:
            if getattr(self, "_h", None) is not None:
                self._lib.ff_cvit_destroy(self._h)
                self._h = None
        except Exception:
            pass

    # ------------------------------------------------------------------ forward
    def _require_ready(self):
        if self._h is None:
            raise EngineError("load_state_dict() must be called before the forward pass")

    def __call__(self, x: torch.Tensor, mask=None) -> torch.Tensor:
        return self.forward(x, mask)

    def forward(self, x: torch.Tensor, mask=None) -> torch.Tensor:
        """``model(x)``: fp32 NCHW [b,3,224,224] -> logits [b,2]; slot = batch index.

        Like the reference, b > 32 raises RuntimeError (``x += pos_embedding[0:b]`` cannot
        broadcast, model/cvit.py:175); use ``forward_slots`` for larger batches.
        """
        if mask is not None:
            raise EngineError("mask is always None on the reference prediction path (cvit_prediction.py:229)")
        if x.dim() != 4 or tuple(x.shape[1:]) != (3, 224, 224):
            raise ValueError(f"expected [b,3,224,224], got {tuple(x.shape)}")
        if x.shape[0] > 32:
            raise RuntimeError(
                f"The size of tensor a ({x.shape[0]}) must match the size of tensor b (32) at non-singleton "
                "dimension 0 (CViT.forward: batch > 32, model/cvit.py:175)")
        return self.forward_slots(x, None)

    def _checked_input(self, x: torch.Tensor):
        """(contiguous tensor, layout code) for a crop batch on the engine's device: uint8 NHWC [n,224,224,3] raw crops
        or fp32 NCHW [n,3,224,224] normalised — anything else raises instead of handing a wild pointer to the library."""
        if not isinstance(x, torch.Tensor):
            raise ValueError("input must be a torch.Tensor")
        if x.device != self._device:
            raise EngineError(f"input is on {x.device}, engine on {self._device}")
        if x.dtype == torch.uint8:
            if x.dim() != 4 or tuple(x.shape[1:]) != (224, 224, 3):
                raise ValueError(f"uint8 input must be [n,224,224,3], got {tuple(x.shape)}")
            layout = L.FF_X_NHWC_U8
        elif x.dtype == torch.float32:
            if x.dim() != 4 or tuple(x.shape[1:]) != (3, 224, 224):
                raise ValueError(f"fp32 input must be [n,3,224,224], got {tuple(x.shape)}")
            layout = L.FF_X_NCHW_F32
        else:
            raise ValueError(f"unsupported input dtype {x.dtype}")
        return x.contiguous(), layout

    def forward_slots(self, x: torch.Tensor, slots: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Any batch size.  x: fp32 NCHW normalised, or uint8 NHWC [n,224,224,3] raw crops.
        slots: int32 [n] in [0,32) (default i % 32)."""
        self._require_ready()
        x, layout = self._checked_input(x)
        n = x.shape[0]
        logits = torch.empty((n, 2), dtype=torch.float32, device=self._device)
        sp = None
        if slots is not None:
            slots = slots.to(self._device, torch.int32).contiguous()
            if slots.numel() != n:
                raise ValueError("slots must have one entry per crop")
            if n and (int(slots.min()) < 0 or int(slots.max()) > 31):
                raise ValueError("slots must be in [0,32)")
            sp = C.c_void_p(slots.data_ptr())
        if n == 0:
            return logits
        rc = self._lib.ff_cvit_forward(self._h, C.c_void_p(x.data_ptr()), layout, sp, n, C.c_void_p(logits.data_ptr()),
                                       C.c_void_p(_stream_ptr(self._device)))
        self._check(rc, "ff_cvit_forward")
        return logits

    # ------------------------------------------------------------------ reduction / fused predict
    def video_scores(self, logits: torch.Tensor, offsets: torch.Tensor, mode: int = L.FF_REDUCE_REFERENCE) -> torch.Tensor:
        self._require_ready()
        logits = logits.to(self._device, torch.float32).contiguous()
        offsets = offsets.to(self._device, torch.int32).contiguous()
        nv = offsets.numel() - 1
        scores = torch.empty((max(nv, 0),), dtype=torch.float32, device=self._device)
        if nv <= 0:
            return scores
        rc = self._lib.ff_video_scores(self._h, C.c_void_p(logits.data_ptr()), C.c_void_p(offsets.data_ptr()), nv, mode,
                                       C.c_void_p(scores.data_ptr()), C.c_void_p(_stream_ptr(self._device)))
        self._check(rc, "ff_video_scores")
        return scores

    def predict_videos(self, crops: torch.Tensor, offsets: Sequence[int], mode: int = L.FF_REDUCE_REFERENCE,
                       return_logits: bool = False):
        """Model half of ``predict()`` for many videos (cvit_prediction.py:209-242): crops of video v are rows
        [offsets[v], offsets[v+1]) — uint8 NHWC or normalised fp32 NCHW, on the engine's device."""
        self._require_ready()
        off_host = torch.as_tensor(list(offsets), dtype=torch.int32)
        nv = off_host.numel() - 1
        if nv < 0:
            raise ValueError("offsets needs at least one entry")
        crops, layout = self._checked_input(crops)
        if off_host.numel() and (int(off_host[0]) != 0 or bool((off_host[1:] < off_host[:-1]).any())):
            raise ValueError("offsets must start at 0 and be non-decreasing")
        n = int(off_host[-1]) if nv >= 0 and off_host.numel() else 0
        if n > crops.shape[0]:
            raise ValueError("offsets exceed the number of crops")
        off_dev = off_host.to(self._device)
        scores = torch.empty((nv,), dtype=torch.float32, device=self._device)
        logits = torch.empty((n, 2), dtype=torch.float32, device=self._device) if return_logits else None
        if nv == 0:
            return (scores, logits) if return_logits else scores
        off_c = (C.c_int32 * (nv + 1))(*off_host.tolist())
        rc = self._lib.ff_cvit_predict(self._h, C.c_void_p(crops.data_ptr()), layout, off_c, C.c_void_p(off_dev.data_ptr()), nv,
                                       mode, C.c_void_p(logits.data_ptr()) if logits is not None else None,
                                       C.c_void_p(scores.data_ptr()), C.c_void_p(_stream_ptr(self._device)))
        self._check(rc, "ff_cvit_predict")
        return (scores, logits) if return_logits else scores

    def predict_videos_host(self, crops_host: torch.Tensor, offsets: Sequence[int], mode: int = L.FF_REDUCE_REFERENCE) -> torch.Tensor:
        """End-to-end call with HOST buffers (pinned preferred): H2D of the uint8 crops, forward, reduction, D2H of
        the scores, stream-synchronised on return."""
        self._require_ready()
        if (crops_host.device.type != "cpu" or crops_host.dtype != torch.uint8 or crops_host.dim() != 4
                or tuple(crops_host.shape[1:]) != (224, 224, 3)):
            raise ValueError("crops_host must be a CPU uint8 tensor [n,224,224,3]")
        crops_host = crops_host.contiguous()
        off = list(offsets)
        nv = len(off) - 1
        scores = torch.empty((max(nv, 0),), dtype=torch.float32)
        if nv <= 0:
            return scores
        if off[0] != 0 or any(b < a for a, b in zip(off, off[1:])):
            raise ValueError("offsets must start at 0 and be non-decreasing")
        if off[-1] > crops_host.shape[0]:
            raise ValueError("offsets exceed the number of crops")
        off_c = (C.c_int32 * (nv + 1))(*off)
        rc = self._lib.ff_cvit_predict_host(self._h, C.c_void_p(crops_host.data_ptr()), off_c, nv, mode,
                                            C.c_void_p(scores.data_ptr()), C.c_void_p(_stream_ptr(self._device)))
        self._check(rc, "ff_cvit_predict_host")
        return scores

    # ------------------------------------------------------------------ preprocessing (K0)
    def preprocess_crops(self, crops: Sequence[torch.Tensor], swap_rb: bool = True, normalized: bool = False):
        """cv2.resize(INTER_AREA, 224) + RGB<->BGR swap for variable-size uint8 HWC CUDA crops
        (cvit_prediction.py:114-115).  Returns uint8 [n,224,224,3] (and the fp32 NCHW normalised tensor)."""
        self._ensure_handle()
        n = len(crops)
        out = torch.empty((n, 224, 224, 3), dtype=torch.uint8, device=self._device)
        norm = torch.empty((n, 3, 224, 224), dtype=torch.float32, device=self._device) if normalized else None
        if n == 0:
            return (out, norm) if normalized else out
        keep = []
        ptrs = (C.c_void_p * n)()
        hw = (C.c_int32 * (2 * n))()
        pitch = (C.c_int32 * n)()
        for i, c in enumerate(crops):
            if c.dtype != torch.uint8 or c.dim() != 3 or c.shape[2] != 3:
                raise ValueError("each crop must be uint8 [h,w,3]")
            if c.device != self._device:
                raise EngineError("crops must be on the engine's device")
            if c.stride(2) != 1 or c.stride(1) != 3:
                c = c.contiguous()
            keep.append(c)
            ptrs[i] = c.data_ptr()
            hw[2 * i], hw[2 * i + 1] = c.shape[0], c.shape[1]
            pitch[i] = c.stride(0)
        rc = self._lib.ff_preprocess_crops(self._h, ptrs, hw, pitch, n, int(bool(swap_rb)), C.c_void_p(out.data_ptr()),
                                           C.c_void_p(norm.data_ptr()) if norm is not None else None,
                                           C.c_void_p(_stream_ptr(self._device)))
        self._check(rc, "ff_preprocess_crops")
        return (out, norm) if normalized else out

    # ------------------------------------------------------------------ introspection / debug
    def launch_count(self) -> int:
        return int(self._lib.ff_cvit_launch_count(self._h)) if self._h is not None else 0

    def set_tuning(self, stage12_sub_batch: int = 0):
        self._require_ready()
        self._check(self._lib.ff_cvit_set_tuning(self._h, stage12_sub_batch), "ff_cvit_set_tuning")

    KERNEL_CLASSES = ("conv1", "tcgen05_conv", "tcgen05_gemm", "small_kernels")

    def set_profiling(self, enable):
        """False/0 off, True/1 per-launch event pairs, 2 coarse (three phases per pass, launches stay PDL-chained)."""
        self._require_ready()
        self._check(self._lib.ff_cvit_set_profiling(self._h, int(enable)), "ff_cvit_set_profiling")

    def get_profile(self, per_layer: bool = False):
        """{class: (milliseconds, launches)} accumulated since set_profiling(True).
        per_layer=True returns the raw 21 slots (see include/facfake.h)."""
        self._require_ready()
        ms = (C.c_double * 21)()
        cnt = (C.c_int64 * 21)()
        self._check(self._lib.ff_cvit_get_profile(self._h, ms, cnt), "ff_cvit_get_profile")
        if per_layer:
            names = ["conv1"] + [f"conv{i + 1}" for i in range(1, 17)] + ["gemm_embed", "gemm_transformer", "gemm_head", "small"]
            return {k: (float(ms[i]), int(cnt[i])) for i, k in enumerate(names)}
        agg = {"conv1": (float(ms[0]), int(cnt[0])),
               "tcgen05_conv": (float(sum(ms[1:17])), int(sum(cnt[1:17]))),
               "tcgen05_gemm": (float(sum(ms[17:20])), int(sum(cnt[17:20]))),
               "small_kernels": (float(ms[20]), int(cnt[20]))}
        return agg

    def debug_activation(self, x: torch.Tensor, stop_after: int, slots: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Activation after step `stop_after` (see include/facfake.h) as a flat fp32 CPU tensor."""
        self._require_ready()
        x, layout = self._checked_input(x)
        n = x.shape[0]
        cap = n * 224 * 224 * 32
        out = torch.empty((cap,), dtype=torch.float32)
        sp = None
        if slots is not None:
            slots = slots.to(self._device, torch.int32).contiguous()
            if slots.numel() != n:
                raise ValueError("slots must have one entry per crop")
            sp = C.c_void_p(slots.data_ptr())
        cnt = self._lib.ff_cvit_debug_activation(self._h, C.c_void_p(x.data_ptr()), layout, sp, n, stop_after,
                                                 C.c_void_p(out.data_ptr()), cap, C.c_void_p(_stream_ptr(self._device)))
        if cnt < 0:
            self._check(int(cnt), "ff_cvit_debug_activation")
        return out[:cnt]


class ResVitKanEngine(CViTEngine):
    """Drop-in for the ResVitKan ``CViT`` at inference time
    (/root/reference/CViT-main/ResVitKan/ResVitKan.py:284-329: ResNet-50 features + ViT + KAN head).

    Same surface as ``CViTEngine`` (``.to()``, ``.load_state_dict()``, ``.eval()``, ``model(x)``,
    ``forward_slots``, ``predict_videos``...); the state_dict keys are the reference module's.
    ``compute_dtype='fp32'`` selects the CUDA-core path (logits within 1e-4 of the reference, slow).
    """

    _IGNORED_KEYS = ("mlp_head.*",)      # defined by the module, not on its forward path (ResVitKan.py:296-300)

    def __init__(self, image_size=224, patch_size=7, num_classes=2, channels=512, dim=1024, depth=6, heads=8,
                 mlp_dim=2048, *, max_crops: int = 256, compute_dtype: str = "bf16"):
        super().__init__(image_size, patch_size, num_classes, channels, dim, depth, heads, mlp_dim,
                         max_crops=max_crops, compute_dtype=compute_dtype)

    def _create(self, h) -> int:
        return self._lib.ff_resvitkan_create(C.byref(h), self._device.index, self._max_crops, self._compute)


class CViTGGCAEngine(CViTEngine):
    """Drop-in for ``cvit_GGCA_ADD_DEConv_RepBn8.CViT`` at inference time
    (/root/reference/CViT-main/model/cvit_GGCA_ADD_DEConv_RepBn8.py:353-455): the CViT conv plan with DEConv blocks
    (folded to plain 3x3 kernels at load), one BN-less conv pair, the GGCA gate on the 7x7x512 map and
    LinearNorm (= LayerNorm eps 1e-6 in eval) in the MLP branches.  Same surface as ``CViTEngine``;
    ``compute_dtype='bf16'`` = tensor-core path (fp16 conv stack, logits within 2e-2), ``'fp32'`` = CUDA-core path (1e-4).
    """

    # RepBN / LinearNorm training-schedule state and the unused Deconv block (cvit_GGCA_ADD_DEConv_RepBn8.py:22-60,425)
    _IGNORED_KEYS = ("Deconv.*", "transformer.layers.*.1.fn.norm.norm2.*", "transformer.layers.*.1.fn.norm.warm",
                     "transformer.layers.*.1.fn.norm.iter", "transformer.layers.*.1.fn.norm.total_step",
                     "transformer.layers.*.1.fn.norm.r0", "*.num_batches_tracked")

    def __init__(self, image_size=224, patch_size=7, num_classes=2, channels=512, dim=1024, depth=6, heads=8,
                 mlp_dim=2048, *, max_crops: int = 512, compute_dtype: str = "bf16"):
        super().__init__(image_size, patch_size, num_classes, channels, dim, depth, heads, mlp_dim,
                         max_crops=max_crops, compute_dtype=compute_dtype)

    def _create(self, h) -> int:
        return self._lib.ff_cvit_ggca_create(C.byref(h), self._device.index, self._max_crops, self._compute)
