"""fac_fake_b200 — B200-native (sm_100a) engine for FAC_fake's CViT forgery-classification hot path.

Host side of the drop-in: a ctypes binding of ``libfacfake.so`` (C-ABI in ``include/facfake.h``)
plus mirrors of the reference's Python seam (``CViT`` module call, ``cvit_prediction`` helpers).
There is no CPU / PyTorch fallback: without the CUDA library the package raises.
"""
from .engine import CViTEngine, CViTGGCAEngine, EngineError, ResVitKanEngine  # noqa: F401
from .blazeface import BlazeFaceEngine  # noqa: F401
from .face_extract import FaceExtractorEngine  # noqa: F401
from .s3d import S3DEngine  # noqa: F401
from . import weights  # noqa: F401

__all__ = ["CViTEngine", "CViTGGCAEngine", "ResVitKanEngine", "BlazeFaceEngine", "FaceExtractorEngine", "S3DEngine", "EngineError", "weights"]
