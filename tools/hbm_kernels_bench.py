"""Achieved HBM bandwidth of the memory-bound kernels of the path (K0 preprocess, LayerNorm, per-video reduction).

    python tools/hbm_kernels_bench.py

Algorithmic bytes: K0 = h*w*3 read + 150,528 written per crop; LayerNorm = 1024*(4+2) B per row;
reduction = 8 B per frame + 4 B per video.  Timed with CUDA events on the launching stream, inputs larger than L2.
"""
import ctypes as C
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fac_fake_b200 import CViTEngine, weights as W  # noqa: E402


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
    eng = CViTEngine(max_crops=512).to("cuda:0").load_state_dict(W.make_state_dict(0, "default"))
    out = {}
    # ---- K0: 512 crops, mixed sizes (integer-ratio, fractional area, up-scaling)
    g = torch.Generator().manual_seed(0)
    sizes = [(448, 448), (400, 380), (300, 300), (640, 520), (224, 224), (180, 200), (905, 640), (500, 333)]
    crops = [torch.randint(0, 256, (sizes[i % len(sizes)][0], sizes[i % len(sizes)][1], 3), generator=g, dtype=torch.uint8).cuda()
             for i in range(512)]
    in_bytes = sum(c.numel() for c in crops)
    out_bytes = 512 * 224 * 224 * 3
    ms = timeit(lambda: eng.preprocess_crops(crops, swap_rb=True), iters=5, warm=2)
    out["K0_preprocess"] = {"ms": ms, "algorithmic_GB": (in_bytes + out_bytes) / 1e9, "GBps": (in_bytes + out_bytes) / ms / 1e6,
                            "note": "whole C-ABI call incl. Python/ctypes marshalling of 512 crop descriptors"}

    def kernel_only(cs):
        eng.set_profiling(True)
        for _ in range(5):
            eng.preprocess_crops(cs, swap_rb=True)
        torch.cuda.synchronize()
        k = eng.get_profile()["small_kernels"]
        eng.set_profiling(False)
        return k[0] / max(k[1], 1)
    kms = kernel_only(crops)
    out["K0_preprocess"]["kernel_ms"] = kms
    out["K0_preprocess"]["kernel_GBps"] = (in_bytes + out_bytes) / kms / 1e6
    for name, hw in (("K0_area_fast_448", (448, 448)), ("K0_area_frac_400x380", (400, 380)), ("K0_linear_180x200", (180, 200))):
        cs = [torch.randint(0, 256, (hw[0], hw[1], 3), generator=g, dtype=torch.uint8).cuda() for _ in range(512)]
        b = sum(c.numel() for c in cs) + out_bytes
        ms = timeit(lambda: eng.preprocess_crops(cs, swap_rb=True), iters=5, warm=2)
        kms = kernel_only(cs)
        out[name] = {"ms": ms, "GBps": b / ms / 1e6, "kernel_ms": kms, "kernel_GBps": b / kms / 1e6}
    # ---- per-video reduction: 8192 videos x 30 frames
    logits = torch.randn((8192 * 30, 2), device="cuda")
    offs = torch.arange(0, 8192 * 30 + 1, 30, dtype=torch.int32, device="cuda")
    ms = timeit(lambda: eng.video_scores(logits, offs), iters=20)
    out["K8_video_reduce"] = {"ms": ms, "GBps": (8192 * 30 * 8 + 8192 * 4) / ms / 1e6, "note": "245,760 frames: latency-bound (2 MB)"}
    out["hbm_peak_GBps"] = peaks["hbm_gbs"]
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
