"""Device time of the encoder (transformer) class for several batch sizes: python tools/xf_probe.py
(FF_XF=0 selects the per-op launches, FF_VERBOSE=1 prints the cluster occupancy)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fac_fake_b200 import CViTEngine, weights as W  # noqa: E402

eng = CViTEngine(max_crops=512).to("cuda:0").load_state_dict(W.make_state_dict(0, "default"))
for n in (32, 64, 128, 256, 512):
    crops = W.synthetic_crops(n, seed=0).cuda()
    offs = list(range(0, n + 1, 32))
    for _ in range(3):
        eng.predict_videos(crops, offs)
    torch.cuda.synchronize()
    eng.set_profiling(True)
    steps = 5
    for _ in range(steps):
        eng.predict_videos(crops, offs)
    torch.cuda.synchronize()
    prof = eng.get_profile(per_layer=True)
    eng.set_profiling(False)
    ms, cnt = prof["gemm_transformer"]
    print(f"n={n:4d} rows={2*n:5d}: encoder class {ms/steps*1e3:8.1f} us per step, {cnt/steps:.0f} launches")
