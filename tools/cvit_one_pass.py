"""One CViT pass (for ncu): python tools/cvit_one_pass.py [crops]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fac_fake_b200 import CViTEngine, weights as W  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
eng = CViTEngine(max_crops=n).to("cuda:0").load_state_dict(W.make_state_dict(0, "default"))
crops = W.synthetic_crops(n, seed=0).cuda()
for _ in range(2):
    scores = eng.predict_videos(crops, list(range(0, n + 1, 32)))
torch.cuda.synchronize()
print("scores", scores[:4].tolist())
