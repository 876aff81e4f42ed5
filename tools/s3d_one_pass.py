"""One S3D pass (for ncu): python tools/s3d_one_pass.py [clips] [frames]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fac_fake_b200 import S3DEngine, weights as W  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
T = int(sys.argv[2]) if len(sys.argv) > 2 else 64
eng = S3DEngine(1, "no", frames_per_clip=T, max_clips=n).to("cuda:0").load_state_dict(W.make_s3d_state_dict(0, "default"))
clips = W.synthetic_clips(n, T, seed=0).cuda()
for _ in range(2):
    lg = eng(clips)
torch.cuda.synchronize()
print("logits", lg[:4].flatten().tolist())
