"""One BlazeFace batch (for ncu): python tools/blazeface_one_pass.py [tiles]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fac_fake_b200 import BlazeFaceEngine  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
w = np.load(os.path.join(ROOT, "tests", "golden", "blazeface_weights.npz"))
eng = BlazeFaceEngine(max_tiles=n).to("cuda:0")
eng.load_weights({k: torch.from_numpy(w[k]) for k in w.files if k != "anchors"})
eng.load_anchors(w["anchors"])
g = np.load(os.path.join(ROOT, "tests", "golden", "blazeface_golden.npz"))
tiles = torch.from_numpy(np.concatenate([g["tiles"]] * ((n + 11) // 12))[:n]).cuda()
for _ in range(2):
    det = eng.predict_dense(tiles)
torch.cuda.synchronize()
print("max score", det[..., 16].max().item())
