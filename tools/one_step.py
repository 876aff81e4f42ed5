"""A short command for ncu: N warm-up steps + one step of the default CViT path (512 uint8 crops, 16 videos x 32).

    python tools/one_step.py [--crops 512] [--steps 2]
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fac_fake_b200 import CViTEngine, weights as W  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--crops", type=int, default=512)
    ap.add_argument("--steps", type=int, default=2)
    args = ap.parse_args()
    eng = CViTEngine(max_crops=512).to("cuda:0").load_state_dict(W.make_state_dict(0, "default"))
    n = args.crops
    crops = W.synthetic_crops(n, seed=0).cuda()
    offs = list(range(0, n + 1, 32))
    for _ in range(args.steps):
        scores = eng.predict_videos(crops, offs)
    torch.cuda.synchronize()
    print("ok", float(scores.sum()), "launches", eng.launch_count())


if __name__ == "__main__":
    main()
