"""ResVitKan: device time of one pass by phase (CUDA-event pairs around every launch).

    python tools/rvk_profile.py [--crops 256] [--steps 5]
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fac_fake_b200 import ResVitKanEngine, weights as W  # noqa: E402

# 2 * MACs per crop of the ResNet-50 trunk as the reference defines it (ResVitKan.py:185-240)
def resnet_flops_per_crop():
    fl = {"stem": 2 * 112 * 112 * 64 * 3 * 49}
    inpl, hw = 64, 56
    for li, (planes, blocks, stride) in enumerate(((64, 3, 1), (128, 4, 2), (256, 6, 2), (512, 3, 2)), start=1):
        f = 0
        for b in range(blocks):
            s = stride if b == 0 else 1
            f += 2 * hw * hw * inpl * planes
            oh = hw // s
            f += 2 * oh * oh * planes * planes * 9
            f += 2 * oh * oh * planes * planes * 4
            if b == 0:
                f += 2 * oh * oh * inpl * planes * 4
            inpl, hw = planes * 4, oh
        fl[f"layer{li}"] = f
    fl["channel"] = 2 * 49 * 2048 * 512
    return fl


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--crops", type=int, default=256)
    ap.add_argument("--steps", type=int, default=5)
    args = ap.parse_args()
    n = args.crops
    eng = ResVitKanEngine(max_crops=n).to("cuda:0").load_state_dict(W.make_resvitkan_state_dict(0, "default"))
    crops = [W.synthetic_crops(n, seed=i).cuda() for i in range(2)]
    offs = list(range(0, n + 1, 32))
    for i in range(3):
        eng.predict_videos(crops[i % 2], offs)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        eng.predict_videos(crops[i % 2], offs)
    e1.record()
    torch.cuda.synchronize()
    plain_ms = e0.elapsed_time(e1) / args.steps
    eng.set_profiling(True)
    for i in range(args.steps):
        eng.predict_videos(crops[i % 2], offs)
    torch.cuda.synchronize()
    prof = list(eng.get_profile(per_layer=True).values())
    fl = resnet_flops_per_crop()
    print(f"ResVitKan crops={n}: pass {plain_ms:.3f} ms ({n / plain_ms * 1e3:.0f} crops/s), "
          f"{sum(fl.values()) / 1e9:.2f} GFLOP/crop in the ResNet trunk")
    names = ["stem", "layer1", "layer2", "layer3", "layer4", "channel"]
    for i, name in enumerate(names):
        ms, cnt = prof[i]
        ms /= args.steps
        print(f"  {name:>8s}: {ms:8.3f} ms  {cnt / args.steps:4.0f} launches  {fl[name] * n / ms / 1e9:8.1f} TFLOP/s")
    for name, i in (("embed GEMM", 17), ("transformer GEMMs", 18), ("head Linear", 19), ("small kernels + KAN", 20)):
        print(f"  {name:>20s}: {prof[i][0] / args.steps:8.3f} ms  {prof[i][1] / args.steps:4.0f} launches")


if __name__ == "__main__":
    main()
