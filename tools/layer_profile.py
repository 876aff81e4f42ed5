"""Per-layer device time of one 512-crop step (CUDA-event pairs around every launch) + achieved TFLOP/s.

    python tools/layer_profile.py [--crops 512] [--steps 5] [--s12 16]
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fac_fake_b200 import CViTEngine, weights as W  # noqa: E402

PLAN = [(3, 32, 224), (32, 32, 224), (32, 32, 224), (32, 64, 112), (64, 64, 112), (64, 64, 112),
        (64, 128, 56), (128, 128, 56), (128, 128, 56), (128, 256, 28), (256, 256, 28), (256, 256, 28),
        (256, 256, 28), (256, 512, 14), (512, 512, 14), (512, 512, 14), (512, 512, 14)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--crops", type=int, default=512)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--s12", type=int, default=0)
    args = ap.parse_args()
    sd = W.make_state_dict(0, "default")
    eng = CViTEngine(max_crops=512).to("cuda:0").load_state_dict(sd)
    if args.s12:
        eng.set_tuning(stage12_sub_batch=args.s12)
    n = args.crops
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
    eng.set_profiling(2)
    for i in range(args.steps):
        eng.predict_videos(crops[i % 2], offs)
    torch.cuda.synchronize()
    coarse = eng.get_profile(per_layer=True)
    names = list(coarse.keys())
    ph = [coarse[names[i]][0] / args.steps for i in range(3)]
    print(f"coarse phases (PDL intact): layers1-6 {ph[0]:.3f} ms | layers7-17 {ph[1]:.3f} ms | embed+transformer+head {ph[2]:.3f} ms")
    eng.set_profiling(True)
    e0.record()
    for i in range(args.steps):
        eng.predict_videos(crops[i % 2], offs)
    e1.record()
    torch.cuda.synchronize()
    prof_ms = e0.elapsed_time(e1) / args.steps
    prof = eng.get_profile(per_layer=True)
    print(f"s12={args.s12 or 'default'} crops={n}: step {plain_ms:.3f} ms "
          f"({n/plain_ms*1e3:.0f} crops/s), with per-launch events {prof_ms:.3f} ms")
    tot = 0.0
    for i, (name, (ms, cnt)) in enumerate(prof.items()):
        ms /= args.steps
        tot += ms
        if i < 17:
            cin, cout, hw = PLAN[i]
            fl = 2 * 9 * cin * cout * hw * hw * n
            print(f"  {name:>16s} {cin:3d}->{cout:3d} @{hw:3d}: {ms:8.3f} ms  {cnt/args.steps:5.0f} launches  {fl/max(ms, 1e-9)/1e9:8.1f} TFLOP/s")
        else:
            print(f"  {name:>16s}              : {ms:8.3f} ms  {cnt/args.steps:5.0f} launches")
    print(f"  sum of kernels {tot:.3f} ms")


if __name__ == "__main__":
    main()
