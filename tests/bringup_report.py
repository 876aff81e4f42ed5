"""GPU bring-up report: per-step error of the engine against the CPU oracle (no asserts).

    python tests/bringup_report.py [--n 4] [--variant bn] [--compute bf16]

Prints one line per debug tap (conv layers 1..17, tokens, transformer layers, logits) with
max-abs error, relative error and the oracle's scale, so that a single GPU call localises
a broken kernel.  Test infrastructure (uses oracle/).
"""
import argparse
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))  # repo root (this file lives in tests/)
sys.path.insert(0, ROOT)
from fac_fake_b200 import CViTEngine, weights as W  # noqa: E402
from oracle import cvit_oracle as O  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=4)
    ap.add_argument("--variant", default="bn")
    ap.add_argument("--compute", default="bf16")
    ap.add_argument("--steps", default="all")
    args = ap.parse_args()
    torch.set_num_threads(os.cpu_count() or 8)
    sd = W.make_state_dict(0, args.variant)
    crops = W.synthetic_crops(args.n, seed=1)
    x = O.normalize_crops(crops)
    slots = torch.arange(args.n) % 32
    t0 = time.time()
    eng = CViTEngine(max_crops=64, compute_dtype=args.compute).to("cuda:0").load_state_dict(sd)
    print(f"engine ready in {time.time()-t0:.1f}s", flush=True)
    xg = crops.cuda()
    # oracle activations
    acts = {}
    with torch.no_grad():
        h = x
        for li in range(17):
            h = O.feature_layer(h, sd, li)
            acts[li + 1] = h.permute(0, 2, 3, 1).contiguous().flatten()
        t = O.embed_tokens(h, sd, slots)
        acts[18] = t.flatten()
        for l in range(6):
            # run the oracle transformer one layer at a time
            sub = {k.replace(f"transformer.layers.{l}.", "transformer.layers.0."): v for k, v in sd.items()
                   if k.startswith(f"transformer.layers.{l}.")}
            t = O.transformer(t, sub, depth=1)
            acts[19 + l] = t.flatten()
        acts[25] = O.head(t, sd).flatten()
    steps = range(1, 26) if args.steps == "all" else [int(s) for s in args.steps.split(",")]
    for step in steps:
        try:
            got = eng.debug_activation(xg, step)
        except Exception as e:  # noqa: BLE001
            print(f"step {step:2d}: ERROR {e}", flush=True)
            break
        ref = acts[step]
        if got.numel() != ref.numel():
            print(f"step {step:2d}: size mismatch got {got.numel()} want {ref.numel()}")
            continue
        d = (got - ref).abs()
        scale = ref.abs().max().item()
        bad = (~torch.isfinite(got)).sum().item()
        print(f"step {step:2d}: max|d|={d.max().item():.4e} mean|d|={d.mean().item():.4e} ref_max={scale:.4e} "
              f"rel={d.max().item()/max(scale,1e-30):.3e} nonfinite={bad}", flush=True)
    lg = eng.forward_slots(xg, slots.cuda()).cpu()
    print("logits engine:", lg[:4].tolist())
    print("logits oracle:", acts[25].view(-1, 2)[:4].tolist())
    xf = x.cuda()
    lg2 = eng.forward_slots(xf, None).cpu()
    print("fp32-NCHW input path max diff vs u8 path:", (lg2 - lg).abs().max().item())


if __name__ == "__main__":
    main()
