"""BlazeFace throughput: tiles/s on the GPU (inputs resident / from pinned host memory) next to the
CPU oracle, one JSON line.

    python tools/blazeface_bench.py [--tiles 512] [--steps 20]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fac_fake_b200 import BlazeFaceEngine  # noqa: E402

BLOCKS = ((24, 24, 1, 64), (24, 28, 1, 64), (28, 32, 2, 64), (32, 36, 1, 32), (36, 42, 1, 32), (42, 48, 2, 32), (48, 56, 1, 16),
          (56, 64, 1, 16), (64, 72, 1, 16), (72, 80, 1, 16), (80, 88, 1, 16), (88, 96, 2, 16), (96, 96, 1, 8), (96, 96, 1, 8),
          (96, 96, 1, 8), (96, 96, 1, 8))


def work_per_tile():
    """(2*MAC flops, fp32 activation bytes moved by the 19 launches: each reads its input and writes its output once)."""
    fl = 2 * 64 * 64 * 24 * 75
    by = 128 * 128 * 3 + 64 * 64 * 24 * 4
    for cin, cout, s, hw in BLOCKS:
        ho = hw // s
        fl += 2 * ho * ho * cin * 9 + 2 * ho * ho * cin * cout
        by += 4 * (hw * hw * cin + ho * ho * cout)
    fl += 2 * (256 * 88 * 34 + 64 * 96 * 102)
    by += 4 * (256 * 88 + 64 * 96 + 896 * 17 * 2 + 896 * 17)
    return fl, by


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tiles", type=int, default=512)
    ap.add_argument("--steps", type=int, default=20)
    args = ap.parse_args()
    w = np.load(os.path.join(ROOT, "tests", "golden", "blazeface_weights.npz"))
    sd = {k: torch.from_numpy(w[k]) for k in w.files if k != "anchors"}
    eng = BlazeFaceEngine(max_tiles=args.tiles).to("cuda:0")
    eng.load_weights(sd)
    eng.load_anchors(w["anchors"])
    g = np.load(os.path.join(ROOT, "tests", "golden", "blazeface_golden.npz"))
    reps = (args.tiles + len(g["tiles"]) - 1) // len(g["tiles"])
    host = torch.from_numpy(np.concatenate([g["tiles"]] * reps)[: args.tiles]).pin_memory()
    dev = host.cuda()
    for _ in range(3):
        eng.predict_dense(dev)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = eng.launch_count()
    e0.record()
    for _ in range(args.steps):
        eng.predict_dense(dev)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    launches = (eng.launch_count() - l0) // args.steps
    t0 = time.perf_counter()
    for _ in range(args.steps):
        faces = eng.predict_on_batch(host, apply_nms=True)           # H2D + network + D2H of the dense result + host NMS
    e2e_s = (time.perf_counter() - t0) / args.steps
    # CPU oracle on a bounded sample
    from oracle import blazeface_oracle as B
    torch.set_num_threads(os.cpu_count() or 1)
    sample = g["tiles"]
    B.predict_on_batch(sample, sd, torch.from_numpy(w["anchors"]))
    t0 = time.perf_counter()
    for _ in range(5):
        B.predict_on_batch(sample, sd, torch.from_numpy(w["anchors"]))
    cpu_s = (time.perf_counter() - t0) / 5
    fl, by = work_per_tile()
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
    gbs = by * args.tiles / (ms * 1e-3) / 1e9
    print(json.dumps({
        "metric": "BlazeFace tiles/sec", "value": args.tiles / ms * 1e3, "unit": "tiles/s", "ms_per_step": ms, "tiles_per_step": args.tiles,
        "gpu_launches_per_step": int(launches), "dtype": "f32", "faces_found": int(sum(len(f) for f in faces)),
        "e2e": {"value": args.tiles / e2e_s, "unit": "tiles/s", "what": "pinned host uint8 tiles -> H2D -> network+decode -> score mask on the device -> D2H of the survivors -> host blending NMS"},
        "roofline": {"bound": "hbm", "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": gbs / peaks["hbm_gbs"],
                     "algorithmic_bytes_per_tile": by, "flops_per_tile": fl, "gflops": fl * args.tiles / (ms * 1e-3) / 1e9},
        "cpu_baseline": {"value": len(sample) / cpu_s, "unit": "tiles/s", "cores": os.cpu_count(), "kind": "port",
                         "sample": f"{len(sample)} tiles x 5 runs, torch fp32 oracle incl. NMS"},
    }))


if __name__ == "__main__":
    main()
