"""Developer aid: run the encoder once with the trace build of the library (`make trace` -> build/libfacfake_trace.so) and let
block 0 print the cycle stamps of one layer's barriers / accumulators.  python tools/xf_trace.py [n_crops]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fac_fake_b200 import _lib  # noqa: E402

_lib.LIB_PATH = os.path.join(ROOT, "build", os.environ.get("FF_TRACE_LIB", "libfacfake_trace.so"))
from fac_fake_b200 import CViTEngine, weights as W  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
eng = CViTEngine(max_crops=512).to("cuda:0").load_state_dict(W.make_state_dict(0, "default"))
crops = W.synthetic_crops(n, seed=0).cuda()
offs = list(range(0, n + 1, 32))
for _ in range(3):
    eng.predict_videos(crops, offs)
torch.cuda.synchronize()
