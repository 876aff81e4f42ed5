"""Extra comparator: the same networks as plain torch modules in bf16 channels_last on the B200,
i.e. what the reference would get from cuDNN / cuBLAS by calling `.cuda().to(bfloat16)`.  NOT the reference path
(that is fp32 on host cores, see bench.py --impl reference) and NOT product code: it only says how far the
hand-written sm_100a kernels are from the vendor-library path on the same box.

    python tools/eager_comparator.py [--crops 512] [--steps 10] [--model cvit|resvitkan|both]

The module definitions restate the architectures (/root/reference/CViT-main/model/cvit.py:80-179,
ResVitKan/ResVitKan.py:150-240,284-329) with random weights; only the timing is used.
"""
import argparse
import json
import os
import sys

import torch
import torch.nn as nn
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CVIT_PLAN = [(3, 32, 0), (32, 32, 0), (32, 32, 1), (32, 64, 0), (64, 64, 0), (64, 64, 1), (64, 128, 0), (128, 128, 0),
             (128, 128, 1), (128, 256, 0), (256, 256, 0), (256, 256, 0), (256, 256, 1), (256, 512, 0), (512, 512, 0),
             (512, 512, 0), (512, 512, 1)]


class Block(nn.Module):
    def __init__(self, dim=1024, heads=8, mlp=2048):
        super().__init__()
        self.n1, self.n2 = nn.LayerNorm(dim), nn.LayerNorm(dim)
        self.qkv = nn.Linear(dim, 3 * dim, bias=False)
        self.out = nn.Linear(dim, dim)
        self.f1, self.f2 = nn.Linear(dim, mlp), nn.Linear(mlp, dim)
        self.heads, self.scale = heads, dim ** -0.5

    def forward(self, x):
        b, n, d = x.shape
        q, k, v = self.qkv(self.n1(x)).view(b, n, 3, self.heads, d // self.heads).permute(2, 0, 3, 1, 4)
        att = ((q @ k.transpose(-1, -2)) * self.scale).softmax(-1)
        x = x + self.out((att @ v).transpose(1, 2).reshape(b, n, d))
        return x + self.f2(F.gelu(self.f1(self.n2(x))))


class Tail(nn.Module):
    def __init__(self, kan=False):
        super().__init__()
        self.embed = nn.Linear(25088, 1024)
        self.cls = nn.Parameter(torch.randn(1, 1, 1024))
        self.pos = nn.Parameter(torch.randn(32, 1, 1024))
        self.blocks = nn.Sequential(*[Block() for _ in range(6)])
        self.h1 = nn.Linear(1024, 2048)
        self.h2 = nn.Linear(2048, 2)      # the KAN head is < 0.1 % of the FLOPs; a Linear stands in for it here

    def forward(self, f):                 # f: [b, 512, 7, 7] channels_last == NHWC flatten
        b = f.shape[0]
        y = self.embed(f.permute(0, 2, 3, 1).reshape(b, 1, 25088))
        x = torch.cat([self.cls.expand(b, -1, -1), y], 1) + self.pos[torch.arange(b, device=f.device) % 32]
        x = self.blocks(x)
        return self.h2(F.relu(self.h1(x[:, 0])))


class CViT(nn.Module):
    def __init__(self):
        super().__init__()
        layers = []
        for cin, cout, pool in CVIT_PLAN:
            layers += [nn.Conv2d(cin, cout, 3, padding=1), nn.BatchNorm2d(cout), nn.ReLU(inplace=True)]
            if pool:
                layers.append(nn.MaxPool2d(2))
        self.features = nn.Sequential(*layers)
        self.tail = Tail()

    def forward(self, x):
        return self.tail(self.features(x))


class Bottleneck(nn.Module):
    def __init__(self, inpl, planes, stride, down):
        super().__init__()
        self.c1, self.b1 = nn.Conv2d(inpl, planes, 1, bias=False), nn.BatchNorm2d(planes)
        self.c2, self.b2 = nn.Conv2d(planes, planes, 3, stride, 1, bias=False), nn.BatchNorm2d(planes)
        self.c3, self.b3 = nn.Conv2d(planes, planes * 4, 1, bias=False), nn.BatchNorm2d(planes * 4)
        self.down = nn.Sequential(nn.Conv2d(inpl, planes * 4, 1, stride, bias=False), nn.BatchNorm2d(planes * 4)) if down else None

    def forward(self, x):
        o = F.relu(self.b1(self.c1(x)))
        o = F.relu(self.b2(self.c2(o)))
        o = F.relu(self.b3(self.c3(o)))
        return F.relu(o + (x if self.down is None else self.down(x)))


class ResVitKan(nn.Module):
    def __init__(self):
        super().__init__()
        mods = [nn.Conv2d(3, 64, 7, 2, 3, bias=False), nn.BatchNorm2d(64), nn.ReLU(inplace=True), nn.MaxPool2d(3, 2, 1)]
        inpl = 64
        for planes, blocks, stride in ((64, 3, 1), (128, 4, 2), (256, 6, 2), (512, 3, 2)):
            for b in range(blocks):
                mods.append(Bottleneck(inpl, planes, stride if b == 0 else 1, b == 0))
                inpl = planes * 4
        mods += [nn.Conv2d(2048, 512, 1, bias=False), nn.BatchNorm2d(512)]
        self.features = nn.Sequential(*mods)
        self.tail = Tail(kan=True)

    def forward(self, x):
        return self.tail(self.features(x))


def time_model(name, model, n, steps, chunk):
    model = model.eval().cuda().to(torch.bfloat16).to(memory_format=torch.channels_last)
    xs = [torch.randn(n, 3, 224, 224, device="cuda", dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
          for _ in range(2)]
    torch.backends.cudnn.benchmark = True

    def step(x):
        with torch.no_grad():
            return torch.cat([model(x[i:i + chunk]) for i in range(0, n, chunk)])

    for i in range(3):
        step(xs[i % 2])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        step(xs[i % 2])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return {"comparator": "torch eager bf16 channels_last (cuDNN/cuBLAS), not the reference path", "model": name,
            "crops_per_step": n, "chunk": chunk, "ms_per_step": round(ms, 3), "crops_per_s": round(n / ms * 1e3, 1),
            "torch": torch.__version__, "cudnn": torch.backends.cudnn.version()}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--crops", type=int, default=512)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--model", default="both", choices=["cvit", "resvitkan", "both"])
    args = ap.parse_args()
    torch.manual_seed(0)
    for name, ctor in (("cvit", CViT), ("resvitkan", ResVitKan)):
        if args.model in (name, "both"):
            for chunk in (32, 128):       # 32 = the reference's chunking (cvit_prediction.py:229-238); 128 = library-friendly
                print(json.dumps(time_model(name, ctor(), args.crops, args.steps, chunk)), flush=True)


if __name__ == "__main__":
    main()
