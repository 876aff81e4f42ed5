#!/usr/bin/env python
"""bench.py — CViT hot-path throughput on B200 (BASELINE.json metric: CViT face-crops/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One step = one pass of the hot path (crop normalise -> CViT forward -> per-video score) over one
synthetic batch of 512 uint8 face crops 224x224 (BASELINE.json configs[1]; 16 videos x 32 crops,
slot = i % 32) per GPU.  With N > 1 (torchrun) every rank owns its own whole videos (weak scaling,
BASELINE.json configs[2] sharding); the only exchange is the all_gather of the per-video scores.

Printed JSON line (rank 0):
  value      crops/s, inputs already resident in HBM, CUDA-event timed, max over ranks
  e2e        same metric through the host-buffer C-ABI call (pinned host uint8 -> H2D -> forward -> scores D2H)
  roofline   the tcgen05 conv kernels: algorithmic conv FLOPs / their summed launch time (a CUDA-event pair around
             every launch on the launching stream, second pass of the same K steps) against the measured bf16 peak
  cpu_baseline  the CPU oracle port (torch fp32 on the host cores) on a bounded sample of the workload
--impl reference times that CPU path alone (the reference has no GPU kernels of its own).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOPS_PER_CROP = 13_291_528_192
CONV1_FLOPS = 2 * 27 * 32 * 224 * 224                 # CUDA-core conv1
CONV_FLOPS = 13_034_520_576                          # whole conv stack (17 layers), 2*MAC
TC_CONV_FLOPS = CONV_FLOPS - CONV1_FLOPS              # the 16 tcgen05 conv layers
CROPS_PER_STEP = 512
CROPS_PER_VIDEO = 32
METRIC = "CViT face-crops/sec"
UNIT = "crops/s"


def resvitkan_trunk_work():
    """(2*MAC flops, unfused bf16 activation bytes) per crop of the ResNet-50 trunk as the reference defines it
    (ResVitKan.py:185-240): every convolution reads its input once and writes its output once, conv3 also reads
    the residual; the stem path is uint8 crop -> bf16 NHWC4 -> conv -> max-pool."""
    fl = 2 * 112 * 112 * 64 * 3 * 49
    by = 224 * 224 * 3 + 2 * 224 * 224 * 8 + 2 * 112 * 112 * 64 * 2 + 56 * 56 * 64 * 2
    inpl, hw = 64, 56
    for planes, blocks, stride in ((64, 3, 1), (128, 4, 2), (256, 6, 2), (512, 3, 2)):
        for b in range(blocks):
            s = stride if b == 0 else 1
            oh = hw // s
            fl += 2 * hw * hw * inpl * planes + 2 * oh * oh * planes * planes * 9 + 2 * oh * oh * planes * planes * 4
            by += 2 * (hw * hw * inpl + hw * hw * planes)                 # conv1
            by += 2 * (hw * hw * planes + oh * oh * planes)               # conv2 (reads every input pixel)
            by += 2 * (oh * oh * planes + 2 * oh * oh * planes * 4)       # conv3: in + residual + out
            if b == 0:
                fl += 2 * oh * oh * inpl * planes * 4
                by += 2 * (oh * oh * inpl + oh * oh * planes * 4)         # strided 1x1 reads only the sampled pixels
            inpl, hw = planes * 4, oh
    fl += 2 * 49 * 2048 * 512
    by += 2 * (49 * 2048 + 49 * 512)
    return fl, by


MODELS = {
    "cvit": {"metric": METRIC, "flops_per_crop": FLOPS_PER_CROP, "max_crops": 512,
             "name": "CViT", "baseline_config": "configs[1]"},
    "ggca": {"metric": "CViT-GGCA-DEConv face-crops/sec", "flops_per_crop": FLOPS_PER_CROP + 2 * 9 * 128 * 128 * 56 * 56,
             "max_crops": 512, "name": "cvit_GGCA_ADD_DEConv_RepBn8", "baseline_config": "SURVEY 8f-4 (no BASELINE config)"},
    "resvitkan": {"metric": "ResVitKan face-crops/sec", "flops_per_crop": None, "max_crops": 512,
                  "name": "ResVitKan (ResNet-50 + ViT + KAN head)", "baseline_config": "configs[3]"},
}


_JSON_FD = None


def keep_stdout_clean():
    """stdout must carry exactly one JSON line, but NCCL prints its version banner there (NCCL_DEBUG=VERSION and up,
    not redirected by NCCL_DEBUG_FILE on this build).  Point fd 1 at stderr for the rest of the run and keep the real
    stdout for emit()."""
    global _JSON_FD
    if _JSON_FD is None and not os.environ.get("FF_KEEP_NCCL_DEBUG"):
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit(obj):
    line = json.dumps(obj) + "\n"
    if _JSON_FD is None:
        sys.stdout.write(line)
        sys.stdout.flush()
    else:
        sys.stdout.flush()
        os.write(_JSON_FD, line.encode())


def read_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"bf16_burst": d["bf16_tflops"], "bf16_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                "hbm_gbs": d["hbm_gbs"], "source": "measured (MEASURED_PEAKS.json)"}
    return {"bf16_burst": 1590.0, "bf16_sustained": 1400.0, "hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None
        self.thread = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons, capped = [], [], set(), 0
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
                    capped += name == "sw_power_cap"
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "samples_power_capped": capped}


def cpu_baseline(sample_crops: int, repeats: int, model: str = "cvit"):
    """The reference's CPU path (torch fp32, all host threads) on `sample_crops` crops of the same workload: the
    unmodified reference class when it was vendored (kind "reference"), else the oracle port (kind "port")."""
    import torch
    from fac_fake_b200 import weights as W
    from oracle import cvit_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    kind, where = "port", "oracle port"
    if model == "resvitkan":
        from oracle import resvitkan_oracle as R
        sd = W.make_resvitkan_state_dict(0, "default")
        fwd = lambda x: torch.cat([R.forward(x[i:i + 32], sd) for i in range(0, x.shape[0], 32)])   # noqa: E731
    elif model == "ggca":
        from oracle import ggca_oracle as G
        sd = W.make_ggca_state_dict(0, "default")
        fwd = lambda x: torch.cat([G.forward(x[i:i + 32], sd) for i in range(0, x.shape[0], 32)])   # noqa: E731
    else:
        f32, kind, where = reference_forward_fn()
        fwd = lambda x: torch.cat([f32(x[i:i + 32]) for i in range(0, x.shape[0], 32)])             # noqa: E731
    crops = W.synthetic_crops(sample_crops, seed=11)
    offs = list(range(0, sample_crops + 1, CROPS_PER_VIDEO))
    if offs[-1] != sample_crops:
        offs.append(sample_crops)

    def one(c=crops, o=offs):
        x = O.normalize_crops(c)
        lg = fwd(x)
        return O.video_scores(lg, o)

    one()                                         # warm-up
    times = []
    for _ in range(repeats):
        t0 = time.perf_counter()
        one()
        times.append(time.perf_counter() - t0)
    med = statistics.median(times)
    out = {"value": sample_crops / med, "unit": UNIT, "cores": cores, "kind": kind,
           "sample": f"{sample_crops} crops (chunks of 32) x {repeats} runs, median; torch {torch.__version__} fp32, {where}",
           "seconds_per_run": med}
    # BASELINE configs[0]: batch 15 = the 15 face crops of one sample clip (tests/golden/sample_clip_crops.npz)
    fx = os.path.join(ROOT, "tests", "golden", "sample_clip_crops.npz")
    if model == "cvit" and os.path.exists(fx):
        import numpy as np
        g = np.load(fx)
        real = torch.from_numpy(g["crops"][:15])
        one(real, [0, 15])
        ts = []
        for _ in range(max(2, repeats)):
            t0 = time.perf_counter()
            one(real, [0, 15])
            ts.append(time.perf_counter() - t0)
        m15 = statistics.median(ts)
        out["configs0_batch15"] = {"value": 15 / m15, "unit": UNIT, "videos_per_s": 1.0 / m15, "seconds_per_video": m15,
                                   "sample": "15 real face crops of one clip of sample__prediction_data (BASELINE configs[0]), one chunk"}
    return out


def s3d_flops_per_clip(T: int) -> int:
    """2*MAC of every convolution of S3D (model.py:17-34) for one [3,T,224,224] clip."""
    t1 = (T - 1) // 2 + 1
    t2 = (t1 - 1) // 2 + 1
    t3 = t2 // 2
    fl = 2 * T * 112 * 112 * 64 * 3 * 49 + 2 * t1 * 112 * 112 * 64 * 64 * 7                     # base.0 spatial + temporal
    fl += 2 * t1 * 56 * 56 * 64 * 64                                                            # base.2
    fl += 2 * t1 * 56 * 56 * 192 * (64 * 9 + 192 * 3)                                           # base.3
    mixed = [(192, 64, 96, 128, 16, 32, 32), (256, 128, 128, 192, 32, 96, 64), (480, 192, 96, 208, 16, 48, 64),
             (512, 160, 112, 224, 24, 64, 64), (512, 128, 128, 256, 24, 64, 64), (512, 112, 144, 288, 32, 64, 64),
             (528, 256, 160, 320, 32, 128, 128), (832, 256, 160, 320, 32, 128, 128), (832, 384, 192, 384, 48, 128, 128)]
    where = [(t1, 28)] * 2 + [(t2, 14)] * 5 + [(t3, 7)] * 2
    for (cin, b0, m1, o1, m2, o2, b3), (t, hw) in zip(mixed, where):
        vox = t * hw * hw
        fl += 2 * vox * (cin * (b0 + m1 + m2 + b3) + o1 * (m1 * 9 + o1 * 3) + o2 * (m2 * 9 + o2 * 3))
    return fl + 2 * 1024 * (t3 - 1)


def bind_to_gpu_numa_node(local_rank):
    """Pin this rank to the CPU cores NVML reports as local to its GPU, BEFORE the pinned host buffers of the e2e leg are
    allocated (first touch puts them on that NUMA node): at 8 ranks the 8 x 77 MB per step of H2D otherwise all cross
    from the node torchrun happened to start on.  Best effort: returns the number of cores, 0 when NVML / affinity
    are unavailable."""
    try:
        import pynvml
        import torch
        pynvml.nvmlInit()
        # NVML numbers the physical GPUs; CUDA_VISIBLE_DEVICES may renumber them for this process: go through the UUID
        props = torch.cuda.get_device_properties(local_rank)
        uuid = str(props.uuid)
        h = pynvml.nvmlDeviceGetHandleByUUID((uuid if uuid.startswith("GPU-") else "GPU-" + uuid).encode())
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cores = [64 * i + b for i, w in enumerate(words) for b in range(64) if (w >> b) & 1]
        allowed = sorted(set(cores) & set(os.sched_getaffinity(0)))
        if allowed:
            os.sched_setaffinity(0, allowed)
            return len(allowed)
    except Exception:
        pass
    return 0


def run_s3d(args):
    """--model s3d: BASELINE configs[4] — 64-frame 224x224 clips, 32 clips per GPU per step."""
    keep_stdout_clean()
    import torch
    import torch.distributed as dist
    from fac_fake_b200 import S3DEngine, weights as W
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    T, n = args.frames, args.clips
    warmup, steps = max(3, args.warmup), max(1, args.steps)
    peaks = read_peaks()
    eng = S3DEngine(1, "yes" if args.srm else "no", frames_per_clip=T, max_clips=n).to(dev).load_state_dict(
        W.make_s3d_state_dict(0, "default", srm=args.srm))
    ROT = 2                                           # 2 x 308 MB of uint8 clips > 126 MB L2; > 10 GB of activations per step
    host = [W.synthetic_clips(n, T, seed=200 + 13 * rank + i).pin_memory() for i in range(ROT)]
    devb = [b.to(dev) for b in host]

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step(i):
        lg = eng(devb[i % ROT])
        if world > 1:                                 # the only exchange: per-clip logits to every rank (n floats)
            out = [torch.empty_like(lg) for _ in range(world)]
            dist.all_gather(out, lg)
            lg = torch.cat(out)
        return lg

    for i in range(warmup):
        step(i)
    sync_all()
    l0 = eng.launch_count()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    e0.record()
    for i in range(steps):
        step(i)
    e1.record()
    sync_all()
    ms = e0.elapsed_time(e1)
    launches = eng.launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * n * steps / (ms_max * 1e-3)
    e2e_steps = max(1, min(steps, 5))
    for i in range(2):
        eng(host[i % ROT]).cpu()
    sync_all()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        eng(host[i % ROT]).cpu()                      # pinned host uint8 clips -> H2D -> forward -> logits D2H
    wall = time.perf_counter() - t0
    t = torch.tensor([wall], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * n * e2e_steps / float(t.item())
    if rank == 0:
        fl = s3d_flops_per_clip(T)
        if args.srm:      # HPF 3->30 5x5 on every frame, and the stem's spatial conv on 30 instead of 3 channels
            fl += 2 * T * 224 * 224 * 30 * 75 + 2 * T * 112 * 112 * 64 * 27 * 49
        tfl = value / world * fl / 1e12
        out = {
            "metric": "S3D clips/sec", "value": value, "unit": "clips/s", "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": ms_max / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic",
            "config": {"workload": f"S3D{' + SRM front-end' if args.srm else ''} bf16 inference, synthetic {T}-frame 224x224 uint8 clips, {n} clips per GPU per step "
                                   f"(BASELINE configs[4]), random-init weights", "clips_per_step_per_gpu": n, "frames_per_clip": T,
                       "parallelism": f"clip-sharded x{world} (no data-path collective)",
                       "l2": "inputs rotate over 2 batches (616 MB > 126 MB L2); > 10 GB of activations per step sweep L2",
                       "timing": "CUDA events on the launching stream, max over ranks"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "clips/s", "h2d_bytes_per_step": n * T * 224 * 224 * 3, "d2h_bytes_per_step": 4 * n,
                    "steps": e2e_steps, "api": "S3DEngine.forward on pinned host uint8 clips (H2D + ff_s3d_forward + logits D2H)"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "tensor", "kernel": "ff::rvk_conv2_kernel (all 61 convolutions) + stem; whole pass timed, so this is a lower bound for the kernels",
                         "achieved": tfl, "peak": peaks["bf16_sustained"], "unit": "TFLOP/s", "frac": tfl / peaks["bf16_sustained"],
                         "peak_source": f"bf16_tflops_sustained, {peaks['source']}", "traffic": None,
                         "algorithmic_flops_per_clip": fl, "frac_of_burst_peak": tfl / peaks["bf16_burst"]},
        }
        if world == 1 and not args.no_cpu_baseline:
            from oracle import s3d_oracle as S
            torch.set_num_threads(os.cpu_count() or 1)
            sd = W.make_s3d_state_dict(0, "default", srm=args.srm)
            x = W.synthetic_clips(1, T, seed=3).permute(0, 4, 1, 2, 3).contiguous().float()
            S.forward(x, sd, srm=args.srm)
            t0 = time.perf_counter()
            for _ in range(2):
                S.forward(x, sd, srm=args.srm)
            dt = (time.perf_counter() - t0) / 2
            out["cpu_baseline"] = {"value": 1.0 / dt, "unit": "clips/s", "cores": os.cpu_count(), "kind": "port",
                                   "sample": f"1 clip of {T} frames x 2 runs; torch {torch.__version__} fp32 CPU oracle", "seconds_per_run": dt}
        emit(out)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


BLAZE_BLOCKS = ((24, 24, 1, 64), (24, 28, 1, 64), (28, 32, 2, 64), (32, 36, 1, 32), (36, 42, 1, 32), (42, 48, 2, 32), (48, 56, 1, 16),
          (56, 64, 1, 16), (64, 72, 1, 16), (72, 80, 1, 16), (80, 88, 1, 16), (88, 96, 2, 16), (96, 96, 1, 8), (96, 96, 1, 8),
          (96, 96, 1, 8), (96, 96, 1, 8))


def blazeface_work_per_tile():
    """(2*MAC flops, fp32 activation bytes moved by the 19 launches: each reads its input and writes its output once)."""
    fl = 2 * 64 * 64 * 24 * 75
    by = 128 * 128 * 3 + 64 * 64 * 24 * 4
    for cin, cout, s, hw in BLAZE_BLOCKS:
        ho = hw // s
        fl += 2 * ho * ho * cin * 9 + 2 * ho * ho * cin * cout
        by += 4 * (hw * hw * cin + ho * ho * cout)
    fl += 2 * (256 * 88 * 34 + 64 * 96 * 102)
    by += 4 * (256 * 88 + 64 * 96 + 896 * 17 * 2 + 896 * 17)
    return fl, by


def run_blazeface(args):
    """--model blazeface: SURVEY 8f-3, tiles/s of the detector (network + decode on the GPU, blending NMS on the host)."""
    keep_stdout_clean()
    import numpy as np
    import torch
    from fac_fake_b200 import BlazeFaceEngine
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
    # the whole front-end of SURVEY 8f-3: frames -> tiles -> detector -> untile / NMS / margin -> 224x224 crops -> CViT scores
    from fac_fake_b200 import CViTEngine, FaceExtractorEngine, weights as W
    fx = np.load(os.path.join(ROOT, "tests", "golden", "face_extract.npz"))
    n_frames = 60
    fr_host = torch.from_numpy(np.concatenate([fx["land_frames"]] * (n_frames // len(fx["land_frames"])))).pin_memory()
    small = BlazeFaceEngine(max_tiles=3 * n_frames).to("cuda:0")
    small.load_weights(sd)
    small.load_anchors(w["anchors"])
    ex = FaceExtractorEngine(None, small)
    model = CViTEngine(max_crops=n_frames * 2).to("cuda:0").load_state_dict(W.make_state_dict(0, "bn"))

    def front_end(frames_host):
        frames = frames_host.cuda(non_blocking=True)
        crops = ex.extract_crops_device(frames, model)
        return model.predict_videos(crops, [0, crops.shape[0]]).cpu(), crops.shape[0]
    for _ in range(3):
        front_end(fr_host)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        _, n_crops = front_end(fr_host)
    fe_s = (time.perf_counter() - t0) / args.steps
    fl, by = blazeface_work_per_tile()
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
    gbs = by * args.tiles / (ms * 1e-3) / 1e9
    emit({
        "metric": "BlazeFace tiles/sec", "value": args.tiles / ms * 1e3, "unit": "tiles/s", "ms_per_step": ms, "tiles_per_step": args.tiles,
        "gpu_launches_per_step": int(launches), "dtype": "f32", "faces_found": int(sum(len(f) for f in faces)),
        "e2e": {"value": args.tiles / e2e_s, "unit": "tiles/s", "what": "pinned host uint8 tiles -> H2D -> network + decode -> mask + blending NMS on the device (ff_blazeface_nms) -> D2H of [n,16,17] faces + counts"},
        "face_front_end": {"value": n_frames / fe_s, "unit": "frames/s", "frames": n_frames, "frame_hw": list(fr_host.shape[1:3]), "crops": int(n_crops),
                           "what": "pinned host uint8 frames (536x500, 3 tiles each) -> H2D -> ff_blazeface_tile_frames -> detector -> ff_blazeface_frame_faces "
                                   "-> D2H of detections/rectangles -> ff_preprocess_crops on device views -> CViT -> per-video score on the host"},
        "roofline": {"bound": "hbm", "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": gbs / peaks["hbm_gbs"],
                     "algorithmic_bytes_per_tile": by, "flops_per_tile": fl, "gflops": fl * args.tiles / (ms * 1e-3) / 1e9},
        "cpu_baseline": {"value": len(sample) / cpu_s, "unit": "tiles/s", "cores": os.cpu_count(), "kind": "port",
                         "sample": f"{len(sample)} tiles x 5 runs, torch fp32 oracle incl. NMS"},
    })




def load_reference_class():
    """The UNMODIFIED reference module model/cvit.py, vendored by __graft_entry__.build() into baseline/_ref/ (git-ignored,
    travels to the GPU box with the snapshot).  Returns (CViT class, where) or (None, why)."""
    ref_dir = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.exists(os.path.join(ref_dir, "cvit.py")):
        return None, "baseline/_ref/cvit.py not present (build() vendors it when /root/reference is mounted)"
    import importlib.util
    spec = importlib.util.spec_from_file_location("ff_reference_cvit", os.path.join(ref_dir, "cvit.py"))
    mod = importlib.util.module_from_spec(spec)
    try:
        spec.loader.exec_module(mod)
    except Exception as e:  # einops missing etc.
        return None, f"import of baseline/_ref/cvit.py failed: {e!r}"
    return mod.CViT, "baseline/_ref/cvit.py (unmodified /root/reference/CViT-main/model/cvit.py)"


def reference_forward_fn():
    """(callable x[n<=32,3,224,224] fp32 -> logits, kind, description): the reference's own class when vendored, else the port."""
    import torch
    from fac_fake_b200 import weights as W
    from oracle import cvit_oracle as O
    sd = W.make_state_dict(0, "default")
    cls, where = load_reference_class()
    if cls is None:
        return (lambda x: O.forward(x, sd)), "port", f"oracle/cvit_oracle.py ({where})"
    model = cls(image_size=224, patch_size=7, num_classes=2, channels=512, dim=1024, depth=6, heads=8, mlp_dim=2048)
    model.load_state_dict(sd)
    model.eval()

    def fwd(x):
        with torch.no_grad():
            return model(x)
    return fwd, "reference", where


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path — the unmodified `cvit.CViT` class under
    torch.no_grad() on all host cores, fed exactly like cvit_prediction.py:209-242 feeds it (normalised fp32 NCHW chunks
    of <= 32 crops, pred_sig + pre_process_prediction per video).  Each step is a bounded sample of configs[1]."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample = 32
    import torch  # noqa: F401
    from fac_fake_b200 import weights as W
    from oracle import cvit_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    fwd, kind, where = reference_forward_fn()
    crops = W.synthetic_crops(sample, seed=11)
    offs = [0, sample]

    def one():
        x = O.normalize_crops(crops)
        lg = torch.cat([fwd(x[i:i + 32]) for i in range(0, sample, 32)])
        return O.video_scores(lg, offs)

    for _ in range(max(1, min(args.warmup, 2))):
        one()
    steps = max(1, min(args.steps, 10))
    t0 = time.perf_counter()
    for _ in range(steps):
        one()
    dt = time.perf_counter() - t0
    val = sample * steps / dt
    out = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": max(1, min(args.warmup, 2)), "ms_per_step": 1e3 * dt / steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "CViT inference, synthetic 224x224 uint8 face crops, bounded CPU sample of BASELINE configs[1]",
                   "crops_per_step": sample, "videos_per_s_30f": val / 30.0, "reference_code": where},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": f"{sample} crops per step (one reference-sized chunk), {steps} steps; torch fp32 on host cores"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(out)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--crops", type=int, default=CROPS_PER_STEP)
    ap.add_argument("--tiles", type=int, default=512, help="--model blazeface: 128x128 tiles per step")
    ap.add_argument("--clips", type=int, default=32, help="--model s3d: clips per GPU per step")
    ap.add_argument("--frames", type=int, default=64, help="--model s3d: frames per clip")
    ap.add_argument("--srm", action="store_true", help="--model s3d: S3D(num_class, 'yes'), the SRM high-pass front-end")
    ap.add_argument("--model", default="cvit", choices=sorted(MODELS) + ["s3d", "blazeface"],
                    help="cvit = the north-star path (default, what the driver runs); resvitkan = SURVEY 8f-1 / BASELINE configs[3]")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-config2", action="store_true", help="skip the configs[2] strong-scaling block (8192 videos x 30 frames)")
    ap.add_argument("--config2-videos", type=int, default=8192)
    ap.add_argument("--e2e-steps", type=int, default=0, help="default: same as --steps")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    if args.model == "s3d":
        run_s3d(args)
        return
    if args.model == "blazeface":
        run_blazeface(args)
        return

    keep_stdout_clean()
    import torch
    import torch.distributed as dist
    from fac_fake_b200 import CViTEngine, CViTGGCAEngine, ResVitKanEngine, weights as W
    from fac_fake_b200.sharding import gather_scores, shard_range
    model = MODELS[args.model]
    rvk = args.model == "resvitkan"

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa_cores = bind_to_gpu_numa_node(local_rank) if world > 1 else 0
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    warmup = max(3, args.warmup)
    steps = max(1, args.steps)
    n = args.crops
    peaks = read_peaks()

    cap = min(model["max_crops"], (n + 31) // 32 * 32)
    if rvk:
        eng = ResVitKanEngine(max_crops=cap).to(dev).load_state_dict(W.make_resvitkan_state_dict(0, "default"))
    elif args.model == "ggca":
        eng = CViTGGCAEngine(max_crops=cap).to(dev).load_state_dict(W.make_ggca_state_dict(0, "default"))
    else:
        eng = CViTEngine(max_crops=cap).to(dev).load_state_dict(W.make_state_dict(0, "default"))
    offsets = list(range(0, n + 1, CROPS_PER_VIDEO))
    if offsets[-1] != n:
        offsets.append(n)
    n_videos = len(offsets) - 1
    # rotating inputs: 4 distinct batches (4 x 77 MB > 126 MB L2); activations (>1.6 GB / step) also sweep L2
    ROT = 4
    host_batches = [W.synthetic_crops(n, seed=100 + 17 * rank + i).pin_memory() for i in range(ROT)]
    dev_batches = [b.to(dev) for b in host_batches]
    torch.cuda.synchronize()

    # per-rank score store: every step's scores stay on the device; ONE gather after the last step
    score_store = torch.empty((max(warmup, steps), n_videos), dtype=torch.float32, device=dev)

    def step(i):
        score_store[i] = eng.predict_videos(dev_batches[i % ROT], offsets)

    def gather_all(k):
        flat = score_store[:k].reshape(-1)
        return gather_scores(flat, flat.numel() * world, rank, world) if world > 1 else flat

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(warmup):
        step(i)
    gather_all(warmup)
    sync_all()
    l0 = eng.launch_count()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # ---- timed region A: K steps + the one score gather, no instrumentation -> `value`
    sync_all()
    e0.record()
    for i in range(steps):
        step(i)
    all_scores = gather_all(steps)
    e1.record()
    sync_all()
    ms = e0.elapsed_time(e1)
    launches = eng.launch_count() - l0
    assert bool(torch.isfinite(all_scores).all())
    # ---- timed region B: the same K steps with a CUDA-event pair around every kernel launch (on the launching
    #      stream) -> per-layer device time for the rooflines.  The event records serialise the launches
    #      (no programmatic-dependent-launch overlap), so this pass is slower than A and is NOT the reported value.
    eng.set_profiling(True)
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for i in range(steps):
        step(i)
    f1.record()
    sync_all()
    ms_instr = f0.elapsed_time(f1)
    clocks = sampler.stop() if rank == 0 else None
    prof = eng.get_profile()
    prof_layers = eng.get_profile(per_layer=True)
    eng.set_profiling(False)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * n * steps / (ms_max * 1e-3)

    # ---- e2e: host uint8 crops -> H2D -> forward -> scores D2H, through the C-ABI host entry point
    e2e_steps = args.e2e_steps or steps
    for i in range(3):
        eng.predict_videos_host(host_batches[i % ROT], offsets)
    sync_all()
    t0 = time.perf_counter()
    e0.record()
    for i in range(e2e_steps):
        eng.predict_videos_host(host_batches[i % ROT], offsets)
    e1.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    ms_e2e = max(e0.elapsed_time(e1), wall * 1e3)
    t = torch.tensor([ms_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * n * e2e_steps / (float(t.item()) * 1e-3)

    # ---- BASELINE configs[2] as written: 8192 videos x 30 uint8 frames, per-video seed = video id,
    #      video_offsets = 30 * arange, slot = frame index, rank r owns videos [r*V/W, (r+1)*V/W); STRONG scaling: the
    #      total is fixed, one gather of the V scores at the end.  Inputs are generated on the device before the timed
    #      region (37 GB of uint8 over all ranks), frames of a video are contiguous.
    config2 = None
    if args.model == "cvit" and not args.no_config2:
        V, FR = args.config2_videos, 30
        lo, hi = shard_range(V, rank, world)
        nv_local = hi - lo
        store = torch.empty((nv_local * FR, 224, 224, 3), dtype=torch.uint8, device=dev)
        gen = torch.Generator(device=dev)
        for v in range(lo, hi):
            gen.manual_seed(v)
            torch.randint(0, 256, (FR, 224, 224, 3), dtype=torch.uint8, device=dev, generator=gen, out=store[(v - lo) * FR:(v - lo + 1) * FR])
        CH = 256                                   # videos per call = 7680 crops = 15 full passes of 512
        local_scores = torch.empty((nv_local,), dtype=torch.float32, device=dev)

        def run_config2():
            for c0 in range(0, nv_local, CH):
                c1 = min(nv_local, c0 + CH)
                local_scores[c0:c1] = eng.predict_videos(store[c0 * FR:c1 * FR], [FR * k for k in range(c1 - c0 + 1)])
            return gather_scores(local_scores, V, rank, world)

        eng.predict_videos(store[:FR * min(nv_local, 34)], [FR * k for k in range(min(nv_local, 34) + 1)])     # warm-up
        sync_all()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        sc = run_config2()
        g1.record()
        sync_all()
        t = torch.tensor([g0.elapsed_time(g1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms2 = float(t.item())
        config2 = {"what": f"BASELINE configs[2]: {V} synthetic videos x {FR} uint8 frames (per-video seed), sharded by whole videos over {world} GPU(s), "
                           "slot = frame index, one gather of the scores at the end",
                   "scaling": "strong", "videos": V, "frames_per_video": FR, "videos_per_s": V / (ms2 * 1e-3),
                   "crops_per_s": V * FR / (ms2 * 1e-3), "seconds": ms2 * 1e-3, "n_gpus": world,
                   "scores_finite": bool(torch.isfinite(sc).all()), "scores_gathered": int(sc.numel())}
        del store

    if rank == 0 and rvk:
        fl, by = resvitkan_trunk_work()
        conv_ms, conv_launches = prof["tcgen05_conv"]
        stem_ms = prof["conv1"][0]
        trunk_ms = conv_ms + stem_ms
        gbs = by * n * steps / (trunk_ms * 1e-3) / 1e9 if trunk_ms > 0 else 0.0
        out = {
            "metric": model["metric"], "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": ms_max / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"ResVitKan bf16 inference, synthetic batch of {n} uint8 face crops 224x224 per GPU "
                                   f"(BASELINE {model['baseline_config']}; {n_videos} videos x {CROPS_PER_VIDEO} crops, passes of {cap}), random-init weights",
                       "crops_per_step_per_gpu": n, "parallelism": f"video-sharded x{world} (no data-path collective)",
                       "l2": "inputs rotate over 4 distinct batches; > 10 GB of activations per step sweep L2",
                       "timing": "CUDA events on the launching stream, max over ranks"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": n * 224 * 224 * 3 + 4 * (n_videos + 1),
                    "d2h_bytes_per_step": 4 * n_videos, "steps": e2e_steps,
                    "api": "ff_cvit_predict_host on an ff_resvitkan_create handle (ResVitKanEngine.predict_videos_host)"},
            "gpu_launches": int(launches),
            "roofline": {
                "bound": "hbm", "kernel": "ResNet-50 trunk: ff::rvk_conv_kernel (52 bottleneck convolutions + channel conv) + stem kernels",
                "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": gbs / peaks["hbm_gbs"],
                "peak_source": f"hbm_gbs, {peaks['source']}", "traffic": None,
                "algorithmic_bytes_per_crop": by, "algorithmic_flops_per_crop": fl,
                "trunk_tflops": fl * n * steps / (trunk_ms * 1e-3) / 1e12 if trunk_ms > 0 else 0.0,
                "kernel_ms_per_step": trunk_ms / steps, "kernel_launches_per_step": (conv_launches + prof["conv1"][1]) / steps,
                "step_share": trunk_ms / ms_instr if ms_instr > 0 else None,
                "how": "second pass of the same K steps with a CUDA-event pair around every launch; unfused per-layer bf16 activation bytes (each conv reads its input and writes its output once)",
                "by_class_ms_per_step": {k: v[0] / steps for k, v in prof.items()},
            },
        }
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline(32, 2, "resvitkan")
        emit(out)
    elif rank == 0:
        conv_ms, conv_launches = prof["tcgen05_conv"]
        flops_per_crop = model["flops_per_crop"]
        # Which peak: ALWAYS the burst figure (the larger, conservative denominator).  The sustained cuBLAS figure of
        # MEASURED_PEAKS.json (1393 TFLOP/s) is below what the CTA-pair conv kernels reach inside this very step under the power
        # cap (1.50-1.56 PFLOP/s), so a fraction of it would read > 1; it is kept as `frac_of_sustained_peak` for information.
        peak = peaks["bf16_burst"]
        peak_name = "bf16_tflops (burst)"
        # per-layer event times -> one roofline per kernel family.  Profile slot k (k >= 1) = feature layer k+1; slot 1 also
        # carries layer 1 (layers 1+2 are one kernel, ff::c12_kernel).
        plan = [(3, 32, 224), (32, 32, 224), (32, 32, 224), (32, 64, 112), (64, 64, 112), (64, 64, 112), (64, 128, 56), (128, 128, 56),
                (128, 128, 56), (128, 256, 28), (256, 256, 28), (256, 256, 28), (256, 256, 28), (256, 512, 14), (512, 512, 14),
                (512, 512, 14), (512, 512, 14)]
        lf = [2 * 9 * ci * co * hw * hw for ci, co, hw in plan]
        lt = list(prof_layers.values())
        fams = {"ff::c12_kernel (layers 1+2 fused)": ([1], lf[0] + lf[1]), "ff::ws2conv_kernel (layers 3, 4)": ([2, 3], lf[2] + lf[3]),
                "ff::ws2x_conv_kernel (layers 5, 6)": ([4, 5], lf[4] + lf[5]), "ff::ptcw_conv_kernel (layers 7-9)": ([6, 7, 8], sum(lf[6:9])),
                "ff::ptc2_conv_kernel (layers 10-17)": (list(range(9, 17)), sum(lf[9:17]))}
        by_kernel = {}
        for name, (slots, fl) in fams.items():
            kms = sum(lt[k][0] for k in slots)
            kl = sum(lt[k][1] for k in slots)
            tf = fl * n * steps / (kms * 1e-3) / 1e12 if kms > 0 else 0.0
            by_kernel[name] = {"ms_per_step": kms / steps, "launches_per_step": kl / steps, "tflops": tf, "frac": tf / peak,
                               "flops_per_crop": fl}
        gemm_ms = prof["tcgen05_gemm"][0]
        by_kernel["ff::xf_kernel + ff::tc_gemm_kernel (embed, encoder, head)"] = {
            "ms_per_step": gemm_ms / steps, "launches_per_step": prof["tcgen05_gemm"][1] / steps,
            "tflops": (FLOPS_PER_CROP - CONV_FLOPS) * n * steps / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0,
            "flops_per_crop": FLOPS_PER_CROP - CONV_FLOPS}
        by_kernel[list(by_kernel)[-1]]["frac"] = by_kernel[list(by_kernel)[-1]]["tflops"] / peak
        dom_name = max(fams, key=lambda k: by_kernel[k]["ms_per_step"])
        dom = by_kernel[dom_name]
        # traffic of the dominant kernel: dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu
        # --set full capture of THIS kernel (profiles/r02_roofline_traffic.json, written by tools/ncu_summary.py)
        traffic, traffic_src = None, "no committed ncu capture for this kernel"
        tf_path = os.path.join(ROOT, "profiles", "r02_roofline_traffic.json")
        if os.path.exists(tf_path):
            tj = json.load(open(tf_path))
            for key, ent in tj.items():
                if key in dom_name:
                    traffic, traffic_src = ent["dram_bytes_per_launch"], ent["source"]
        class_tflops = (CONV_FLOPS * n * steps) / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else 0.0
        roofline = {
            "bound": "tensor", "kernel": dom_name,
            "achieved": dom["tflops"], "peak": peak, "unit": "TFLOP/s", "frac": dom["tflops"] / peak,
            "peak_source": f"{peak_name}, {peaks['source']}; power cap on {clocks.get('samples_power_capped') if clocks else None} of "
                           f"{clocks.get('samples') if clocks else None} clock samples",
            "traffic": traffic, "traffic_source": traffic_src,
            "algorithmic_flops_per_launch": dom["flops_per_crop"] * n / max(dom["launches_per_step"], 1.0),
            "avg_launch_ms": dom["ms_per_step"] / max(dom["launches_per_step"], 1.0),
            "kernel_ms_per_step": dom["ms_per_step"], "kernel_launches_per_step": dom["launches_per_step"],
            "step_share": dom["ms_per_step"] * steps / ms_instr if ms_instr > 0 else None,
            "how": "second pass of the same K steps with a CUDA-event pair around every launch on the launching stream; "
                   "value/ms_per_step come from the uninstrumented pass",
            "by_kernel": by_kernel,
            "conv_stack": {"tflops": class_tflops, "frac": class_tflops / peak, "frac_of_burst_peak": class_tflops / peaks["bf16_burst"],
                           "frac_of_sustained_peak": class_tflops / peaks["bf16_sustained"], "ms_per_step": conv_ms / steps,
                           "launches_per_step": conv_launches / steps, "algorithmic_flops_per_crop": CONV_FLOPS},
            "whole_step": {"tflops": value / world * flops_per_crop / 1e12, "frac_of_burst_peak": value / world * flops_per_crop / 1e12 / peaks["bf16_burst"],
                           "frac_of_sustained_peak": value / world * flops_per_crop / 1e12 / peaks["bf16_sustained"]},
            "instrumented_ms_per_step": ms_instr / steps,
            "by_class_ms_per_step": {k: v[0] / steps for k, v in prof.items()},
        }
        out = {
            "metric": model["metric"], "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": ms_max / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"{model['name']} bf16 inference, synthetic batch of {n} uint8 face crops 224x224 per GPU "
                                   f"({n_videos} videos x {CROPS_PER_VIDEO} crops, slot = i % 32; BASELINE {model['baseline_config']}), random-init weights",
                       "crops_per_step_per_gpu": n, "videos_per_step_per_gpu": n_videos,
                       "videos_per_s_30f": value / 30.0,
                       "parallelism": f"video-sharded x{world} (no data-path collective; one gather of the per-video scores after the last step, inside the timed region)",
                       "l2": "inputs rotate over 4 distinct batches (308 MB > 126 MB L2); >1.6 GB of activations per step sweep L2",
                       "timing": "CUDA events on the launching stream, max over ranks",
                       "host": f"rank bound to the {numa_cores} cores NVML reports local to its GPU" if numa_cores else "no CPU binding"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": n * 224 * 224 * 3 + 4 * (n_videos + 1),
                    "d2h_bytes_per_step": 4 * n_videos, "steps": e2e_steps,
                    "api": "ff_cvit_predict_host (CViTEngine.predict_videos_host), pinned host uint8 crops"},
            "gpu_launches": int(launches),
            "roofline": roofline,
        }
        if config2 is not None:
            out["config2"] = config2
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline(32, 3, args.model)
        emit(out)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
