"""Generate tests/golden/*.npz by running the UNMODIFIED reference here (build container only).

    python oracle/make_golden.py

Imports ``cvit.CViT`` from /root/reference/CViT-main/model (cvit.py:80-179), loads the
deterministic synthetic state_dicts of ``fac_fake_b200.weights`` and records the
reference's own outputs on seeded inputs.  /root/reference does not exist on the GPU
box, so the vectors are committed; the oracle and the CUDA engine are both checked
against them.  The per-video reduction vectors are produced with the reference's
``pred_sig`` / ``pre_process_prediction`` source text (cvit_prediction.py:258-281),
exec'd verbatim from the reference file because that script is not importable
(missing facenet_pytorch / face_recognition, os.chdir to a Windows path).
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference/CViT-main"
sys.path.insert(0, os.path.join(REF, "model"))

from cvit import CViT  # noqa: E402  (reference class)
from fac_fake_b200 import weights as W  # noqa: E402
from oracle import cvit_oracle as O  # noqa: E402


def reference_reduction_functions():
    """exec the reference's pred_sig / pre_process_prediction source lines verbatim."""
    lines = open(os.path.join(REF, "cvit_prediction.py"), encoding="utf-8").read().splitlines()
    src = "\n".join(lines[257:260] + [""] + lines[265:282])     # :258-260 and :266-282 (1-based)
    ns = {"torch": torch}
    exec(src, ns)
    return ns["pred_sig"], ns["pre_process_prediction"]


def main():
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    torch.set_num_threads(8)
    for variant in ("default", "bn"):
        sd = W.make_state_dict(0, variant)
        model = CViT(image_size=224, patch_size=7, num_classes=2, channels=512,
                     dim=1024, depth=6, heads=8, mlp_dim=2048).eval()
        model.load_state_dict(sd, strict=True)
        crops = W.synthetic_crops(40, seed=1)
        x = O.normalize_crops(crops)
        with torch.no_grad():
            logits_a = model(x[0:32])                  # slots 0..31
            logits_b = model(x[32:40])                 # slots 0..7 (second chunk of a 40-crop video)
            # intermediate checkpoints of the first 4 crops
            feats = {}
            h = x[0:4]
            conv_no = 0
            for idx, mod in enumerate(model.features):
                h = mod(h)
                if isinstance(mod, torch.nn.ReLU):
                    nxt = model.features[idx + 1] if idx + 1 < len(model.features) else None
                    if not isinstance(nxt, torch.nn.MaxPool2d):
                        feats[conv_no] = h
                        conv_no += 1
                elif isinstance(mod, torch.nn.MaxPool2d):
                    feats[conv_no] = h
                    conv_no += 1
            assert conv_no == 17
        layer_stats = np.stack([np.array([feats[i].double().mean().item(), feats[i].double().abs().mean().item(),
                                          feats[i].double().pow(2).mean().sqrt().item()]) for i in range(17)])
        np.savez_compressed(
            os.path.join(out_dir, f"cvit_logits_{variant}.npz"),
            logits=torch.cat([logits_a, logits_b]).numpy(),
            layer_stats=layer_stats,
            feat_final=feats[16].numpy().astype(np.float32),        # [4,512,7,7]
            feat_l2_sample=feats[2][0, :, :8, :8].numpy(),           # crop 0, 32ch, 8x8 corner after pool
            seed_weights=0, seed_crops=1, n=40,
        )
        print(variant, "logits[0:3] =", logits_a[0:3].tolist())

    # --- per-video reduction vectors from the reference's own source text
    pred_sig, pre_process_prediction = reference_reduction_functions()
    g = torch.Generator().manual_seed(5)
    cases, lens, scores = [], [], []
    for n in (1, 2, 3, 4, 15, 29, 30, 32, 33, 64, 90):
        for scale, shift in ((1.0, 0.0), (3.0, 1.5), (3.0, -1.5), (0.01, 0.0)):
            lg = torch.randn((n, 2), generator=g) * scale
            lg[:, 0] += shift
            s = pre_process_prediction(pred_sig(lg))
            cases.append(lg.numpy())
            lens.append(n)
            scores.append(float(s.item()))
    np.savez_compressed(os.path.join(out_dir, "video_reduction.npz"),
                        logits=np.concatenate(cases, 0), lens=np.array(lens), scores=np.array(scores, dtype=np.float64))
    print("reduction cases:", len(lens))


if __name__ == "__main__":
    main()


def main_resvitkan():
    """tests/golden/resvitkan_{default,bn}.npz from the reference ResVitKan class (ResVitKan/ResVitKan.py:284-329)."""
    sys.path.insert(0, os.path.join(REF, "ResVitKan"))
    from ResVitKan import CViT as RVK  # noqa: E402  (reference class; imports kan.KAN from its own directory)
    from oracle import resvitkan_oracle as R  # noqa: E402
    out_dir = os.path.join(ROOT, "tests", "golden")
    for variant in ("default", "bn"):
        sd = W.make_resvitkan_state_dict(0, variant)
        model = RVK().eval()
        model.load_state_dict(sd, strict=True)
        crops = W.synthetic_crops(8, seed=2)
        x = O.normalize_crops(crops)
        with torch.no_grad():
            logits = model(x)
            f = model.features
            h = f.maxpool(f.relu(f.bn1(f.conv1(x[:2]))))
            stats = [h]
            for layer in (f.layer1, f.layer2, f.layer3, f.layer4):
                h = layer(h)
                stats.append(h)
            h = f.bn2(f.channel(h))
            stats.append(h)
        layer_stats = np.stack([np.array([s.double().mean().item(), s.double().abs().mean().item(),
                                          s.double().pow(2).mean().sqrt().item()]) for s in stats])
        np.savez_compressed(os.path.join(out_dir, f"resvitkan_{variant}.npz"), logits=logits.numpy(), layer_stats=layer_stats,
                            feat_final=stats[-1].numpy().astype(np.float32), stem_sample=stats[0][0, :, :6, :6].numpy(),
                            seed_weights=0, seed_crops=2, n=8)
        print("resvitkan", variant, "logits[0:2] =", logits[0:2].tolist())


if __name__ == "__main__" and os.environ.get("FF_GOLDEN_RESVITKAN", "1") == "1":
    main_resvitkan()


def load_reference_ggca_class():
    """Import `cvit_GGCA_ADD_DEConv_RepBn8.CViT` on a CPU-only host.

    The reference file is CUDA-only in three places that do not change the arithmetic: DEConv's ``get_weight`` builds
    its folded kernels with ``torch.cuda.FloatTensor(...)`` (:227,297,318), ``Conv2d_vd.__init__`` calls
    ``self.conv.cuda()`` (:313), and the module imports ``torchsummary`` (:7).  Shims: ``torch.cuda.FloatTensor`` ->
    ``torch.FloatTensor``, ``nn.Module.cuda`` -> identity, an empty ``torchsummary`` module.  A fourth, layout-only
    shim is applied per instance by ``make_ggca_model``: on this CPU build the conv stack returns a channels_last
    tensor, on which GGCA's ``x.view(b*groups, ...)`` (:181) raises; a forward pre-hook makes it contiguous (on CUDA
    the reference gets a contiguous NCHW tensor and needs no hook).  Everything else is the reference's own code.
    """
    import types
    sys.modules.setdefault("torchsummary", types.SimpleNamespace(summary=lambda *a, **k: None))
    torch.cuda.FloatTensor = torch.FloatTensor
    torch.nn.Module.cuda = lambda self, device=None: self
    sys.path.insert(0, os.path.join(REF, "model"))
    import importlib
    return importlib.import_module("cvit_GGCA_ADD_DEConv_RepBn8").CViT


def make_ggca_model(sd):
    model = load_reference_ggca_class()().eval()
    model.load_state_dict(sd, strict=True)
    model.ggca.register_forward_pre_hook(lambda mod, args: (args[0].contiguous(),))
    return model


def main_ggca():
    """tests/golden/ggca_{default,bn}.npz from the reference class (model/cvit_GGCA_ADD_DEConv_RepBn8.py:353-455)."""
    out_dir = os.path.join(ROOT, "tests", "golden")
    for variant in ("default", "bn"):
        sd = W.make_ggca_state_dict(0, variant)
        model = make_ggca_model(sd)
        crops = W.synthetic_crops(8, seed=3)
        x = O.normalize_crops(crops)
        with torch.no_grad():
            logits = model(x)
            f = model.features2(model.features1(x[:2]))
            g = f * model.ggca(f)
        stats = np.stack([np.array([s.double().mean().item(), s.double().abs().mean().item(), s.double().pow(2).mean().sqrt().item()])
                          for s in (f, g)])
        np.savez_compressed(os.path.join(out_dir, f"ggca_{variant}.npz"), logits=logits.numpy(), feat_stats=stats,
                            gated_sample=g[0, :16].numpy().astype(np.float32), seed_weights=0, seed_crops=3, n=8)
        print("ggca", variant, "logits[0:2] =", logits[0:2].tolist(), "feat stats", stats.tolist())


if __name__ == "__main__" and os.environ.get("FF_GOLDEN_GGCA", "1") == "1":
    main_ggca()


def main_blazeface():
    """tests/golden/blazeface_{weights,golden}.npz: the reference's shipped detector weights / anchors
    (helpers/blazeface.pth, helpers/anchors.npy) and outputs of the reference BlazeFace class (helpers/blazeface.py) on
    tiles cut from the reference's sample videos exactly as FaceExtractor._tile_frames does
    (helpers/helpers_face_extract_1.py:139-204: three min(H,W) windows per landscape frame, cv2 INTER_AREA to 128x128)."""
    import cv2
    helpers = os.path.join(REF, "helpers")
    sys.path.insert(0, helpers)
    from blazeface import BlazeFace  # noqa: E402  (reference class)
    net = BlazeFace()
    net.load_weights(os.path.join(helpers, "blazeface.pth"))
    net.load_anchors(os.path.join(helpers, "anchors.npy"))
    out_dir = os.path.join(ROOT, "tests", "golden")
    sd = {k: v.numpy() for k, v in net.state_dict().items()}
    np.savez_compressed(os.path.join(out_dir, "blazeface_weights.npz"), anchors=net.anchors.numpy(), **sd)
    tiles = []
    for name, frame_ids in (("aajsqyyjni.mp4", (0, 40)), ("sample_2.mp4", (10,)), ("0017_fake.mp4.mp4", (5,))):
        cap = cv2.VideoCapture(os.path.join(REF, "sample__prediction_data", name))
        for fid in frame_ids:
            cap.set(cv2.CAP_PROP_POS_FRAMES, fid)
            ok, frame = cap.read()
            if not ok:
                continue
            frame = cv2.cvtColor(frame, cv2.COLOR_BGR2RGB)
            H, W, _ = frame.shape
            split = min(H, W)
            x_step = (W - split) // 2
            for t in range(3 if W > H else 1):
                crop = frame[0:split, t * x_step:t * x_step + split, :]
                tiles.append(cv2.resize(crop, (128, 128), interpolation=cv2.INTER_AREA))
        cap.release()
    tiles = np.stack(tiles)
    with torch.no_grad():
        x = net._preprocess(torch.from_numpy(tiles).permute(0, 3, 1, 2))
        r, c = net(x)
    det = net.predict_on_batch(tiles, apply_nms=False)
    faces = net.nms(det)
    counts = np.array([len(d) for d in det]), np.array([len(f) for f in faces])
    np.savez_compressed(os.path.join(out_dir, "blazeface_golden.npz"), tiles=tiles, raw_scores=c.numpy()[..., 0],
                        raw_boxes=r.numpy().astype(np.float32), det_counts=counts[0], face_counts=counts[1],
                        faces=np.concatenate([f.numpy() for f in faces]) if sum(counts[1]) else np.zeros((0, 17), np.float32))
    print("blazeface: tiles", tiles.shape, "detections per tile", counts[0].tolist(), "faces per tile", counts[1].tolist())


if __name__ == "__main__" and os.environ.get("FF_GOLDEN_BLAZEFACE", "1") == "1":
    main_blazeface()


def load_reference_s3d_class():
    s3d_dir = os.path.join(os.path.dirname(REF), "sx_exp_deepfakedetect-master", "S3D")
    if not os.path.isdir(s3d_dir):
        s3d_dir = "/root/reference/sx_exp_deepfakedetect-master/S3D"
    sys.path.insert(0, s3d_dir)
    import importlib
    return importlib.import_module("model").S3D


def clips_to_reference_input(clips_u8):
    """uint8 [b,T,H,W,3] -> float [b,3,T,H,W] raw 0..255 (S3D-test.py:94-96)."""
    return clips_u8.permute(0, 4, 1, 2, 3).contiguous().float()


def main_s3d():
    """tests/golden/s3d_{default,bn}.npz from the reference S3D class (S3D/model.py), 2 clips x 16 frames x 224x224."""
    from oracle import s3d_oracle as S  # noqa: E402
    S3D = load_reference_s3d_class()
    out_dir = os.path.join(ROOT, "tests", "golden")
    for variant in ("default", "bn"):
        sd = W.make_s3d_state_dict(0, variant)
        model = S3D(1, "no").eval()
        model.load_state_dict(sd, strict=True)
        x = clips_to_reference_input(W.synthetic_clips(2, 16, seed=4))
        with torch.no_grad():
            logits = model(x)
            stats = []
            h = x
            for i, m in enumerate(model.base):
                h = m(h)
                stats.append([h.double().mean().item(), h.double().abs().mean().item(), h.double().pow(2).mean().sqrt().item()])
        np.savez_compressed(os.path.join(out_dir, f"s3d_{variant}.npz"), logits=logits.numpy(), layer_stats=np.array(stats),
                            feat_sample=h[0, :32, 0].numpy().astype(np.float32), seed_weights=0, seed_clips=4, b=2, t=16)
        print("s3d", variant, "logits", logits.flatten().tolist(), "last stats", stats[-1])
        # BASELINE configs[4] geometry: 64-frame clips (t1 = 32, t2 = 16, t3 = 8 frames at the head)
        x64 = clips_to_reference_input(W.synthetic_clips(1, 64, seed=5))
        with torch.no_grad():
            logits64 = model(x64)
            stats64 = []
            h = x64
            for m in model.base:
                h = m(h)
                stats64.append([h.double().mean().item(), h.double().abs().mean().item(), h.double().pow(2).mean().sqrt().item()])
        np.savez_compressed(os.path.join(out_dir, f"s3d_t64_{variant}.npz"), logits=logits64.numpy(), layer_stats=np.array(stats64),
                            seed_weights=0, seed_clips=5, b=1, t=64)
        print("s3d T=64", variant, "logits", logits64.flatten().tolist())


if __name__ == "__main__" and os.environ.get("FF_GOLDEN_S3D", "1") == "1":
    main_s3d()


def main_sample_clips():
    """tests/golden/sample_clip_crops.npz — BASELINE configs[0]: the reference's own front-end on its own sample clips.

    For every clip of sample__prediction_data: 15 frames by ``VideoReader.read_random_frames(path, 15, seed=0)``
    (helpers_read_video_1.py:50-69), ``FaceExtractor.process_video`` with the reference BlazeFace and its shipped weights
    (helpers_face_extract_1.py:120-317, blazeface.py), first 15 face crops, ``cv2.resize(INTER_AREA, 224)`` +
    ``cvtColor(RGB2BGR)`` exactly as ``face_blaze`` does (cvit_prediction.py:124-149).  Stored: the uint8 crops, per-clip
    offsets, one clip's RAW variable-size crops (for the resize kernel), and the reference CViT class's logits / scores for
    the "bn" and "decisive" synthetic state_dicts (chunks of <= 32 crops per clip, cvit_prediction.py:224-242)."""
    import cv2
    helpers = os.path.join(REF, "helpers")
    sys.path.insert(0, helpers)
    from blazeface import BlazeFace  # noqa: E402
    from helpers_face_extract_1 import FaceExtractor  # noqa: E402
    from helpers_read_video_1 import VideoReader  # noqa: E402
    net = BlazeFace()
    net.load_weights(os.path.join(helpers, "blazeface.pth"))
    net.load_anchors(os.path.join(helpers, "anchors.npy"))
    net.train(False)
    reader = VideoReader(verbose=False)
    extractor = FaceExtractor(lambda p: reader.read_random_frames(p, num_frames=15, seed=0), net)
    clip_dir = os.path.join(REF, "sample__prediction_data")
    names = sorted(f for f in os.listdir(clip_dir) if f.endswith(".mp4"))
    crops, offsets, raw, raw_name = [], [0], [], None
    for name in names:
        frames = extractor.process_video(os.path.join(clip_dir, name))
        got = []
        for fd in frames:
            for face in fd["faces"]:
                if len(got) < 15 and face.size > 0:
                    if raw_name in (None, name) and len(raw) < 15:
                        raw_name = name
                        raw.append(np.ascontiguousarray(face))
                    f224 = cv2.resize(face, (224, 224), interpolation=cv2.INTER_AREA)
                    got.append(cv2.cvtColor(f224, cv2.COLOR_RGB2BGR))
        crops += got
        offsets.append(offsets[-1] + len(got))
        print("sample clip", name, "frames", len(frames), "crops kept", len(got),
              "raw sizes", sorted({f.shape[:2] for fd in frames for f in fd["faces"]})[:4])
    crops = np.stack(crops)
    pred_sig, pre_process_prediction = reference_reduction_functions()
    out = {"crops": crops, "offsets": np.array(offsets, np.int32), "clip_names": np.array(names),
           "raw_clip": np.array(raw_name), "raw_hw": np.array([r.shape[:2] for r in raw], np.int32),
           "raw_bytes": np.concatenate([r.reshape(-1) for r in raw])}
    x = O.normalize_crops(torch.from_numpy(crops))
    for variant in ("bn", "decisive"):
        sd = W.make_state_dict(0, variant)
        model = CViT(image_size=224, patch_size=7, num_classes=2, channels=512, dim=1024, depth=6, heads=8, mlp_dim=2048).eval()
        model.load_state_dict(sd, strict=True)
        logits, scores = [], []
        with torch.no_grad():
            for v in range(len(names)):
                a, b = offsets[v], offsets[v + 1]
                if b == a:
                    scores.append(0.5)
                    continue
                lg = torch.cat([model(x[c:min(b, c + 32)]) for c in range(a, b, 32)])
                logits.append(lg)
                scores.append(float(pre_process_prediction(pred_sig(lg))))
        out[f"logits_{variant}"] = torch.cat(logits).numpy()
        out[f"scores_{variant}"] = np.array(scores, np.float32)
        print("sample clips", variant, "scores", [round(s, 4) for s in scores])
    path = os.path.join(ROOT, "tests", "golden", "sample_clip_crops.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) / 1e6, "MB")


if __name__ == "__main__" and os.environ.get("FF_GOLDEN_SAMPLE_CLIPS", "1") == "1":
    main_sample_clips()


def main_s3d_srm():
    """tests/golden/s3d_srm.npz — the SRM front-end (`S3D(1, 'yes')`, model.py:11-16,38-39): logits of the reference class
    for (a) the synthetic state_dict with synthetic zero-sum residual filters and (b) the same state_dict with the
    reference's OWN 30 SRM filters (the value `HPF()` builds from SRM/srm_filter_kernel.py, stored here as data)."""
    from oracle import s3d_oracle as S  # noqa: E402
    S3D = load_reference_s3d_class()
    out_dir = os.path.join(ROOT, "tests", "golden")
    model = S3D(1, "yes").eval()
    real_bank = model.SRM.hpf.weight.detach().clone()
    out = {"hpf_weight_reference": real_bank.numpy().astype(np.float32), "seed_weights": 0, "seed_clips": 8, "b": 1, "t": 16}
    x = clips_to_reference_input(W.synthetic_clips(1, 16, seed=8))
    for tag, bank in (("synthetic", None), ("reference_bank", real_bank)):
        sd = W.make_s3d_state_dict(0, "bn", srm=True)
        if bank is not None:
            sd["SRM.hpf.weight"] = bank
        model.load_state_dict(sd, strict=True)
        with torch.no_grad():
            logits = model(x)
            h0 = model.SRM(x)
        assert (S.forward(x, sd, srm=True) - logits).abs().max().item() <= 1e-4
        out[f"logits_{tag}"] = logits.numpy()
        out[f"hpf_rms_{tag}"] = np.float32(h0.double().pow(2).mean().sqrt().item())
        print("s3d srm", tag, "logits", logits.flatten().tolist(), "hpf rms", float(out[f"hpf_rms_{tag}"]))
    np.savez_compressed(os.path.join(out_dir, "s3d_srm.npz"), **out)


if __name__ == "__main__" and os.environ.get("FF_GOLDEN_S3D_SRM", "1") == "1":
    main_s3d_srm()


def main_face_extract():
    """tests/golden/face_extract.npz — the reference FaceExtractor (helpers/helpers_face_extract_1.py) with the reference
    BlazeFace and its shipped weights on frames of two small sample clips: a landscape one (3 tiles per frame) and a
    portrait one (1 tile).  Stored: the frames, the 128x128 tiles of `_tile_frames`, per frame the detections after
    `_resize_detections` / `_untile_detections` / `nms` and the crop rectangles after `_add_margin_to_detections`."""
    helpers = os.path.join(REF, "helpers")
    sys.path.insert(0, helpers)
    from blazeface import BlazeFace  # noqa: E402
    from helpers_face_extract_1 import FaceExtractor  # noqa: E402
    from helpers_read_video_1 import VideoReader  # noqa: E402
    net = BlazeFace()
    net.load_weights(os.path.join(helpers, "blazeface.pth"))
    net.load_anchors(os.path.join(helpers, "anchors.npy"))
    net.train(False)
    reader = VideoReader(verbose=False)
    out = {}
    for tag, name, idxs in (("land", "0017_fake.mp4.mp4", [0, 30, 60, 90]), ("port", "0048_fake.mp4.mp4", [0, 200, 400, 800])):
        path = os.path.join(REF, "sample__prediction_data", name)
        fx = FaceExtractor(lambda p: reader.read_frames_at_indices(p, idxs), net)
        frames, got_idxs = reader.read_frames_at_indices(path, idxs)
        tiles, resize_info = fx._tile_frames(frames, net.input_size)
        det = net.predict_on_batch(tiles, apply_nms=False)
        det = fx._resize_detections(det, net.input_size, resize_info)
        det = fx._untile_detections(frames.shape[0], (frames.shape[2], frames.shape[1]), det)
        det = net.nms(det)
        F = frames.shape[0]
        faces = np.zeros((F, 16, 17), np.float32)
        rects = np.zeros((F, 16, 4), np.int32)
        counts = np.zeros((F,), np.int32)
        for i in range(F):
            k = len(det[i])
            counts[i] = k
            faces[i, :k] = det[i].numpy()
            m = fx._add_margin_to_detections(det[i], (frames.shape[2], frames.shape[1]), 0.2)
            rects[i, :k] = m[:, :4].cpu().numpy().astype(int)
        res = fx.process_video(path)
        shapes = [[f.shape[:2] for f in r["faces"]] for r in res]
        out.update({f"{tag}_frames": frames, f"{tag}_tiles": tiles, f"{tag}_faces": faces, f"{tag}_rects": rects, f"{tag}_counts": counts,
                    f"{tag}_idxs": np.array(got_idxs)})
        print("face_extract", tag, frames.shape, "tiles", tiles.shape, "faces per frame", counts.tolist(), "crop shapes", shapes)
        for i, r in enumerate(res):      # the public result agrees with the internal steps replayed above
            assert [tuple(f.shape[:2]) for f in r["faces"]] == [(int(y1 - y0), int(x1 - x0)) for y0, x0, y1, x1 in rects[i, :counts[i]]]
    path = os.path.join(ROOT, "tests", "golden", "face_extract.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) / 1e6, "MB")


if __name__ == "__main__" and os.environ.get("FF_GOLDEN_FACE_EXTRACT", "1") == "1":
    main_face_extract()
