"""The face front-end around the detector: tiling, untiling, NMS across tiles, margin and crop of the
reference's ``FaceExtractor`` (helpers/helpers_face_extract_1.py) against tests/golden/face_extract.npz — outputs of the
reference class with the reference BlazeFace and its shipped weights on frames of two sample clips (landscape: 3 tiles per
frame, portrait: 1 tile)."""
import os

import numpy as np
import pytest
import torch

from oracle import blazeface_oracle as B
from oracle import resize_oracle as R


@pytest.fixture(scope="module")
def fx(golden_dir):
    g = np.load(os.path.join(golden_dir, "face_extract.npz"))
    w = np.load(os.path.join(golden_dir, "blazeface_weights.npz"))
    sd = {k: torch.from_numpy(w[k]) for k in w.files if k != "anchors"}
    return g, sd, w["anchors"]


def _engines(sd, anchors):
    from fac_fake_b200 import BlazeFaceEngine, FaceExtractorEngine
    det = BlazeFaceEngine(max_tiles=8).to("cuda:0")
    det.load_weights(sd)
    det.load_anchors(anchors)
    return det


def test_host_fallback_restates_the_reference_steps(fx):
    """The host path used for frames the kernel declines (> 64 candidates / > 16 faces): `_resize_detections`,
    `_untile_detections`, blending NMS and the margin rectangle, fed with the ORACLE's dense detections of the golden
    tiles — no GPU involved."""
    from fac_fake_b200.face_extract import FaceExtractorEngine

    class _Det:                                    # the two attributes + the NMS the fallback uses
        min_score_thresh, min_suppression_threshold = 0.75, 0.3

        def _weighted_non_max_suppression(self, d):
            return B.weighted_nms(d)

    g, sd, anchors = fx
    ex = FaceExtractorEngine(None, _Det())
    for tag, T in (("land", 3), ("port", 1)):
        tiles = torch.from_numpy(g[f"{tag}_tiles"])
        with torch.no_grad():
            rb, rs = B.forward(B.preprocess(tiles.permute(0, 3, 1, 2)), sd)
            dense = B.dense_detections(rb, rs, torch.from_numpy(anchors))
        _, H, W, _ = g[f"{tag}_frames"].shape
        for f in range(g[f"{tag}_frames"].shape[0]):
            det, rect = ex._frame_on_host(dense[f * T:(f + 1) * T], H, W)
            k = int(g[f"{tag}_counts"][f])
            assert det.shape[0] == k
            assert np.abs(det.numpy() - g[f"{tag}_faces"][f, :k]).max() <= 0.05          # pixels / scores
            assert np.abs(rect - g[f"{tag}_rects"][f, :k]).max() <= 1


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["land", "port"])
def test_device_tiling_detections_and_rectangles(fx, tag):
    from fac_fake_b200 import FaceExtractorEngine
    g, sd, anchors = fx
    ex = FaceExtractorEngine(None, _engines(sd, anchors))
    frames = torch.from_numpy(g[f"{tag}_frames"]).cuda()
    faces, boxes, counts, tiles = ex.detect_frames_device(frames)
    np.testing.assert_array_equal(tiles.cpu().numpy(), g[f"{tag}_tiles"])                 # _tile_frames: byte work, bit-exact
    np.testing.assert_array_equal(counts.cpu().numpy(), g[f"{tag}_counts"])
    for f in range(frames.shape[0]):
        k = int(g[f"{tag}_counts"][f])
        assert np.abs(faces[f, :k].cpu().numpy() - g[f"{tag}_faces"][f, :k]).max() <= 0.05
        assert np.abs(boxes[f, :k].cpu().numpy() - g[f"{tag}_rects"][f, :k]).max() <= 1   # int truncation of a float within 1e-3 px


@pytest.mark.gpu
def test_process_video_and_device_crops_feed_the_classifier(fx):
    """`process_video` returns the reference's dictionaries; `extract_crops_device` hands 224x224 crops to the CViT engine
    without the pixels leaving the GPU, and they equal cv2.resize(INTER_AREA) + cvtColor of the same rectangles."""
    from fac_fake_b200 import CViTEngine, FaceExtractorEngine, weights as W
    g, sd, anchors = fx
    frames_np = g["land_frames"]
    idxs = g["land_idxs"].tolist()
    ex = FaceExtractorEngine(lambda path: (frames_np, idxs), _engines(sd, anchors))
    res = ex.process_video("/nowhere/clip.mp4")
    assert [r["frame_idx"] for r in res] == idxs and all(r["frame_w"] == 536 and r["frame_h"] == 500 for r in res)
    for f, r in enumerate(res):
        k = int(g["land_counts"][f])
        assert len(r["faces"]) == k == len(r["scores"])
        for j, face in enumerate(r["faces"]):
            y0, x0, y1, x1 = g["land_rects"][f, j]
            assert abs(face.shape[0] - (y1 - y0)) <= 1 and abs(face.shape[1] - (x1 - x0)) <= 1
            assert abs(r["scores"][j] - g["land_faces"][f, j, 16]) <= 1e-3
    model = CViTEngine(max_crops=32).to("cuda:0").load_state_dict(W.make_state_dict(0, "bn"))
    frames = torch.from_numpy(frames_np).cuda()
    crops = ex.extract_crops_device(frames, model)
    assert crops.shape == (int(g["land_counts"].sum()), 224, 224, 3) and crops.is_cuda
    rects = [r for f in range(frames.shape[0]) for r in ex._frames_to_lists(frames)[f][1]]
    for i, (y0, x0, y1, x1) in enumerate(rects):
        f = i                                                  # one face per frame in this clip
        np.testing.assert_array_equal(crops[i].cpu().numpy(), R.crop_to_model_input(frames_np[f][y0:y1, x0:x1]))
    scores = model.predict_videos(crops, [0, crops.shape[0]])
    assert torch.isfinite(scores).all()
