"""CPU: the BlazeFace oracle against outputs of the reference class run with the reference's own
shipped weights on tiles of its sample videos (tests/golden/blazeface_*.npz, oracle/make_golden.py:main_blazeface)."""
import os
import sys

import numpy as np
import pytest
import torch

from oracle import blazeface_oracle as B

REF_HELPERS = "/root/reference/CViT-main/helpers"


@pytest.fixture(scope="module")
def blaze(golden_dir):
    w = np.load(os.path.join(golden_dir, "blazeface_weights.npz"))
    sd = {k: torch.from_numpy(w[k]) for k in w.files if k != "anchors"}
    return sd, torch.from_numpy(w["anchors"]), np.load(os.path.join(golden_dir, "blazeface_golden.npz"))


def test_raw_outputs_match_reference(blaze):
    sd, anchors, g = blaze
    x = torch.from_numpy(g["tiles"]).permute(0, 3, 1, 2)
    with torch.no_grad():
        r, c = B.forward(B.preprocess(x), sd)
    np.testing.assert_allclose(c[..., 0].numpy(), g["raw_scores"], rtol=0, atol=2e-4)
    np.testing.assert_allclose(r.numpy(), g["raw_boxes"], rtol=0, atol=2e-4)


def test_detections_and_blending_nms_match_reference(blaze):
    sd, anchors, g = blaze
    det = B.predict_on_batch(g["tiles"], sd, anchors, apply_nms=False)
    assert [len(d) for d in det] == g["det_counts"].tolist()
    faces = B.nms(det)
    assert [len(f) for f in faces] == g["face_counts"].tolist()
    np.testing.assert_allclose(torch.cat(faces).numpy(), g["faces"], rtol=0, atol=1e-5)
    # real frames of the reference's sample clips: one face per tile, well inside the tile, confident
    f = torch.cat(faces)
    assert (f[:, 16] > 0.75).all() and (f[:, :4] > -0.2).all() and (f[:, :4] < 1.2).all()


def test_nms_edge_cases():
    assert B.weighted_nms(torch.zeros((0, 17))) == []
    assert B.nms([torch.zeros((0, 17))])[0].shape == (0, 17)
    a = torch.zeros(17); a[:4] = torch.tensor([0.1, 0.1, 0.5, 0.5]); a[16] = 0.9
    b = a.clone(); b[:4] += 0.02; b[16] = 0.8                       # overlaps a  -> blended
    c = a.clone(); c[:4] = torch.tensor([0.6, 0.6, 0.9, 0.9]); c[16] = 0.7   # disjoint -> kept
    out = B.weighted_nms(torch.stack([b, c, a]))
    assert len(out) == 2
    blended = out[0]
    assert abs(blended[16].item() - 0.85) < 1e-6                    # mean score of the overlapping pair (blazeface.py:353)
    assert abs(blended[0].item() - (0.1 * 0.9 + 0.12 * 0.8) / 1.7) < 1e-6


@pytest.mark.skipif(not os.path.isdir(REF_HELPERS), reason="reference not mounted (GPU box)")
def test_oracle_matches_live_reference_class(blaze):
    sd, anchors, g = blaze
    sys.path.insert(0, REF_HELPERS)
    try:
        from blazeface import BlazeFace
    finally:
        sys.path.pop(0)
    net = BlazeFace()
    net.load_weights(os.path.join(REF_HELPERS, "blazeface.pth"))
    net.load_anchors(os.path.join(REF_HELPERS, "anchors.npy"))
    gen = torch.Generator().manual_seed(5)
    tiles = torch.randint(0, 256, (3, 128, 128, 3), generator=gen, dtype=torch.uint8).numpy()
    tiles = np.concatenate([tiles, g["tiles"][:2]])
    ref = net.predict_on_batch(tiles, apply_nms=True)
    got = B.predict_on_batch(tiles, sd, anchors, apply_nms=True)
    assert [len(a) for a in ref] == [len(b) for b in got]
    for a, b in zip(ref, got):
        assert torch.allclose(a, b, atol=1e-5)
