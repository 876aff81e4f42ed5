"""GPU (B200): the BlazeFace detector through the C-ABI against the reference's own outputs
(reference class + shipped weights on tiles of its sample videos, tests/golden/blazeface_*.npz) and the oracle."""
import os

import numpy as np
import pytest
import torch

from oracle import blazeface_oracle as B

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def blaze(golden_dir):
    from fac_fake_b200 import BlazeFaceEngine
    w = np.load(os.path.join(golden_dir, "blazeface_weights.npz"))
    sd = {k: torch.from_numpy(w[k]) for k in w.files if k != "anchors"}
    eng = BlazeFaceEngine(max_tiles=8).to("cuda:0")
    eng.load_weights(sd)
    eng.load_anchors(w["anchors"])
    return eng, sd, torch.from_numpy(w["anchors"]), np.load(os.path.join(golden_dir, "blazeface_golden.npz"))


def test_raw_network_outputs_match_reference(blaze):
    eng, sd, anchors, g = blaze
    det, rb, rs = eng.predict_dense(g["tiles"], return_raw=True)       # 12 tiles > max_tiles = 8: two passes
    # fp32 on both sides, different summation order only (raw box regressors are O(100))
    np.testing.assert_allclose(rs.cpu().numpy(), g["raw_scores"], rtol=0, atol=2e-3)
    np.testing.assert_allclose(rb.cpu().numpy(), g["raw_boxes"], rtol=0, atol=5e-3)
    dense = B.dense_detections(torch.from_numpy(g["raw_boxes"]), torch.from_numpy(g["raw_scores"]).unsqueeze(-1), anchors)
    np.testing.assert_allclose(det.cpu().numpy(), dense.numpy(), rtol=0, atol=1e-4)


def test_every_block_against_oracle_on_random_tiles(blaze):
    """Random uint8 tiles exercise all borders (asymmetric TFLite padding of the stem and of the stride-2 blocks)."""
    eng, sd, anchors, g = blaze
    gen = torch.Generator().manual_seed(11)
    tiles = torch.randint(0, 256, (5, 128, 128, 3), generator=gen, dtype=torch.uint8)
    with torch.no_grad():
        r, c = B.forward(B.preprocess(tiles.permute(0, 3, 1, 2)), sd)
    det, rb, rs = eng.predict_dense(tiles.numpy(), return_raw=True)
    assert (rs.cpu() - c[..., 0]).abs().max().item() <= 2e-3 * max(1.0, c.abs().max().item())
    assert (rb.cpu() - r).abs().max().item() <= 2e-3 * max(1.0, r.abs().max().item())
    # the reference also accepts (b,3,H,W) uint8 tensors
    det2 = eng.predict_dense(tiles.permute(0, 3, 1, 2))
    assert torch.equal(det2, det)


def test_detections_and_faces_match_reference(blaze):
    eng, sd, anchors, g = blaze
    det = eng.predict_on_batch(g["tiles"], apply_nms=False)
    assert [len(d) for d in det] == g["det_counts"].tolist()
    faces = eng.nms(det)
    assert [len(f) for f in faces] == g["face_counts"].tolist()
    np.testing.assert_allclose(torch.cat(faces).numpy(), g["faces"], rtol=0, atol=2e-4)
    faces2 = eng.predict_on_batch(g["tiles"])                          # apply_nms=True path: mask + NMS on the device
    assert [len(f) for f in faces2] == g["face_counts"].tolist()
    assert all(torch.allclose(a, b, atol=1e-6) for a, b in zip(faces, faces2))
    one = eng.predict_on_image(g["tiles"][0])
    assert torch.allclose(one, faces[0], atol=1e-6)
    assert eng.predict_on_batch(np.zeros((1, 128, 128, 3), np.uint8))[0].shape[1] == 17


def test_device_nms_matches_oracle_on_crowded_tiles(blaze):
    """ff_blazeface_nms on fabricated dense detections: 0 / few / many overlapping candidates, chains of merges, and a
    tile with more than 64 candidates (host fallback)."""
    eng, sd, anchors, g = blaze
    eng.predict_dense(g["tiles"][:1])                                  # make sure the handle exists
    gen = torch.Generator().manual_seed(3)
    n = 9
    dense = torch.zeros((n, 896, 17))
    dense[..., 16] = torch.rand((n, 896), generator=gen) * 0.7         # below the 0.75 threshold
    for t, k in enumerate((0, 1, 2, 5, 12, 30, 50, 64, 90)):
        idx = torch.randperm(896, generator=gen)[:k]
        c = torch.rand((k, 2), generator=gen) * 0.5 + 0.1
        sz = torch.rand((k, 2), generator=gen) * 0.25 + 0.05
        dense[t, idx, 0:2] = c
        dense[t, idx, 2:4] = c + sz
        dense[t, idx, 4:16] = torch.rand((k, 12), generator=gen)
        dense[t, idx, 16] = torch.rand((k,), generator=gen) * 0.24 + 0.755
    got = eng._nms_on_device(dense.cuda())
    for t in range(n):
        d = dense[t]
        ref = B.weighted_nms(d[d[:, 16] >= B.MIN_SCORE])
        assert len(got[t]) == len(ref), t
        if ref:
            assert torch.allclose(got[t], torch.stack(ref), atol=2e-6), t
    # `nms(list)` (blazeface.py:225-234) = the same kernel in list mode, on the masked detections in arbitrary order
    lists = []
    for t in range(n):
        d = dense[t][dense[t][:, 16] >= B.MIN_SCORE]
        lists.append(d[torch.randperm(d.shape[0], generator=gen)])
    got = eng.nms(lists)
    assert len(got) == n
    for t in range(n):
        ref = B.weighted_nms(lists[t])
        assert got[t].shape == (len(ref), 17), t
        if ref:
            assert torch.allclose(got[t], torch.stack(ref), atol=2e-6), t
    assert eng.nms([]) == [] and eng.nms([torch.zeros((0, 17))])[0].shape == (0, 17)


def test_errors_are_loud(blaze):
    from fac_fake_b200 import BlazeFaceEngine, EngineError
    eng = BlazeFaceEngine().to("cuda:0")
    with pytest.raises(EngineError):
        eng.predict_on_batch(np.zeros((1, 128, 128, 3), np.uint8))     # no weights
    with pytest.raises(ValueError):
        blaze[0].predict_dense(np.zeros((1, 128, 128, 3), np.float32))
    bad = BlazeFaceEngine().to("cuda:0")
    sd = dict(blaze[1])
    sd.pop("classifier_8.bias")
    bad.load_weights(sd)
    bad.load_anchors(blaze[2].numpy())
    with pytest.raises(EngineError):
        bad.predict_on_batch(np.zeros((1, 128, 128, 3), np.uint8))
