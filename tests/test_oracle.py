"""CPU: the oracle against the committed golden vectors (outputs of the reference itself, see
oracle/make_golden.py) and — when /root/reference is present — against the live reference class."""
import os
import sys

import numpy as np
import pytest
import torch

from fac_fake_b200 import weights as W
from oracle import cvit_oracle as O

REF_MODEL_DIR = "/root/reference/CViT-main/model"


@pytest.mark.parametrize("variant", ["default", "bn"])
def test_oracle_logits_match_reference_golden(golden_dir, variant):
    g = np.load(os.path.join(golden_dir, f"cvit_logits_{variant}.npz"))
    sd = W.make_state_dict(int(g["seed_weights"]), variant)
    crops = W.synthetic_crops(int(g["n"]), seed=int(g["seed_crops"]))
    x = O.normalize_crops(crops)
    torch.set_num_threads(os.cpu_count() or 4)
    got = torch.cat([O.forward(x[0:32], sd), O.forward(x[32:40], sd)]).numpy()
    np.testing.assert_allclose(got, g["logits"], rtol=0, atol=2e-5)
    # the chunked helper reproduces the reference's slot = index-in-chunk rule
    got2 = O.forward_chunked(x, sd).numpy()
    np.testing.assert_allclose(got2, g["logits"], rtol=0, atol=2e-5)


@pytest.mark.parametrize("variant", ["default", "bn"])
def test_oracle_features_match_reference_golden(golden_dir, variant):
    g = np.load(os.path.join(golden_dir, f"cvit_logits_{variant}.npz"))
    sd = W.make_state_dict(0, variant)
    x = O.normalize_crops(W.synthetic_crops(40, seed=1))[0:4]
    h = x
    for li in range(17):
        h = O.feature_layer(h, sd, li)
        st = g["layer_stats"][li]
        hd = h.double()
        np.testing.assert_allclose([hd.mean().item(), hd.abs().mean().item(), hd.pow(2).mean().sqrt().item()], st, rtol=1e-4)
        if li == 2:
            np.testing.assert_allclose(h[0, :, :8, :8].numpy(), g["feat_l2_sample"], rtol=1e-4, atol=1e-6)
    scale = np.abs(g["feat_final"]).max()
    np.testing.assert_allclose(h.numpy(), g["feat_final"], rtol=0, atol=1e-4 * scale)


def test_reduction_matches_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "video_reduction.npz"))
    logits = torch.from_numpy(g["logits"])
    a = 0
    for n, want in zip(g["lens"], g["scores"]):
        got = O.video_score(logits[a:a + n])
        assert got == pytest.approx(float(want), abs=1e-7), (n, got, want)
        a += int(n)


def test_reduction_sentinels():
    assert O.video_score(torch.zeros((0, 2))) == 0.5                    # no faces   (cvit_prediction.py:218-219)
    assert O.video_score(torch.tensor([[3.0, -1.0]])) == 0.5            # 1 frame: squeeze -> len 2 -> 0.5
    assert O.video_score(torch.tensor([[3.0, -1.0], [2.0, 0.0]])) == 0.5
    s = O.video_score(torch.tensor([[3.0, -1.0], [2.0, 0.0], [1.0, 0.5]]))
    assert s == pytest.approx(float(torch.sigmoid(torch.tensor([3.0, 2.0, 1.0])).mean()), abs=1e-6)
    s = O.video_score(torch.tensor([[-3.0, 1.0], [-2.0, 0.0], [-1.0, 0.5]]))
    assert s == pytest.approx(abs(1 - float(torch.sigmoid(torch.tensor([1.0, 0.0, 0.5])).mean())), abs=1e-6)
    assert O.real_or_fake(0.5) == "FAKE" and O.real_or_fake(0.4999) == "REAL"


def test_batch_over_32_raises_like_reference():
    sd = W.make_state_dict(0, "shape_only") if False else None
    with pytest.raises(RuntimeError):
        O.forward(torch.zeros((33, 3, 224, 224)), {})


def test_slot_dependence_and_flops():
    assert O.count_flops_per_crop() == O.FLOPS_PER_CROP == 13_291_528_192
    sd = W.make_state_dict(0, "default")
    x = O.normalize_crops(W.synthetic_crops(2, seed=3))
    a = O.forward_slots(x, sd, torch.tensor([0, 1]))
    b = O.forward_slots(x, sd, torch.tensor([1, 0]))
    assert (a - b).abs().max() > 1e-3          # logits depend on the batch slot
    c = O.forward_slots(x.flip(0), sd, torch.tensor([1, 0])).flip(0)
    assert torch.allclose(a, c, atol=1e-5)     # ...and on nothing else about the batch


def test_bf16_contract_within_tolerance():
    """The engine's precision contract (bf16 operands, fp32 accumulate) stays inside the 2e-2 gate."""
    for variant in ("default", "bn"):
        sd = W.make_state_dict(0, variant)
        x = O.normalize_crops(W.synthetic_crops(4, seed=5))
        ref = O.forward(x, sd)
        sim = O.forward(x, sd, bf16_sim=True)
        assert (ref - sim).abs().max().item() < 1e-2


@pytest.mark.skipif(not os.path.isdir(REF_MODEL_DIR), reason="reference not mounted (GPU box)")
def test_oracle_matches_live_reference_class():
    sys.path.insert(0, REF_MODEL_DIR)
    try:
        from cvit import CViT
    finally:
        sys.path.pop(0)
    sd = W.make_state_dict(3, "bn")
    m = CViT().eval()
    m.load_state_dict(sd, strict=True)
    x = O.normalize_crops(W.synthetic_crops(5, seed=9))
    with torch.no_grad():
        ref = m(x)
    assert torch.allclose(O.forward(x, sd), ref, atol=1e-5)
    # key set equals the reference's state_dict
    assert set(sd.keys()) == set(m.state_dict().keys())
