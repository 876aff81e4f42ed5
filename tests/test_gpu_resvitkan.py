"""GPU (B200): the ResVitKan CUDA path through the C-ABI against its oracle and the golden vectors
produced by the reference class (/root/reference/CViT-main/ResVitKan/ResVitKan.py:284-329)."""
import os

import numpy as np
import pytest
import torch

from fac_fake_b200 import weights as W
from oracle import cvit_oracle as O
from oracle import resvitkan_oracle as R

pytestmark = pytest.mark.gpu

BF16_TOL = 2e-2      # same gate as the CViT path: per-frame logits within 2e-2 abs of the fp32 reference


def _engine(variant, max_crops=64):
    from fac_fake_b200 import ResVitKanEngine
    sd = W.make_resvitkan_state_dict(0, variant)
    return ResVitKanEngine(max_crops=max_crops).to("cuda:0").load_state_dict(sd), sd


@pytest.fixture(scope="module")
def rvk_bn():
    return _engine("bn")


@pytest.fixture(scope="module")
def rvk_default():
    return _engine("default")


def test_resnet_stages_match_oracle(rvk_bn):
    """Stem+pool, layer1..4 and channel conv against the fp32 oracle (bf16 activations => relative gate)."""
    eng, sd = rvk_bn
    torch.set_num_threads(os.cpu_count() or 4)
    crops = W.synthetic_crops(3, seed=31)
    taps = {}
    with torch.no_grad():
        R.features(O.normalize_crops(crops), sd, taps=taps)
    xg = crops.cuda()
    for step, name in enumerate(("stem", "layer1", "layer2", "layer3", "layer4", "channel"), start=1):
        ref = taps[name].permute(0, 2, 3, 1).contiguous().flatten()
        got = eng.debug_activation(xg, step)
        assert got.numel() == ref.numel(), name
        assert torch.isfinite(got).all(), name
        scale = ref.abs().max().item()
        err = (got - ref).abs().max().item()
        # 2^-8 relative rounding per stored activation, 3..6 stores per bottleneck; never above 3 %
        assert err <= scale * (0.006 if step == 1 else 0.03), f"{name}: err {err} scale {scale}"


def test_resnet_stages_fp32_nchw_input(rvk_bn):
    """The reference-compatible input (normalised fp32 NCHW) takes the same path after the conversion kernel."""
    eng, sd = rvk_bn
    crops = W.synthetic_crops(2, seed=32)
    a = eng.debug_activation(crops.cuda(), 2)
    b = eng.debug_activation(O.normalize_crops(crops).cuda(), 2)
    assert (a - b).abs().max().item() <= 0.02 * a.abs().max().item()


@pytest.mark.parametrize("variant", ["default", "bn"])
def test_logits_match_reference_golden(golden_dir, variant, rvk_bn, rvk_default):
    eng, _ = rvk_bn if variant == "bn" else rvk_default
    g = np.load(os.path.join(golden_dir, f"resvitkan_{variant}.npz"))
    n = int(g["n"])
    crops = W.synthetic_crops(n, seed=int(g["seed_crops"]))
    got = eng.forward_slots(crops.cuda(), torch.arange(n)).cpu().numpy()
    assert np.isfinite(got).all()
    assert np.abs(got - g["logits"]).max() <= BF16_TOL * max(1.0, np.abs(g["logits"]).max())
    # reference-compatible call: fp32 NCHW normalised input, slot = batch index
    got2 = eng(O.normalize_crops(crops).cuda()).cpu().numpy()
    assert np.abs(got2 - g["logits"]).max() <= BF16_TOL * max(1.0, np.abs(g["logits"]).max())
    with pytest.raises(RuntimeError):
        eng(torch.zeros((33, 3, 224, 224), device="cuda"))


@pytest.mark.parametrize("variant", ["default", "bn"])
def test_fp32_path_logits(golden_dir, variant):
    """compute_dtype='fp32' (CUDA-core trunk + encoder + KAN head): the north-star fp32 gate, 1e-4, against the
    reference class's golden logits; uint8 and reference-style fp32 NCHW input."""
    from fac_fake_b200 import ResVitKanEngine
    sd = W.make_resvitkan_state_dict(0, variant)
    eng = ResVitKanEngine(max_crops=32, compute_dtype="fp32").to("cuda:0").load_state_dict(sd)
    g = np.load(os.path.join(golden_dir, f"resvitkan_{variant}.npz"))
    n = min(int(g["n"]), 6)
    crops = W.synthetic_crops(int(g["n"]), seed=int(g["seed_crops"]))[:n]
    tol = 1e-4 * max(1.0, np.abs(g["logits"]).max())
    got = eng.forward_slots(crops.cuda(), torch.arange(n)).cpu().numpy()
    assert np.abs(got - g["logits"][:n]).max() <= tol
    got2 = eng(O.normalize_crops(crops).cuda()).cpu().numpy()
    assert np.abs(got2 - g["logits"][:n]).max() <= tol


def test_kan_head_alone_is_fp32_exact(rvk_bn):
    """Tokens after the last transformer layer -> oracle kan_head vs the engine's logits on the same pass."""
    eng, sd = rvk_bn
    crops = W.synthetic_crops(4, seed=33).cuda()
    t6 = eng.debug_activation(crops, 24).view(4, 2, 1024)
    logits = eng.debug_activation(crops, 25).view(4, 2)
    with torch.no_grad():
        ref = R.kan_head(t6[:, 0], sd)
    # the Linear runs on bf16 tensor cores (cls token and weight rounded), the KAN layers in fp32
    assert (logits - ref).abs().max().item() <= 5e-3 * max(1.0, ref.abs().max().item())


def test_batch_composition_and_ragged_sizes(rvk_bn):
    """A crop's logits do not depend on what else is in the pass (tiles hold 2..32 images), nor on pass splitting."""
    eng, sd = rvk_bn
    crops = W.synthetic_crops(70, seed=34).cuda()          # > max_crops=64: two passes
    slots = torch.arange(70) % 32
    full = eng.forward_slots(crops, slots).cpu()
    for lo, hi in ((0, 1), (5, 8), (33, 66)):
        part = eng.forward_slots(crops[lo:hi], slots[lo:hi]).cpu()
        assert torch.equal(part, full[lo:hi]), (lo, hi)


def test_predict_videos_decisions(rvk_bn):
    eng, sd = rvk_bn
    lens = [0, 2, 3, 9]
    offsets = np.concatenate([[0], np.cumsum(lens)]).tolist()
    crops = W.synthetic_crops(offsets[-1], seed=35)
    scores, logits = eng.predict_videos(crops.cuda(), offsets, return_logits=True)
    scores, logits = scores.cpu(), logits.cpu()
    x = O.normalize_crops(crops)
    for v, ln in enumerate(lens):
        if ln <= 2:
            assert scores[v].item() == 0.5
            continue
        sl = slice(offsets[v], offsets[v + 1])
        ref_logits = R.forward(x[sl], sd)
        assert (logits[sl] - ref_logits).abs().max().item() <= BF16_TOL * max(1.0, ref_logits.abs().max().item())
        ref_score = O.video_score(ref_logits)
        assert abs(scores[v].item() - ref_score) <= 1e-2
        if abs(ref_score - 0.5) > 2e-2:
            assert O.real_or_fake(scores[v].item()) == O.real_or_fake(ref_score)
