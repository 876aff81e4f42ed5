"""GPU (B200): the S3D clip classifier through the C-ABI against its oracle and the golden vectors
produced by the reference class (/root/reference/sx_exp_deepfakedetect-master/S3D/model.py)."""
import os

import numpy as np
import pytest
import torch

from fac_fake_b200 import weights as W
from oracle import s3d_oracle as S

pytestmark = pytest.mark.gpu

T = 16


def _engine(variant):
    from fac_fake_b200 import S3DEngine
    sd = W.make_s3d_state_dict(0, variant)
    return S3DEngine(1, "no", frames_per_clip=T, max_clips=2).to("cuda:0").load_state_dict(sd), sd


@pytest.fixture(scope="module")
def s3d_bn():
    return _engine("bn")


@pytest.fixture(scope="module")
def s3d_default():
    return _engine("default")


def _ref_input(clips):
    return clips.permute(0, 4, 1, 2, 3).contiguous().float()


@pytest.mark.parametrize("variant", ["default", "bn"])
def test_every_base_module_matches_oracle(variant, s3d_bn, s3d_default):
    """Activation after each of the 16 `base` modules (stem, pools, 9 Inception blocks with in-place concat)."""
    eng, sd = s3d_bn if variant == "bn" else s3d_default
    torch.set_num_threads(os.cpu_count() or 4)
    clips = W.synthetic_clips(1, T, seed=51)
    taps = {}
    S.forward(_ref_input(clips), sd, taps)
    xg = clips.cuda()
    for idx in range(16):
        ref = taps[idx].permute(0, 2, 3, 4, 1).contiguous().flatten()        # NCDHW -> NDHWC
        got = eng.debug_activation(xg, idx)
        assert got.numel() == ref.numel(), idx
        assert torch.isfinite(got).all(), idx
        rel_rms = ((got - ref).pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()).item()
        err = (got - ref).abs().max().item()
        # bf16 activations through up to 40 conv layers: 2^-9 per stored value, compounding
        assert rel_rms <= 0.004 * (idx + 2) and err <= 0.06 * ref.abs().max().item(), f"base.{idx}: rms {rel_rms} max err {err}"


@pytest.mark.parametrize("variant", ["default", "bn"])
def test_logits_match_reference_golden(golden_dir, variant, s3d_bn, s3d_default):
    eng, _ = s3d_bn if variant == "bn" else s3d_default
    g = np.load(os.path.join(golden_dir, f"s3d_{variant}.npz"))
    clips = W.synthetic_clips(int(g["b"]), int(g["t"]), seed=int(g["seed_clips"]))
    got = eng(clips.cuda()).cpu().numpy()                                    # uint8 frames as decoded
    tol = 3e-2 * max(1.0, np.abs(g["logits"]).max())
    assert np.isfinite(got).all() and np.abs(got - g["logits"]).max() <= tol
    got2 = eng(_ref_input(clips).cuda()).cpu().numpy()                       # the module's own fp32 NCDHW input
    assert np.abs(got2 - got).max() <= 1e-6                                  # 0..255 integers are exact in bf16
    # passes of max_clips = 2: three clips -> two passes, same per-clip results
    three = torch.cat([clips, clips[:1]])
    got3 = eng(three.cuda()).cpu().numpy()
    assert np.array_equal(got3[:2], got) and np.array_equal(got3[2], got[0])


def test_streamed_host_input_matches_device_input(s3d_bn):
    """Pinned host clips take the chunked copy/compute-overlap path; results must equal the resident-input path."""
    eng, sd = s3d_bn
    clips = W.synthetic_clips(3, T, seed=53)
    a = eng(clips.cuda()).cpu()
    b = eng(clips.pin_memory()).cpu()
    assert torch.equal(a, b)


def test_video_score_and_errors(s3d_bn):
    eng, sd = s3d_bn
    clips = W.synthetic_clips(2, T, seed=52)
    ref = S.video_score(S.forward(_ref_input(clips), sd))
    assert abs(eng.video_score(clips.cuda()) - ref) <= 1e-2
    with pytest.raises(ValueError):
        eng(torch.zeros((1, 3, T + 1, 224, 224)))
    from fac_fake_b200 import S3DEngine
    with pytest.raises(ValueError):
        S3DEngine(1, "maybe")
    with pytest.raises(ValueError):
        S3DEngine(1, "no", frames_per_clip=8).to("cuda:0").load_state_dict(sd)    # head would see < 2 frames


# ---- BASELINE configs[4] geometry: 64-frame clips (the benchmarked shape: t1 = 32, t2 = 16, t3 = 8 frames at the head,
#      other temporal TMA extents than T = 16)
@pytest.mark.parametrize("variant", ["default", "bn"])
def test_t64_every_base_module_and_logits(golden_dir, variant):
    from fac_fake_b200 import S3DEngine
    T64 = 64
    sd = W.make_s3d_state_dict(0, variant)
    eng = S3DEngine(1, "no", frames_per_clip=T64, max_clips=2).to("cuda:0").load_state_dict(sd)
    torch.set_num_threads(os.cpu_count() or 4)
    g = np.load(os.path.join(golden_dir, f"s3d_t64_{variant}.npz"))
    clips = W.synthetic_clips(int(g["b"]), T64, seed=int(g["seed_clips"]))
    taps = {}
    ref_logits = S.forward(_ref_input(clips), sd, taps)
    assert np.abs(ref_logits.numpy() - g["logits"]).max() <= 1e-4            # oracle == reference class at T = 64
    xg = clips.cuda()
    for idx in range(16):
        ref = taps[idx].permute(0, 2, 3, 4, 1).contiguous().flatten()
        got = eng.debug_activation(xg, idx)
        assert got.numel() == ref.numel(), idx
        assert torch.isfinite(got).all(), idx
        rel_rms = ((got - ref).pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()).item()
        assert rel_rms <= 0.004 * (idx + 2), f"base.{idx}: rms {rel_rms}"
    got = eng(xg).cpu().numpy()
    tol = 3e-2 * max(1.0, np.abs(g["logits"]).max())
    assert np.isfinite(got).all() and np.abs(got - g["logits"]).max() <= tol
    # two clips in one pass and the same clips one by one give the same bits
    two = torch.cat([clips, W.synthetic_clips(1, T64, seed=6)]).cuda()
    both = eng(two).cpu().numpy()
    assert np.array_equal(both[0], got[0])


# ---- SRM front-end: S3D(num_class, 'yes') — 30 high-pass residual filters in front of `base` (model.py:38-39, SRM/HPF.py)
@pytest.mark.parametrize("bank", ["synthetic", "reference_bank"])
def test_srm_front_end(golden_dir, bank):
    from fac_fake_b200 import S3DEngine
    g = np.load(os.path.join(golden_dir, "s3d_srm.npz"))
    sd = W.make_s3d_state_dict(0, "bn", srm=True)
    if bank == "reference_bank":                       # the reference's own 30 SRM filters (data fixture)
        sd["SRM.hpf.weight"] = torch.from_numpy(g["hpf_weight_reference"])
    eng = S3DEngine(1, "yes", frames_per_clip=int(g["t"]), max_clips=2).to("cuda:0").load_state_dict(sd)
    torch.set_num_threads(os.cpu_count() or 4)
    clips = W.synthetic_clips(int(g["b"]), int(g["t"]), seed=int(g["seed_clips"]))
    taps = {}
    ref_logits = S.forward(_ref_input(clips), sd, taps, srm=True)
    assert np.abs(ref_logits.numpy() - g[f"logits_{bank}"]).max() <= 1e-4            # oracle == reference class
    xg = clips.cuda()
    for idx in range(16):
        ref = taps[idx].permute(0, 2, 3, 4, 1).contiguous().flatten()
        got = eng.debug_activation(xg, idx)
        assert got.numel() == ref.numel() and torch.isfinite(got).all(), idx
        rel_rms = ((got - ref).pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()).item()
        assert rel_rms <= 0.004 * (idx + 3), f"base.{idx}: rms {rel_rms}"
    got = eng(xg).cpu().numpy()
    tol = 3e-2 * max(1.0, np.abs(g[f"logits_{bank}"]).max())
    assert np.isfinite(got).all() and np.abs(got - g[f"logits_{bank}"]).max() <= tol
    got2 = eng(_ref_input(clips).cuda()).cpu().numpy()                                # fp32 NCDHW entry
    assert np.abs(got2 - got).max() <= 1e-6
    # a no-SRM checkpoint must not load into an SRM engine (first conv has 30 input channels)
    with pytest.raises(ValueError):
        S3DEngine(1, "yes", frames_per_clip=16).to("cuda:0").load_state_dict(W.make_s3d_state_dict(0, "bn"))
