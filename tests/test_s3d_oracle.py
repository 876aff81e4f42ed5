"""CPU: the S3D oracle against golden vectors produced by the reference class."""
import importlib.util
import os

import numpy as np
import pytest
import torch

from fac_fake_b200 import weights as W
from oracle import s3d_oracle as S

REF_DIR = "/root/reference/sx_exp_deepfakedetect-master/S3D"


def _ref_input(clips):
    return clips.permute(0, 4, 1, 2, 3).contiguous().float()


@pytest.mark.parametrize("variant", ["default", "bn"])
def test_s3d_oracle_matches_reference_golden(golden_dir, variant):
    g = np.load(os.path.join(golden_dir, f"s3d_{variant}.npz"))
    sd = W.make_s3d_state_dict(int(g["seed_weights"]), variant)
    x = _ref_input(W.synthetic_clips(int(g["b"]), int(g["t"]), seed=int(g["seed_clips"])))
    torch.set_num_threads(os.cpu_count() or 4)
    taps = {}
    got = S.forward(x, sd, taps).numpy()
    np.testing.assert_allclose(got, g["logits"], rtol=0, atol=2e-5 * max(1.0, np.abs(g["logits"]).max()))
    for i in range(16):
        h = taps[i].double()
        np.testing.assert_allclose([h.mean().item(), h.abs().mean().item(), h.pow(2).mean().sqrt().item()], g["layer_stats"][i], rtol=1e-4)
    np.testing.assert_allclose(taps[15][0, :32, 0].numpy(), g["feat_sample"], rtol=1e-3, atol=1e-5)


def test_shapes_and_time_reduction():
    """T = 32 -> 16 after the stem, 8 after Mixed_3's pool, 4 after Mixed_4's pool; avg-pool window 2 -> 3 scores, mean."""
    sd = W.make_s3d_state_dict(0, "default")
    x = torch.zeros(1, 3, 32, 64, 64)
    taps = {}
    y = S.forward(x, sd, taps)
    assert y.shape == (1, 1)
    assert taps[0].shape == (1, 64, 16, 32, 32) and taps[7].shape[2] == 8 and taps[13].shape == (1, 832, 4, 2, 2)
    assert abs(S.video_score(torch.tensor([0.0, 0.0])) - 0.5) < 1e-12


@pytest.mark.skipif(not os.path.isdir(REF_DIR), reason="reference not mounted (GPU box)")
def test_s3d_oracle_matches_live_reference_class():
    spec = importlib.util.spec_from_file_location("mk_golden", os.path.join(os.path.dirname(__file__), "..", "oracle", "make_golden.py"))
    for k in ("FF_GOLDEN_GGCA", "FF_GOLDEN_RESVITKAN", "FF_GOLDEN_BLAZEFACE", "FF_GOLDEN_S3D"):
        os.environ[k] = "0"
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    sd = W.make_s3d_state_dict(3, "bn")
    model = mg.load_reference_s3d_class()(1, "no").eval()
    assert set(sd.keys()) == set(model.state_dict().keys())
    model.load_state_dict(sd, strict=True)
    x = _ref_input(W.synthetic_clips(1, 16, seed=6, hw=112))
    with torch.no_grad():
        ref = model(x)
    assert torch.allclose(S.forward(x, sd), ref, atol=1e-5)
