"""CPU: the ResVitKan oracle against golden vectors produced by the reference class."""
import os
import sys

import numpy as np
import pytest
import torch

from fac_fake_b200 import weights as W
from oracle import cvit_oracle as O
from oracle import resvitkan_oracle as R

REF_DIR = "/root/reference/CViT-main/ResVitKan"


@pytest.mark.parametrize("variant", ["default", "bn"])
def test_resvitkan_oracle_matches_reference_golden(golden_dir, variant):
    g = np.load(os.path.join(golden_dir, f"resvitkan_{variant}.npz"))
    sd = W.make_resvitkan_state_dict(int(g["seed_weights"]), variant)
    x = O.normalize_crops(W.synthetic_crops(int(g["n"]), seed=int(g["seed_crops"])))
    torch.set_num_threads(os.cpu_count() or 4)
    got = R.forward(x, sd).numpy()
    scale = max(1.0, np.abs(g["logits"]).max())
    np.testing.assert_allclose(got, g["logits"], rtol=0, atol=2e-5 * scale)
    taps = {}
    with torch.no_grad():
        R.features(x[:2], sd, taps=taps)
    for i, name in enumerate(("stem", "layer1", "layer2", "layer3", "layer4", "channel")):
        h = taps[name].double()
        np.testing.assert_allclose([h.mean().item(), h.abs().mean().item(), h.pow(2).mean().sqrt().item()],
                                   g["layer_stats"][i], rtol=1e-4)
    np.testing.assert_allclose(taps["stem"][0, :, :6, :6].numpy(), g["stem_sample"], rtol=1e-4, atol=1e-6)


def test_bspline_partition_of_unity_and_support():
    grid = (torch.arange(-3, 9) * 0.4 - 1.0).expand(5, -1).contiguous()
    x = torch.linspace(-0.99, 0.99, 40).unsqueeze(1).expand(-1, 5).contiguous()
    b = R.b_splines(x, grid)
    assert b.shape == (40, 5, 8)
    assert torch.allclose(b.sum(-1), torch.ones(40, 5), atol=1e-5)       # inside the grid range the cubic bases sum to 1
    assert (b >= -1e-6).all()
    far = R.b_splines(torch.full((1, 5), 10.0), grid)
    assert far.abs().max() == 0                                           # outside every half-open interval


@pytest.mark.skipif(not os.path.isdir(REF_DIR), reason="reference not mounted (GPU box)")
def test_resvitkan_oracle_matches_live_reference_class():
    sys.path.insert(0, REF_DIR)
    try:
        from ResVitKan import CViT as RVK
    finally:
        sys.path.pop(0)
    sd = W.make_resvitkan_state_dict(2, "bn")
    m = RVK().eval()
    m.load_state_dict(sd, strict=True)
    x = O.normalize_crops(W.synthetic_crops(3, seed=5))
    with torch.no_grad():
        ref = m(x)
    assert torch.allclose(R.forward(x, sd), ref, atol=1e-5)
    assert set(sd.keys()) == set(m.state_dict().keys())
