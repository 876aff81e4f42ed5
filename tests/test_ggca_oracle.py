"""CPU: the oracle of the cvit_GGCA_ADD_DEConv_RepBn8 variant against golden vectors produced by
the reference class (oracle/make_golden.py:main_ggca) and, when /root/reference is mounted, the live class."""
import importlib.util
import os

import numpy as np
import pytest
import torch

from fac_fake_b200 import weights as W
from oracle import cvit_oracle as O
from oracle import ggca_oracle as G

REF_FILE = "/root/reference/CViT-main/model/cvit_GGCA_ADD_DEConv_RepBn8.py"


@pytest.mark.parametrize("variant", ["default", "bn"])
def test_ggca_oracle_matches_reference_golden(golden_dir, variant):
    g = np.load(os.path.join(golden_dir, f"ggca_{variant}.npz"))
    sd = W.make_ggca_state_dict(int(g["seed_weights"]), variant)
    x = O.normalize_crops(W.synthetic_crops(int(g["n"]), seed=int(g["seed_crops"])))
    torch.set_num_threads(os.cpu_count() or 4)
    got = G.forward(x, sd).numpy()
    np.testing.assert_allclose(got, g["logits"], rtol=0, atol=2e-5 * max(1.0, np.abs(g["logits"]).max()))
    with torch.no_grad():
        f = G.features(x[:2], sd)
        gated = f * G.ggca(f, sd)
    for t, ref in zip((f, gated), g["feat_stats"]):
        d = t.double()
        np.testing.assert_allclose([d.mean().item(), d.abs().mean().item(), d.pow(2).mean().sqrt().item()], ref, rtol=1e-4)
    np.testing.assert_allclose(gated[0, :16].numpy(), g["gated_sample"], rtol=1e-3, atol=1e-5)


def test_deconv_fold_is_the_sum_of_its_five_branches():
    """Folding happens on weights; check it against running the five difference convolutions separately."""
    gen = torch.Generator().manual_seed(3)
    sd = {}
    W._deconv(gen, sd, "d", 8, 0.3)
    w, b = G.deconv_weight(sd, "d")
    x = torch.randn(2, 8, 9, 9, generator=gen)
    import torch.nn.functional as F
    w1 = sd["d.conv1_1.conv.weight"]
    cd = F.conv2d(x, w1, padding=1) - F.conv2d(x, w1.sum((2, 3), keepdim=True))          # central difference
    w2 = sd["d.conv1_2.conv.weight"]                                                       # [o,i,3]: column taps
    hd = F.conv2d(x, w2.unsqueeze(-1), padding=(1, 0))
    hd = F.pad(hd, (1, 0))[..., :9] - F.pad(hd, (0, 1))[..., 1:]                          # left minus right neighbour
    w3 = sd["d.conv1_3.conv.weight"]
    vd = F.conv2d(x, w3.unsqueeze(-2), padding=(0, 1))
    vd = F.pad(vd, (0, 0, 1, 0))[..., :9, :] - F.pad(vd, (0, 0, 0, 1))[..., 1:, :]        # upper minus lower neighbour
    w4 = sd["d.conv1_4.conv.weight"]
    ad = F.conv2d(x, w4, padding=1) - F.conv2d(x, w4.reshape(8, 8, 9)[:, :, G.AD_PERM].reshape(8, 8, 3, 3), padding=1)
    plain = F.conv2d(x, sd["d.conv1_5.weight"], padding=1)
    ref = cd + hd + vd + ad + plain + b.view(1, -1, 1, 1)
    got = F.conv2d(x, w, b, padding=1)
    assert torch.allclose(got, ref, atol=2e-5)


def test_state_dict_keys_and_linear_norm_eval_semantics():
    sd = W.make_ggca_state_dict(1, "bn")
    assert sd["features1.26.weight"].shape == (128, 128, 3, 3) and "features1.28.weight" not in sd     # BN-less conv pair
    assert sd["transformer.layers.0.1.fn.norm.norm1.weight"].shape == (1024,)
    # eval(): LinearNorm == norm1 == LayerNorm(eps 1e-6); norm2 (RepBN) must not influence the output
    x = O.normalize_crops(W.synthetic_crops(2, seed=9))
    a = G.forward(x, sd)
    sd2 = dict(sd)
    sd2["transformer.layers.0.1.fn.norm.norm2.alpha"] = torch.full((1,), 7.0)
    assert torch.equal(a, G.forward(x, sd2))


@pytest.mark.skipif(not os.path.isfile(REF_FILE), reason="reference not mounted (GPU box)")
def test_ggca_oracle_matches_live_reference_class():
    spec = importlib.util.spec_from_file_location("mk_golden", os.path.join(os.path.dirname(__file__), "..", "oracle", "make_golden.py"))
    os.environ["FF_GOLDEN_GGCA"] = "0"
    os.environ["FF_GOLDEN_RESVITKAN"] = "0"
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    sd = W.make_ggca_state_dict(2, "bn")
    model = mg.make_ggca_model(sd)
    assert set(sd.keys()) == set(model.state_dict().keys())
    x = O.normalize_crops(W.synthetic_crops(3, seed=5))
    with torch.no_grad():
        ref = model(x)
    assert torch.allclose(G.forward(x, sd), ref, atol=1e-5)
