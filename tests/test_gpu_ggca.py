"""GPU (B200): the cvit_GGCA_ADD_DEConv_RepBn8 variant through the C-ABI against its oracle and the
golden vectors produced by the reference class (/root/reference/CViT-main/model/cvit_GGCA_ADD_DEConv_RepBn8.py)."""
import os

import numpy as np
import pytest
import torch

from fac_fake_b200 import weights as W
from oracle import cvit_oracle as O
from oracle import ggca_oracle as G

pytestmark = pytest.mark.gpu

# Per-frame logits vs the fp32 reference: the north-star gate, 2e-2.  This variant's difference convolutions amplify the
# rounding of their input activations (see test_conv_chain_and_gate_against_fp32_oracle); with bf16 activations the
# logits landed 2.6e-2 away, so its conv stack runs on fp16 activations and filters (kind::f16 MMAs, same tensor rate,
# three more mantissa bits).
BF16_TOL = 2e-2


def _engine(variant, max_crops=64):
    from fac_fake_b200 import CViTGGCAEngine
    sd = W.make_ggca_state_dict(0, variant)
    return CViTGGCAEngine(max_crops=max_crops).to("cuda:0").load_state_dict(sd), sd


@pytest.fixture(scope="module")
def ggca_bn():
    return _engine("bn")


@pytest.fixture(scope="module")
def ggca_default():
    return _engine("default")


# oracle plan entry -> engine debug step (entry 8 is the extra BN-less conv, tap 26)
STEP_OF_ENTRY = [1, 2, 3, 4, 5, 6, 7, 8, 26, 9, 10, 11, 12, 13, 14, 15, 16, 17]


def _q(t):
    """round to the conv stack's 16-bit operand type of this variant (fp16)"""
    return t.to(torch.float16).to(torch.float32)


def _nchw(flat, n, c):
    hw = int(round((flat.numel() // (n * c)) ** 0.5))
    return flat.view(n, hw, hw, c).permute(0, 3, 1, 2).contiguous()


@pytest.mark.parametrize("variant", ["default", "bn"])
def test_each_conv_layer_in_isolation(variant, ggca_bn, ggca_default):
    """Implementation check, one layer at a time: the oracle layer (fp32 accumulate, fp16-rounded folded kernel) applied
    to the ENGINE's previous activation must reproduce the engine's next activation to one fp16 rounding.  Covers the
    DEConv folding done in C++ from the five branch tensors, the BN-less conv pair and every pool."""
    eng, sd = ggca_bn if variant == "bn" else ggca_default
    torch.set_num_threads(os.cpu_count() or 4)
    n = 2
    crops = W.synthetic_crops(n, seed=41)
    xg = crops.cuda()
    prev = _q(O.normalize_crops(crops))
    sdq = dict(sd)
    with torch.no_grad():
        for entry, step in enumerate(STEP_OF_ENTRY):
            seq, ci, kind, bi, relu, pool = G.PLAN[entry]
            p = f"{seq}.{ci}"
            if kind == "de":                      # round the FOLDED kernel, as the engine does
                w, b = G.deconv_weight(sd, p)
                for k in list(sdq):
                    if k.startswith(p + ".conv1_"):
                        sdq[k] = torch.zeros_like(sd[k])
                sdq[p + ".conv1_5.weight"], sdq[p + ".conv1_5.bias"] = _q(w), b
            else:
                sdq[p + ".weight"] = _q(sd[p + ".weight"])
            ref = G.feature_layer(prev, sdq, entry)
            got = _nchw(eng.debug_activation(xg, step), n, ref.shape[1])
            assert got.shape == ref.shape, entry
            scale = ref.abs().max().item()
            err = (got - ref).abs().max().item()
            rel_rms = ((got - ref).pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()).item()
            assert err <= 0.002 * scale and rel_rms <= 0.001, f"plan entry {entry} (step {step}): err {err} scale {scale} rms {rel_rms}"
            prev = got


@pytest.mark.parametrize("variant", ["default", "bn"])
def test_conv_chain_and_gate_against_fp32_oracle(variant, ggca_bn, ggca_default):
    """End to end against the fp32 oracle.  The DEConv kernels are difference filters (centre tap = minus the sum of
    the others): they pass the rounding noise of their input at full gain while attenuating the signal, so the relative
    error of a 16-bit-activation pipeline grows at every DEConv — with bf16 activations a CPU simulation on these weights
    gave 1.9 % rms after entry 7, 4.5 % after entry 12 and 9 % at the feature map (and the engine matched that curve).
    With fp16 activations (8x finer) the gates below are that curve / 4; the implementation itself is pinned by the
    isolation test above."""
    eng, sd = ggca_bn if variant == "bn" else ggca_default
    torch.set_num_threads(os.cpu_count() or 4)
    crops = W.synthetic_crops(3, seed=41)
    xg = crops.cuda()
    h = O.normalize_crops(crops)
    with torch.no_grad():
        for entry, step in enumerate(STEP_OF_ENTRY):
            h = G.feature_layer(h, sd, entry)
            ref = h.permute(0, 2, 3, 1).contiguous().flatten()
            got = eng.debug_activation(xg, step)
            assert got.numel() == ref.numel() and torch.isfinite(got).all(), entry
            rel_rms = ((got - ref).pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()).item()
            gate = 0.008 if entry <= 10 else (0.018 if entry <= 14 else 0.04)
            print(f"ggca {variant} entry {entry}: rel rms {rel_rms:.5f}")
            assert rel_rms <= gate, f"plan entry {entry} (step {step}): relative rms error {rel_rms}"
        gated = (h * G.ggca(h, sd)).permute(0, 2, 3, 1).contiguous().flatten()
    got = eng.debug_activation(xg, 27)
    assert ((got - gated).pow(2).mean().sqrt() / gated.pow(2).mean().sqrt()).item() <= 0.08    # the gate squares x; bf16 store


def test_gate_alone_on_engine_features(ggca_bn):
    """GGCA kernel in isolation: oracle gate applied to the ENGINE's own (fp16) feature map; the result is stored as bf16."""
    eng, sd = ggca_bn
    crops = W.synthetic_crops(4, seed=42).cuda()
    f = eng.debug_activation(crops, 17).view(4, 7, 7, 512).permute(0, 3, 1, 2).contiguous()
    with torch.no_grad():
        ref = (f * G.ggca(f, sd)).permute(0, 2, 3, 1).contiguous().flatten()
    got = eng.debug_activation(crops, 27)
    assert (got - ref).abs().max().item() <= 1e-2 * ref.abs().max().item()      # bf16 store of the result


@pytest.mark.parametrize("variant", ["default", "bn"])
def test_logits_match_reference_golden(golden_dir, variant, ggca_bn, ggca_default):
    eng, _ = ggca_bn if variant == "bn" else ggca_default
    g = np.load(os.path.join(golden_dir, f"ggca_{variant}.npz"))
    n = int(g["n"])
    crops = W.synthetic_crops(n, seed=int(g["seed_crops"]))
    got = eng.forward_slots(crops.cuda(), torch.arange(n)).cpu().numpy()
    assert np.isfinite(got).all()
    tol = BF16_TOL * max(1.0, np.abs(g["logits"]).max())
    assert np.abs(got - g["logits"]).max() <= tol
    got2 = eng(O.normalize_crops(crops).cuda()).cpu().numpy()
    assert np.abs(got2 - g["logits"]).max() <= tol
    with pytest.raises(RuntimeError):
        eng(torch.zeros((33, 3, 224, 224), device="cuda"))


@pytest.mark.parametrize("variant", ["default", "bn"])
def test_fp32_path_logits(golden_dir, variant):
    """compute_dtype='fp32' (CUDA-core conv stack, gate, encoder): the north-star fp32 gate, 1e-4, against the reference
    class's golden logits; uint8 and reference-style fp32 NCHW input."""
    from fac_fake_b200 import CViTGGCAEngine
    sd = W.make_ggca_state_dict(0, variant)
    eng = CViTGGCAEngine(max_crops=32, compute_dtype="fp32").to("cuda:0").load_state_dict(sd)
    g = np.load(os.path.join(golden_dir, f"ggca_{variant}.npz"))
    n = min(int(g["n"]), 6)
    crops = W.synthetic_crops(int(g["n"]), seed=int(g["seed_crops"]))[:n]
    tol = 1e-4 * max(1.0, np.abs(g["logits"]).max())
    got = eng.forward_slots(crops.cuda(), torch.arange(n)).cpu().numpy()
    assert np.abs(got - g["logits"][:n]).max() <= tol
    got2 = eng(O.normalize_crops(crops).cuda()).cpu().numpy()
    assert np.abs(got2 - g["logits"][:n]).max() <= tol


def test_tokens_and_transformer_match_oracle(ggca_bn):
    """LinearNorm (eps 1e-6) in the MLP branch, nn.LayerNorm (eps 1e-5) in the attention branch."""
    eng, sd = ggca_bn
    crops = W.synthetic_crops(5, seed=43)
    slots = torch.tensor([3, 0, 31, 7, 7])
    with torch.no_grad():
        t = O.embed_tokens(G.gated_features(O.normalize_crops(crops), sd), sd, slots)
        t6 = G.transformer(t, sd)
    xg = crops.cuda()
    got = eng.debug_activation(xg, 18, slots).view(5, 2, 1024)
    assert (got - t).abs().max().item() <= 3e-2 * t.abs().max().item()
    got6 = eng.debug_activation(xg, 24, slots).view(5, 2, 1024)
    assert (got6 - t6).abs().max().item() <= 3e-2 * t6.abs().max().item()


def test_batch_composition_and_predict(ggca_bn):
    eng, sd = ggca_bn
    crops = W.synthetic_crops(70, seed=44)
    xg = crops.cuda()
    slots = torch.arange(70) % 32
    full = eng.forward_slots(xg, slots).cpu()
    for lo, hi in ((0, 1), (5, 8), (33, 66)):
        assert torch.equal(eng.forward_slots(xg[lo:hi], slots[lo:hi]).cpu(), full[lo:hi]), (lo, hi)
    lens = [0, 2, 3, 9]
    offsets = np.concatenate([[0], np.cumsum(lens)]).tolist()
    scores, logits = eng.predict_videos(xg[:offsets[-1]], offsets, return_logits=True)
    scores, logits = scores.cpu(), logits.cpu()
    x = O.normalize_crops(crops[:offsets[-1]])
    for v, ln in enumerate(lens):
        if ln <= 2:
            assert scores[v].item() == 0.5
            continue
        sl = slice(offsets[v], offsets[v + 1])
        ref_logits = G.forward(x[sl], sd)
        assert (logits[sl] - ref_logits).abs().max().item() <= BF16_TOL * max(1.0, ref_logits.abs().max().item())
        ref_score = O.video_score(ref_logits)
        assert abs(scores[v].item() - ref_score) <= 1e-2
        if abs(ref_score - 0.5) > 2e-2:
            assert O.real_or_fake(scores[v].item()) == O.real_or_fake(ref_score)
