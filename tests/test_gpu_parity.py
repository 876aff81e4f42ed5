"""GPU (B200): the CUDA path, called through the C-ABI, against the CPU oracle and the golden vectors."""
import os

import numpy as np
import pytest
import torch

from fac_fake_b200 import weights as W
from oracle import cvit_oracle as O

pytestmark = pytest.mark.gpu

BF16_TOL = 2e-2      # north_star: per-frame logits within 2e-2 abs (bf16 path)
FP32_TOL = 1e-4      # north_star: 1e-4 (fp32 path)


def _engine(variant="bn", max_crops=64, compute="bf16", seed=0):
    from fac_fake_b200 import CViTEngine
    sd = W.make_state_dict(seed, variant)
    eng = CViTEngine(max_crops=max_crops, compute_dtype=compute).to("cuda:0").load_state_dict(sd)
    return eng, sd


@pytest.fixture(scope="module")
def eng_bn():
    return _engine("bn")


@pytest.fixture(scope="module")
def eng_default():
    return _engine("default")


def _oracle_layers(sd, x, upto):
    acts = {}
    h = x
    with torch.no_grad():
        for li in range(upto):
            h = O.feature_layer(h, sd, li)
            acts[li + 1] = h.permute(0, 2, 3, 1).contiguous().flatten()
    return acts


def test_conv_layers_match_oracle(eng_bn):
    """Every conv layer (conv+BN+ReLU[+pool]) against the fp32 oracle; bf16 activations => relative gate."""
    eng, sd = eng_bn
    torch.set_num_threads(os.cpu_count() or 4)
    crops = W.synthetic_crops(3, seed=21)
    acts = _oracle_layers(sd, O.normalize_crops(crops), 17)
    xg = crops.cuda()
    for step in range(1, 18):
        got = eng.debug_activation(xg, step)
        ref = acts[step]
        assert got.numel() == ref.numel(), step
        scale = ref.abs().max().item()
        err = (got - ref).abs().max().item()
        assert torch.isfinite(got).all(), step
        # bf16 rounding compounds over layers: 2^-8 per layer budget, never above 3 %
        assert err <= scale * min(0.03, 0.004 * (step + 1)), f"layer {step}: err {err} scale {scale}"


def test_tokens_and_transformer_match_oracle(eng_bn):
    eng, sd = eng_bn
    crops = W.synthetic_crops(5, seed=22)
    x = O.normalize_crops(crops)
    slots = torch.tensor([3, 0, 31, 7, 7])
    with torch.no_grad():
        f = O.features(x, sd)
        t = O.embed_tokens(f, sd, slots)
        xg = crops.cuda()
        got = eng.debug_activation(xg, 18, slots).view(5, 2, 1024)
        assert (got - t).abs().max().item() <= 2e-2 * t.abs().max().item()
        t6 = O.transformer(t, sd)
        got6 = eng.debug_activation(xg, 24, slots).view(5, 2, 1024)
        assert (got6 - t6).abs().max().item() <= 2e-2 * t6.abs().max().item()


@pytest.mark.parametrize("variant", ["default", "bn"])
def test_logits_match_reference_golden(golden_dir, variant, eng_bn, eng_default):
    """Per-frame logits vs the reference's own outputs (tests/golden), 40 crops = chunks [0:32],[32:40]."""
    eng, _ = eng_bn if variant == "bn" else eng_default
    g = np.load(os.path.join(golden_dir, f"cvit_logits_{variant}.npz"))
    crops = W.synthetic_crops(40, seed=1).cuda()
    slots = torch.cat([torch.arange(32), torch.arange(8)])
    got = eng.forward_slots(crops, slots).cpu().numpy()
    assert np.abs(got - g["logits"]).max() <= BF16_TOL
    # reference-compatible call: fp32 NCHW normalised input, b <= 32, slot = batch index
    x = O.normalize_crops(W.synthetic_crops(40, seed=1))[:32].cuda()
    got2 = eng(x).cpu().numpy()
    assert np.abs(got2 - g["logits"][:32]).max() <= BF16_TOL
    with pytest.raises(RuntimeError):
        eng(torch.zeros((33, 3, 224, 224), device="cuda"))


def test_predict_videos_matches_oracle_and_decisions(eng_bn):
    eng, sd = eng_bn
    lens = [0, 1, 2, 3, 15, 30, 33]                 # empty / sentinel / ragged / > 32 (second chunk)
    offsets = np.concatenate([[0], np.cumsum(lens)]).tolist()
    n = offsets[-1]
    crops = W.synthetic_crops(n, seed=23)
    scores, logits = eng.predict_videos(crops.cuda(), offsets, return_logits=True)
    scores, logits = scores.cpu(), logits.cpu()
    x = O.normalize_crops(crops)
    ref_scores, ref_logits = [], []
    for v, ln in enumerate(lens):
        s, lg = O.predict_from_crops(crops[offsets[v]:offsets[v + 1]], sd)
        ref_scores.append(s)
        ref_logits.append(lg)
    ref_logits = torch.cat(ref_logits)
    assert (logits - ref_logits).abs().max().item() <= BF16_TOL
    for v, ln in enumerate(lens):
        if ln <= 2:
            assert scores[v].item() == 0.5
        else:
            assert abs(scores[v].item() - ref_scores[v]) <= 1e-2
            if abs(ref_scores[v] - 0.5) > 2e-2:
                assert O.real_or_fake(scores[v].item()) == O.real_or_fake(ref_scores[v])
    # the reduction kernel alone on the oracle's logits is exact to fp32 rounding
    off_t = torch.tensor(offsets, dtype=torch.int32)
    s2 = eng.video_scores(ref_logits.cuda(), off_t).cpu()
    for v in range(len(lens)):
        assert abs(s2[v].item() - ref_scores[v]) <= 2e-6


def test_reduction_golden(golden_dir, eng_bn):
    eng, _ = eng_bn
    g = np.load(os.path.join(golden_dir, "video_reduction.npz"))
    offsets = np.concatenate([[0], np.cumsum(g["lens"])]).astype(np.int32)
    s = eng.video_scores(torch.from_numpy(g["logits"]).cuda(), torch.from_numpy(offsets)).cpu().numpy()
    assert np.abs(s - g["scores"]).max() <= 2e-6


def test_fp32_path_logits(golden_dir):
    eng, sd = _engine("bn", max_crops=32, compute="fp32")
    g = np.load(os.path.join(golden_dir, "cvit_logits_bn.npz"))
    crops = W.synthetic_crops(40, seed=1)[:6].cuda()
    got = eng.forward_slots(crops, torch.arange(6)).cpu().numpy()
    assert np.abs(got - g["logits"][:6]).max() <= FP32_TOL


def test_full_size_properties(eng_default):
    """BASELINE configs[1] size (512 crops): results do not depend on batch composition, only on (crop, slot)."""
    from fac_fake_b200 import CViTEngine
    sd = W.make_state_dict(0, "default")
    eng = CViTEngine(max_crops=512).to("cuda:0").load_state_dict(sd)
    crops = W.synthetic_crops(512, seed=31).cuda()
    slots = torch.arange(512) % 32
    full = eng.forward_slots(crops, slots)
    assert torch.isfinite(full).all()
    # (a) a crop's logits are identical whether it is evaluated inside the 512 batch or in a small one
    idx = torch.tensor([0, 31, 32, 100, 255, 256, 480, 511])
    sub = eng.forward_slots(crops[idx.cuda()], slots[idx])
    assert (full[idx.cuda()] - sub).abs().max().item() <= 1e-5
    # (b) permuting the batch permutes the logits (slot carried along)
    perm = torch.randperm(512, generator=torch.Generator().manual_seed(0))
    permd = eng.forward_slots(crops[perm.cuda()], slots[perm])
    assert (permd - full[perm.cuda()]).abs().max().item() <= 1e-5
    # (c) ALL 512 logits against the oracle, evaluated the only way the reference can: 16 chunks of 32 (model/cvit.py:175)
    torch.set_num_threads(os.cpu_count() or 4)
    x = O.normalize_crops(crops.cpu())
    with torch.no_grad():
        ref = torch.cat([O.forward(x[c:c + 32], sd) for c in range(0, 512, 32)])
    assert (full.cpu() - ref).abs().max().item() <= BF16_TOL
    # (d) host-buffer entry point gives the same scores as the device one
    offs = list(range(0, 513, 32))
    s_dev = eng.predict_videos(crops, offs).cpu()
    s_host = eng.predict_videos_host(crops.cpu().pin_memory(), offs)
    assert torch.equal(s_dev, s_host)


def test_preprocess_matches_oracle(eng_bn):
    cv2 = pytest.importorskip("cv2")
    from oracle import resize_oracle as R
    eng, _ = eng_bn
    rng = np.random.default_rng(5)
    # integer ratios (incl. 2x2 and non-square), fractional down-scaling, up-scaling in one or both axes, identity,
    # the sample clips' extreme crop sizes (159 / 905 px) and degenerate 1-pixel / 2-row sources
    sizes = [(448, 448), (672, 448), (896, 224), (300, 300), (500, 333), (905, 640), (640, 905), (159, 159), (100, 180),
             (300, 180), (180, 300), (225, 225), (223, 223), (230, 225), (449, 449), (224, 224), (333, 777), (2, 500), (1, 1)]
    srcs = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for h, w in sizes]
    out, norm = eng.preprocess_crops([torch.from_numpy(s).cuda() for s in srcs], swap_rb=True, normalized=True)
    out = out.cpu().numpy()
    for i, s in enumerate(srcs):
        ref = cv2.cvtColor(cv2.resize(s, (224, 224), interpolation=cv2.INTER_AREA), cv2.COLOR_RGB2BGR)
        np.testing.assert_array_equal(out[i], ref, err_msg=str(sizes[i]))           # byte work: bit-exact
        np.testing.assert_array_equal(R.crop_to_model_input(s), ref, err_msg=str(sizes[i]))
    want = O.normalize_crops(torch.from_numpy(out))
    assert (norm.cpu() - want).abs().max().item() <= 1e-6


def test_errors_and_edge_cases(eng_bn):
    from fac_fake_b200 import CViTEngine, EngineError
    eng, _ = eng_bn
    assert eng.forward_slots(torch.zeros((0, 224, 224, 3), dtype=torch.uint8, device="cuda")).shape == (0, 2)
    with pytest.raises(ValueError):
        eng.forward_slots(torch.zeros((2, 3, 100, 100), device="cuda"))
    with pytest.raises(ValueError):
        eng.forward_slots(torch.zeros((2, 224, 224, 3), dtype=torch.uint8, device="cuda"), torch.tensor([0, 32]))
    fresh = CViTEngine().to("cuda:0")
    with pytest.raises(EngineError):
        fresh.forward_slots(torch.zeros((1, 224, 224, 3), dtype=torch.uint8, device="cuda"))
    sd = W.make_state_dict(0, "default")
    bad = dict(sd)
    bad["features.0.weight"] = torch.zeros((32, 3, 5, 5))
    with pytest.raises((ValueError, EngineError)):
        CViTEngine().to("cuda:0").load_state_dict(bad)
    missing = {k: v for k, v in sd.items() if k != "cls_token"}
    with pytest.raises((ValueError, EngineError)):
        CViTEngine().to("cuda:0").load_state_dict(missing)
    assert eng.launch_count() > 0


def test_multi_pass_and_ragged_batches():
    """n > max_crops (several internal passes), n not a multiple of any tile/sub-pass size, n = 1."""
    from fac_fake_b200 import CViTEngine
    sd = W.make_state_dict(0, "bn")
    small = CViTEngine(max_crops=32).to("cuda:0").load_state_dict(sd)     # forces 3 passes for 71 crops
    big = CViTEngine(max_crops=128).to("cuda:0").load_state_dict(sd)
    crops = W.synthetic_crops(71, seed=41).cuda()
    slots = (torch.arange(71) * 7) % 32
    a = small.forward_slots(crops, slots)
    b = big.forward_slots(crops, slots)
    assert torch.isfinite(a).all()
    assert (a - b).abs().max().item() <= 1e-5          # pass structure does not change results
    one = big.forward_slots(crops[5:6], slots[5:6])
    assert (one - b[5:6]).abs().max().item() <= 1e-5
    x = O.normalize_crops(crops[:9].cpu())
    ref = O.forward_slots(x, sd, slots[:9])
    assert (b[:9].cpu() - ref).abs().max().item() <= BF16_TOL
    # ragged videos through the fused entry point, crossing the internal pass boundary
    lens = [3, 30, 0, 17, 21]
    offs = np.concatenate([[0], np.cumsum(lens)]).tolist()
    s_small = small.predict_videos(crops, offs).cpu()
    s_big = big.predict_videos(crops, offs).cpu()
    assert torch.equal(s_small, s_big)
    assert s_big[2].item() == 0.5


def test_softmax_mean_mode(eng_bn):
    eng, _ = eng_bn
    g = torch.Generator().manual_seed(3)
    logits = torch.randn((50, 2), generator=g) * 2
    offs = torch.tensor([0, 7, 7, 30, 50], dtype=torch.int32)
    s = eng.video_scores(logits.cuda(), offs, mode=1).cpu()
    p = torch.softmax(logits, dim=1)[:, 0]
    want = [p[0:7].mean().item(), 0.5, p[7:30].mean().item(), p[30:50].mean().item()]
    for a, b in zip(s.tolist(), want):
        assert abs(a - b) <= 2e-6


def test_prediction_api_mirror_end_to_end(tmp_path, eng_bn):
    """fac_fake_b200.cvit_prediction.predict(): video file in -> score out, same sampling/chunking as the reference
    (cvit_prediction.py:153-242), with an injected face extractor; checked against the oracle on the same crops."""
    cv2 = pytest.importorskip("cv2")
    import fac_fake_b200.cvit_prediction as cp
    eng, sd = eng_bn
    path = str(tmp_path / "clip.avi")
    rng = np.random.default_rng(7)
    vw = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"MJPG"), 25.0, (320, 240))
    if not vw.isOpened():
        pytest.skip("no MJPG encoder in this OpenCV build")
    base = rng.integers(0, 256, (240, 320, 3), dtype=np.uint8)
    for i in range(120):
        vw.write(np.roll(base, 3 * i, axis=1))
    vw.release()
    seen = []

    def extractor(frame, _unused=None):
        # two "faces" per frame: fixed boxes, resized exactly like cvit_prediction.py:113-115
        crops = []
        for (t, r, b, l) in ((20, 200, 180, 40), (60, 300, 220, 140)):
            c = cv2.resize(frame[t:b, l:r], (224, 224), interpolation=cv2.INTER_AREA)
            crops.append(cv2.cvtColor(c, cv2.COLOR_RGB2BGR))
        out = np.stack(crops)
        seen.append(out)
        return out, len(crops)

    cp.configure(model_=eng, face_extractor=extractor, sample_dir=str(tmp_path))
    score = cp.predict_on_video(["clip.avi"], num_workers=1)[0]
    # the reference loop: int(0.1 * n_frames) iterations, 2 crops each, capped at 29 crops
    assert len(seen) == 12
    crops = np.concatenate(seen)[:29]
    ref_score, _ = O.predict_from_crops(torch.from_numpy(crops), sd)
    assert abs(score - ref_score) <= 1e-2
    assert cp.real_or_fake(score) in ("REAL", "FAKE")
    # sentinel: no faces -> 0.5 (cvit_prediction.py:218-219)
    cp.configure(face_extractor=lambda frame, _u=None: ([], 0))
    assert cp.predict_on_video(["clip.avi"], num_workers=1)[0] == 0.5
    # helper parity with the reference's reduction helpers
    lg = torch.randn((9, 2), generator=torch.Generator().manual_seed(1))
    got = cp.pre_process_prediction(cp.pred_sig(lg)).item()
    assert abs(got - O.video_score(lg)) <= 1e-5


def test_fused_layer12_kernel_and_separate_conv1_agree():
    """Feature layers 1 and 2 run in ONE kernel on the uint8 path (conv1's output never leaves shared memory).  Its
    conv2 output against the oracle, incl. the image borders (zero padding of conv1's OUTPUT), a single crop (fewer
    tiles than CTAs) and a ragged count; debug tap 1 runs the stand-alone conv1 kernel, which must match the oracle
    too; and the fp32-NCHW entry (conv1_f32_kernel + the stand-alone conv2) must agree with the fused uint8 entry."""
    eng, sd = _engine("bn", max_crops=64)
    for n, seed in ((3, 28), (1, 29), (5, 30)):
        crops = W.synthetic_crops(n, seed=seed)
        x = O.normalize_crops(crops)
        acts = _oracle_layers(sd, x, 3)
        xg = crops.cuda()
        for step in (1, 2, 3):
            got, ref = eng.debug_activation(xg, step), acts[step]
            assert (got - ref).abs().max().item() <= ref.abs().max().item() * 0.004 * (step + 1), (n, step)
        got_f = eng.debug_activation(x.cuda(), 2)
        assert (got_f - eng.debug_activation(xg, 2)).abs().max().item() <= acts[2].abs().max().item() * 0.008, n


def test_encoder_kernel_tiles_and_launch_count():
    """The six transformer layers run as ONE cooperative kernel (ff_xf.cuh: 16-CTA groups per 128-row token tile).
    Against the fp32 oracle after every layer, and for batch shapes that exercise a partial last tile, a single crop
    and more tiles than co-resident groups (640 crops = 10 tiles on 9 groups) through batch-composition invariance."""
    eng, sd = _engine("bn", max_crops=640)
    crops = W.synthetic_crops(5, seed=31)
    slots = torch.tensor([3, 0, 31, 7, 7])
    with torch.no_grad():
        t = O.embed_tokens(O.features(O.normalize_crops(crops), sd), sd, slots)
        for depth in range(1, 7):          # residual stream after transformer layer `depth`
            ref = O.transformer(t, sd, depth=depth)
            got = eng.debug_activation(crops.cuda(), 18 + depth, slots).view(5, 2, 1024)
            assert torch.isfinite(got).all()
            assert (got - ref).abs().max().item() <= 2e-2 * max(1.0, ref.abs().max().item()), depth
    big = W.synthetic_crops(640, seed=33).cuda()
    before = eng.launch_count()
    full = eng.forward_slots(big).cpu()
    per_call = eng.launch_count() - before
    assert torch.isfinite(full).all()
    for n in (1, 31, 64, 65, 200):
        part = eng.forward_slots(big[:n]).cpu()
        assert (part - full[:n]).abs().max().item() <= 1e-5, n
    # 640 crops = 10 sub-passes x (fused layer-1/2 + 4) + 11 + embed, tokens, ONE encoder launch, cls, head1, head2
    assert per_call == 10 * 5 + 11 + 6, per_call


def test_encoder_folded_layernorm_with_a_large_row_mean():
    """The encoder kernel folds LayerNorm into to_qkv / net.0 and feeds the GEMMs bf16(x - previous row mean) (ff_xf.cuh).
    A residual stream whose rows sit 6 sigma away from zero (pos_embedding + 6) must stay as close to the fp32 oracle as
    a centred one: the shift, not the raw value, sets the bf16 rounding error."""
    from fac_fake_b200 import CViTEngine
    sd = {k: v.clone() for k, v in W.make_state_dict(0, "bn").items()}
    sd["pos_embedding"] = sd["pos_embedding"] + 6.0
    eng = CViTEngine(max_crops=32).to("cuda:0").load_state_dict(sd)
    crops = W.synthetic_crops(4, seed=41)
    slots = torch.tensor([0, 5, 17, 31])
    with torch.no_grad():
        t = O.embed_tokens(O.features(O.normalize_crops(crops), sd), sd, slots)
        assert t.mean().abs().item() > 4.0
        for depth in (1, 3, 6):
            ref = O.transformer(t, sd, depth=depth)
            got = eng.debug_activation(crops.cuda(), 18 + depth, slots).view(4, 2, 1024)
            centred = ref - ref.mean(-1, keepdim=True)
            # gate on the scale of the deviations from the row mean (what LayerNorm sees), not on the 6-sigma offset
            assert (got - ref).abs().max().item() <= 2e-2 * max(1.0, centred.abs().max().item()), depth
    logits = eng.forward_slots(crops.cuda(), slots).cpu()
    ref_logits = O.forward_slots(O.normalize_crops(crops), sd, slots)
    assert (logits - ref_logits).abs().max().item() <= BF16_TOL


def test_repeated_mixed_size_calls_are_bitwise_reproducible():
    """One handle, 48 calls of mixed batch sizes (1 ... 640 crops: partial tiles, more tiles than co-resident encoder
    groups): every repetition of a size returns the same bits.  The encoder's group barriers, its alternating statistics
    buffers and the self-re-arming counters carry state from launch to launch; a protocol race would show up here."""
    eng, _ = _engine("bn", max_crops=640)
    ref = {}
    for it in range(48):
        n = [1, 31, 32, 64, 65, 200, 512, 640][it % 8]
        out = eng.forward_slots(W.synthetic_crops(n, seed=n).cuda()).cpu()
        assert torch.isfinite(out).all()
        if n in ref:
            assert torch.equal(ref[n], out), (it, n)
        else:
            ref[n] = out


def test_cta_pair_conv_kernel_odd_tile_counts():
    """Feature layers 7..9 run on ptcw_conv_kernel (128 channels x 256 pixels = two 8 x 8 x 2-image sub-tiles per CTA), layers
    10..17 on ptc2_conv_kernel (cta_group::2, 256 pixels x 256 channels per CTA pair).  33 crops give an odd image count (the
    second image of the last sub-tile does not exist) and odd pixel-tile counts (the second sub-tile / the pair's second tile
    falls off the end): activations against the oracle."""
    eng, sd = _engine("bn", max_crops=64)
    crops = W.synthetic_crops(33, seed=74)
    acts = _oracle_layers(sd, O.normalize_crops(crops), 17)
    xg = crops.cuda()
    for step in range(7, 18):
        got = eng.debug_activation(xg, step)
        assert torch.isfinite(got).all(), step
        assert (got - acts[step]).abs().max().item() <= acts[step].abs().max().item() * min(0.03, 0.004 * (step + 1)), step


def test_long_videos_follow_the_reference_chunking():
    """cvit_prediction.py:224-238: chunks [0:32], [32:64], [64:90]; frames >= 90 never reach the model.  65-, 90- and
    95-frame videos through the C-ABI predict against the oracle's restatement of those lines."""
    eng, sd = _engine("bn", max_crops=128)
    torch.set_num_threads(os.cpu_count() or 4)
    lens = [65, 90, 95, 3]
    offsets = np.concatenate([[0], np.cumsum(lens)]).tolist()
    crops = W.synthetic_crops(offsets[-1], seed=61)
    scores, logits = eng.predict_videos(crops.cuda(), offsets, return_logits=True)
    scores, logits = scores.cpu(), logits.cpu()
    for v, ln in enumerate(lens):
        ref_score, ref_logits = O.predict_from_crops(crops[offsets[v]:offsets[v + 1]], sd)
        kept = min(ln, 90)
        assert ref_logits.shape[0] == kept
        assert (logits[offsets[v]:offsets[v] + kept] - ref_logits).abs().max().item() <= BF16_TOL, v
        assert abs(scores[v].item() - ref_score) <= 1e-2, v
        # the kernel's own logits through the oracle's reduction: the score must ignore frames >= 90 exactly
        assert abs(scores[v].item() - O.video_score(logits[offsets[v]:offsets[v] + kept])) <= 2e-6, v
    # the mirror module takes the same path (predict_crops -> ff_cvit_predict)
    from fac_fake_b200 import cvit_prediction as cp
    cp.configure(model_=eng)
    got = cp.predict_crops(crops[offsets[2]:offsets[3]].numpy())
    assert abs(got - scores[2].item()) <= 1e-6


def test_config2_video_decisions_at_scale():
    """BASELINE configs[2] shape on one GPU: 64 videos x 30 frames, video_offsets = 30 * arange(65), slot = frame index,
    passes cut at 512 crops (17 videos + 2 frames: videos straddle pass boundaries).  Logits, scores and REAL/FAKE
    decisions against the oracle for every video.  The "decisive" weights (see weights.make_state_dict) spread the scores
    over both sides of the 0.5 threshold so that the decision comparison means something; their last layer is scaled by
    60 and so is the logit gate."""
    from fac_fake_b200 import CViTEngine
    torch.set_num_threads(os.cpu_count() or 4)
    nv, fr = 64, 30
    crops = torch.cat([W.synthetic_video_crops(v, fr) for v in range(nv)])             # per-video seed = video id
    offsets = list(range(0, nv * fr + 1, fr))
    x = O.normalize_crops(crops)
    # (variant, logit gate, score gate = decision margin): the x60 head turns a 3e-3 logit error into up to 0.2, i.e. up to
    # 0.05 in one frame's sigmoid and ~1e-2 in a 30-frame mean
    for variant, tol, stol in (("bn", BF16_TOL, 1e-2), ("decisive", 60 * BF16_TOL, 3e-2)):
        sd = W.make_state_dict(0, variant)
        eng = CViTEngine(max_crops=512).to("cuda:0").load_state_dict(sd)
        scores, logits = eng.predict_videos(crops.cuda(), offsets, return_logits=True)
        host_scores = eng.predict_videos_host(crops.pin_memory(), offsets)
        assert torch.equal(scores.cpu(), host_scores)
        with torch.no_grad():
            ref = torch.cat([O.forward(x[o:o + fr], sd) for o in offsets[:-1]])
        assert (logits.cpu() - ref).abs().max().item() <= tol, variant
        ref_scores = O.video_scores(ref, offsets)
        labels = []
        for v in range(nv):
            # the reference rule `f if f > r else |1 - r|` (cvit_prediction.py:274-279) jumps where the two means cross:
            # a video whose reference means are within the gate of each other has no well-defined score to compare
            pm = torch.sigmoid(ref[offsets[v]:offsets[v + 1]]).mean(0)
            if abs(float(pm[0] - pm[1])) <= stol:
                continue
            assert abs(scores[v].item() - ref_scores[v]) <= stol, (variant, v)
            if abs(ref_scores[v] - 0.5) > stol:      # a video whose reference score IS the threshold has no decision
                assert O.real_or_fake(scores[v].item()) == O.real_or_fake(ref_scores[v]), (variant, v)
                labels.append(O.real_or_fake(ref_scores[v]))
        if variant == "decisive":                    # the comparison is not vacuous: both labels occur, few are skipped
            assert len(labels) >= nv // 2, len(labels)
            assert labels.count("FAKE") >= 8 and labels.count("REAL") >= 8, (labels.count("FAKE"), labels.count("REAL"))
        del eng


def test_strict_state_dict_and_shape_errors():
    """load_state_dict(strict=True) reports keys the model does not have; shape errors are ValueError, missing keys
    EngineError — like nn.Module.load_state_dict tells the two apart."""
    from fac_fake_b200 import CViTEngine, EngineError
    sd = W.make_state_dict(0, "default")
    extra = dict(sd)
    extra["mlp_head.3.weight"] = torch.zeros((2, 2))
    with pytest.raises(EngineError, match="Unexpected key"):
        CViTEngine(max_crops=32).to("cuda:0").load_state_dict(extra)
    CViTEngine(max_crops=32).to("cuda:0").load_state_dict(extra, strict=False)
    bad = dict(sd)
    bad["transformer.layers.0.0.fn.fn.to_out.bias"] = torch.zeros((7,))
    with pytest.raises(ValueError):
        CViTEngine(max_crops=32).to("cuda:0").load_state_dict(bad)
    eng = CViTEngine(max_crops=32).to("cuda:0").load_state_dict(sd)
    for wrong in (torch.zeros((2, 224, 224, 3), dtype=torch.float16, device="cuda"),
                  torch.zeros((2, 224, 224, 3), dtype=torch.float64, device="cuda"),
                  torch.zeros((2, 3, 224, 224), dtype=torch.uint8, device="cuda")):
        with pytest.raises(ValueError):
            eng.predict_videos(wrong, [0, 2])
        with pytest.raises(ValueError):
            eng.debug_activation(wrong, 2)
    with pytest.raises(ValueError):
        eng.predict_videos(torch.zeros((4, 224, 224, 3), dtype=torch.uint8, device="cuda"), [0, 3, 2])


def test_calls_leave_the_current_device_alone_and_threads_are_serialised():
    """Every C-ABI entry point restores the caller's device; concurrent predict_host calls on ONE handle (the
    reference drives predict() from a ThreadPoolExecutor, cvit_prediction.py:73-83) give each caller its own scores."""
    from concurrent.futures import ThreadPoolExecutor
    eng, sd = _engine("bn", max_crops=64)
    dev0 = torch.cuda.current_device()
    batches = [W.synthetic_crops(9 + i, seed=90 + i).pin_memory() for i in range(6)]
    want = [eng.predict_videos(b.cuda(), [0, b.shape[0]]).cpu() for b in batches]
    with ThreadPoolExecutor(max_workers=4) as ex:
        got = list(ex.map(lambda b: eng.predict_videos_host(b, [0, b.shape[0]]), batches * 3))
    for i, g in enumerate(got):
        assert torch.equal(g, want[i % len(batches)]), i
    assert torch.cuda.current_device() == dev0
    if torch.cuda.device_count() > 1:       # a second engine on another GPU of the same process
        from fac_fake_b200 import CViTEngine
        e1 = CViTEngine(max_crops=32).to("cuda:1").load_state_dict(sd)
        c = W.synthetic_crops(4, seed=5)
        a = eng.forward_slots(c.cuda()).cpu()
        b = e1.forward_slots(c.to("cuda:1")).cpu()
        assert torch.equal(a, b)
        assert torch.cuda.current_device() == dev0
        del e1
        assert torch.cuda.current_device() == dev0
