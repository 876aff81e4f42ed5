// ff_resvitkan.cu — host side of the ResVitKan variant: the ResNet-50 `features` trunk and the KAN
// head around the shared patch-embedding / ViT encoder of ff_cvit.cu.
//
// Reference: /root/reference/CViT-main/ResVitKan/ResVitKan.py:185-240 (features), :284-329 (CViT.forward),
// kan.py:90-206 (KANLinear).  The stem is its own kernel (ff_rvk.cuh); every bottleneck convolution is one launch of
// rvk_conv2_kernel: 1x1 stride-1 convs "flat" over all pixels of the pass, 3x3 / strided convs as implicit GEMMs whose
// TMA descriptor carries the stride.  Eval-mode BN is folded into (scale, shift) of the producing launch.
#include <algorithm>
#include <cmath>
#include <cstring>

#include "ff_cvit.h"
#include "ff_fp32.cuh"
#include "ff_rvk.cuh"

namespace ffe {

using namespace ff;

namespace {

static_assert(KAN_CHUNKS == KAN_PART_CHUNKS, "kan_part is sized by ff_cvit.cu");
constexpr int kRvkPlanes[4] = {64, 128, 256, 512}, kRvkBlocks[4] = {3, 4, 6, 3}, kRvkStride[4] = {1, 2, 2, 2};

int rvk_fold_bn(ff_cvit* h, const std::string& bn, int c, std::vector<float>* scale, std::vector<float>* shift) {
  const auto* g = get_w(h, bn + ".weight", {c});
  const auto* be = g ? get_w(h, bn + ".bias", {c}) : nullptr;
  const auto* mu = be ? get_w(h, bn + ".running_mean", {c}) : nullptr;
  const auto* var = mu ? get_w(h, bn + ".running_var", {c}) : nullptr;
  if (!g || !be || !mu || !var) return weight_rc(h);
  scale->resize(c);
  shift->resize(c);
  for (int o = 0; o < c; ++o) {
    const float s = (*g)[o] / std::sqrt((*var)[o] + BN_EPS);
    (*scale)[o] = s;
    (*shift)[o] = (*be)[o] - (*mu)[o] * s;                    // the ResNet convolutions have no bias (ResVitKan.py:191,198)
  }
  return FF_OK;
}

int rvk_add_op(ff_cvit* h, const std::string& conv, const std::string& bn, int cin, int cout, int k, int stride, int in_hw,
               int act, int in_buf, int out_buf, int resid) {
  ff_cvit::RvkOp op;
  op.name = conv;
  op.cin = cin; op.cout = cout; op.taps = k * k; op.stride = stride; op.in_hw = in_hw; op.out_hw = in_hw / stride;
  op.type = (k == 1 && stride == 1) ? 0 : 1;
  op.act = act; op.resid = resid; op.in_buf = in_buf; op.out_buf = out_buf;
  op.bn = std::min(cout, 128);
  const auto* w = get_w(h, conv + ".weight", {cout, cin, k, k});
  if (!w) return weight_rc(h);
  std::vector<float> scale, shift, wr((size_t)cout * k * k * cin);
  int rc = rvk_fold_bn(h, bn, cout, &scale, &shift);
  if (rc) return rc;
  for (int o = 0; o < cout; ++o)
    for (int ci = 0; ci < cin; ++ci)
      for (int t = 0; t < k * k; ++t) wr[((size_t)o * k * k + t) * cin + ci] = (*w)[((size_t)o * cin + ci) * k * k + t];
  if ((rc = dev_upload(h, &op.scale, scale))) return rc;
  if ((rc = dev_upload(h, &op.shift, shift))) return rc;
  if (h->compute == FF_COMPUTE_FP32) {           // CUDA-core path: fp32 filters, no tensor maps
    if ((rc = dev_upload(h, &op.wf, wr))) return rc;
    h->rvk_ops.push_back(op);
    return FF_OK;
  }
  if ((rc = dev_upload(h, &op.w, ffh::to_bf16(wr)))) return rc;
  if ((rc = tmap_2d(h, &op.tmB, op.w, (uint64_t)k * k * cin, cout, 64, op.bn))) return rc;
  const bf16* in = h->rvk_buf[in_buf];
  if (op.type == 0) {
    rc = tmap_4d(h, &op.tmA, in, cin, h->cap * in_hw * in_hw, 1, 1, 64, 128, 1, 1);     // [1][1][pixels][cin], boxes of 128 pixels
  } else {
    rvk_tile_geometry(op.out_hw, &op.bw, &op.bh, &op.bi);
    rc = tmap_4d(h, &op.tmA, in, cin, in_hw, in_hw, h->cap, 64, op.bw, op.bh, op.bi, stride);
  }
  if (rc) return rc;
  const bf16* outp = out_buf < 0 ? h->feat : h->rvk_buf[out_buf];
  const int ohw = op.out_hw;
  for (int j = 0; j < 2; ++j) {
    const bf16* base = j == 0 ? outp : (resid >= 0 ? h->rvk_buf[resid] : outp);
    CUtensorMap* m = j == 0 ? &op.tmO : &op.tmR;
    if (op.type == 0) rc = tmap_4d(h, m, base, cout, h->cap * ohw * ohw, 1, 1, 64, 128, 1, 1);
    else rc = tmap_4d(h, m, base, cout, ohw, ohw, h->cap, 64, op.bw, op.bh, op.bi);
    if (rc) return rc;
  }
  h->rvk_ops.push_back(op);
  return FF_OK;
}

}  // namespace

int finalize_rvk_features(ff_cvit* h) {
  int rc;
  // ---- stem: [64][3][7][7] -> [kh][cout][8 px][4 ch] with kw = px - 1 (ff_rvk.cuh)
  {
    const auto* w = get_w(h, "features.conv1.weight", {64, 3, 7, 7});
    if (!w) return weight_rc(h);
    std::vector<float> ws((size_t)7 * 64 * 32, 0.0f), scale, shift;
    for (int kh = 0; kh < 7; ++kh)
      for (int o = 0; o < 64; ++o)
        for (int kw = 0; kw < 7; ++kw)
          for (int c = 0; c < 3; ++c) ws[((size_t)kh * 64 + o) * 32 + (kw + 1) * 4 + c] = (*w)[(((size_t)o * 3 + c) * 7 + kh) * 7 + kw];
    if ((rc = rvk_fold_bn(h, "features.bn1", 64, &scale, &shift))) return rc;
    for (int o = 0; o < 64; ++o) { h->rvk_stem_scale[o] = scale[o]; h->rvk_stem_shift[o] = shift[o]; }
    if (h->compute == FF_COMPUTE_FP32) {
      std::vector<float> wf((size_t)64 * 49 * 4, 0.0f);      // [cout][kh*7+kw][4]
      for (int o = 0; o < 64; ++o)
        for (int t = 0; t < 49; ++t)
          for (int c = 0; c < 3; ++c) wf[((size_t)o * 49 + t) * 4 + c] = (*w)[((size_t)o * 3 + c) * 49 + t];
      if ((rc = dev_upload(h, &h->rvk_stem_wf, wf))) return rc;
      if ((rc = dev_upload(h, &h->rvk_stem_scale_d, scale))) return rc;
      if ((rc = dev_upload(h, &h->rvk_stem_shift_d, shift))) return rc;
    } else {
    if ((rc = dev_upload(h, &h->rvk_stem_w, ffh::to_bf16(ws)))) return rc;
    cuuint64_t dims[3] = {896, 224, (cuuint64_t)h->cap};
    cuuint64_t strides[2] = {896 * 2, (cuuint64_t)224 * 896 * 2};
    cuuint32_t box[3] = {96, 37, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = ffh::encode_tiled()(&h->rvk_tm_x4, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, h->rvk_x4, dims, strides, box, estr,
                                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(h, FF_ERR_CUDA, "cuTensorMapEncodeTiled(stem input) failed: %d", (int)r);
    }
  }
  // ---- bottlenecks.  Buffers: X in {0,1} (block input / output, alternating), 2 = conv1 out, 3 = conv2 out, 4 = downsample
  h->rvk_ops.clear();
  int inplanes = 64, hw = 56, xb = 1;           // the max-pool writes buffer 1
  for (int li = 0; li < 4; ++li) {
    const int planes = kRvkPlanes[li];
    for (int b = 0; b < kRvkBlocks[li]; ++b) {
      const std::string p = "features.layer" + std::to_string(li + 1) + "." + std::to_string(b);
      const int stride = b == 0 ? kRvkStride[li] : 1;
      const int yb = xb ^ 1;
      if ((rc = rvk_add_op(h, p + ".conv1", p + ".bn1", inplanes, planes, 1, 1, hw, 1, xb, 2, -1))) return rc;
      if ((rc = rvk_add_op(h, p + ".conv2", p + ".bn2", planes, planes, 3, stride, hw, 1, 2, 3, -1))) return rc;
      int resid = xb;
      if (b == 0) {
        if ((rc = rvk_add_op(h, p + ".downsample.0", p + ".downsample.1", inplanes, planes * 4, 1, stride, hw, 0, xb, 4, -1))) return rc;
        resid = 4;
      }
      hw /= stride;
      // conv3 + bn3 + ReLU, + residual, + ReLU (ResVitKan.py:169-176: both ReLUs are in the reference)
      if ((rc = rvk_add_op(h, p + ".conv3", p + ".bn3", planes, planes * 4, 1, 1, hw, 1, 3, yb, resid))) return rc;
      inplanes = planes * 4;
      xb = yb;
    }
    h->rvk_layer_end[li] = (int)h->rvk_ops.size() - 1;
  }
  // features.channel + bn2 (no activation) writes the [n][49][512] patch vector the embedding GEMM reads
  if ((rc = rvk_add_op(h, "features.channel", "features.bn2", 2048, 512, 1, 1, 7, 0, xb, -1, -1))) return rc;
  // ---- KAN([2048, 64, 2])
  const int kin[2] = {MLP, 64}, kout[2] = {64, 2};
  for (int l = 0; l < 2; ++l) {
    const std::string q = "kan_head.3.layers." + std::to_string(l);
    const auto* bw = get_w(h, q + ".base_weight", {kout[l], kin[l]});
    const auto* sw = bw ? get_w(h, q + ".spline_weight", {kout[l], kin[l], 8}) : nullptr;
    const auto* gr = sw ? get_w(h, q + ".grid", {kin[l], 12}) : nullptr;
    if (!bw || !sw || !gr) return weight_rc(h);
    // enable_standalone_scale_spline=True is the KANLinear default (kan.py:19); accept checkpoints without the scaler
    const std::vector<float>* sc = nullptr;
    if (h->host_w.count(q + ".spline_scaler")) {
      sc = get_w(h, q + ".spline_scaler", {kout[l], kin[l]});
      if (!sc) return FF_ERR_SHAPE;
    }
    std::vector<float> pk((size_t)kin[l] * 9 * kout[l]);
    for (int i = 0; i < kin[l]; ++i)
      for (int o = 0; o < kout[l]; ++o) {
        pk[((size_t)i * 9) * kout[l] + o] = (*bw)[(size_t)o * kin[l] + i];
        const float s = sc ? (*sc)[(size_t)o * kin[l] + i] : 1.0f;
        for (int k = 0; k < 8; ++k) pk[((size_t)i * 9 + 1 + k) * kout[l] + o] = (*sw)[((size_t)o * kin[l] + i) * 8 + k] * s;
      }
    if ((rc = dev_upload(h, l == 0 ? &h->kan_w0 : &h->kan_w1, pk))) return rc;
    if ((rc = dev_upload(h, l == 0 ? &h->kan_g0 : &h->kan_g1, *gr))) return rc;
  }
  return FF_OK;
}

// One convolution of the trunk (also the GGCA variant's extra BN-less conv) on the persistent TMA-epilogue kernel.
int rvk_launch_op(ff_cvit* h, const ff_cvit::RvkOp& op, int n, cudaStream_t st, int prof_cls) {
  TcArgs a;
  memset(&a, 0, sizeof(a));
  a.scale = op.scale; a.shift = op.shift;
  a.out = op.out_ptr ? op.out_ptr : (op.out_buf < 0 ? h->feat : h->rvk_buf[op.out_buf]);
  a.kb_per_tap = op.cin / 64;
  a.kb_total = op.taps * a.kb_per_tap;
  a.kb_per_split = a.kb_total;
  a.cin = op.cin;
  a.cout = op.cout;
  a.taps = op.taps; a.stride = op.stride;
  a.conv_act = op.act ? 0 : 1;
  int bi = 1;
  if (op.type == 0) {           // flat: W = every pixel of the pass
    a.H = 1; a.W = n * op.out_hw * op.out_hw;
    a.tiles_w = (a.W + 127) / 128; a.tiles_h = 1;
    a.lg_bw = 7; a.lg_bh = 0;
    a.n_img = 1;
  } else {
    a.H = op.out_hw; a.W = op.out_hw;
    a.tiles_w = (op.out_hw + op.bw - 1) / op.bw; a.tiles_h = (op.out_hw + op.bh - 1) / op.bh;
    a.lg_bw = ffh::ilog2(op.bw); a.lg_bh = ffh::ilog2(op.bh);
    a.n_img = n;
    bi = op.bi;
  }
  const int m_tiles = a.tiles_w * a.tiles_h * ((a.n_img + bi - 1) / bi);
  const int tiles = ((m_tiles + 1) / 2) * (op.cout / op.bn);
  const int grid = std::min(tiles, h->num_sms);
  ProfScope ps(h, st, prof_cls);
  cudaError_t e;
  if (op.resid >= 0 && op.bn == 128) e = launch_rvk_conv2<128, 2, true>(grid, st, op.tmA, op.tmB, op.tmO, op.tmR, a);
  else if (op.resid >= 0) return fail(h, FF_ERR_STATE, "%s: residual epilogue needs cout >= 128", op.name.c_str());
  else if (op.bn == 128 && h->act_f16) e = launch_rvk_conv2<128, 3, false, true>(grid, st, op.tmA, op.tmB, op.tmO, op.tmO, a);
  else if (op.bn == 128) e = launch_rvk_conv2<128, 3, false>(grid, st, op.tmA, op.tmB, op.tmO, op.tmO, a);
  else e = launch_rvk_conv2<64, 4, false>(grid, st, op.tmA, op.tmB, op.tmO, op.tmO, a);
  if (e != cudaSuccess) return fail(h, FF_ERR_CUDA, "launch of %s failed: %s", op.name.c_str(), cudaGetErrorString(e));
  ++h->launches;
  return FF_OK;
}

// ---- ResVitKan feature extractor: input conversion, stem, max-pool, 16 bottlenecks, channel conv -> h->feat
int rvk_features(ff_cvit* h, const void* x, int layout, int slot_base, int n, cudaStream_t st, DebugTap* tap) {
  const int stop = tap ? tap->stop_after : 0;
  auto tap_hit = [&](int step, const void* p, int64_t elems) {
    if (stop == step) { tap->ptr = p; tap->elems = elems; tap->is_16 = true; tap->hit = true; return true; }
    return false;
  };
  if (h->h2d_chunks_pending > 0) {     // host-buffer entry point: wait for the chunks covering this pass
    const int c1 = std::min(h->h2d_chunks_pending - 1, (slot_base + n - 1) / h->h2d_chunk);
    FF_CUDA(h, cudaStreamWaitEvent(st, h->h2d_ready[c1], 0));
  }
  {
    ProfScope ps(h, st, KC_CONV1);
    const unsigned blocks = (unsigned)(((size_t)n * 224 * 224 + 255) / 256);
    if (layout == FF_X_NHWC_U8) {
      // (u/255 - mean)/std as one fp32 FMA per channel (cvit_prediction.py:41-45 convention)
      const float mean[3] = {0.485f, 0.456f, 0.406f}, sd[3] = {0.229f, 0.224f, 0.225f};
      rvk_convert_kernel<2><<<(blocks + 3) / 4, 256, 0, st>>>(x, h->rvk_x4, n, 1.0f / (255.0f * sd[0]), -mean[0] / sd[0],
                                                    1.0f / (255.0f * sd[1]), -mean[1] / sd[1], 1.0f / (255.0f * sd[2]), -mean[2] / sd[2]);
    } else {
      rvk_convert_kernel<0><<<blocks, 256, 0, st>>>(x, h->rvk_x4, n, 1.f, 0.f, 1.f, 0.f, 1.f, 0.f);
    }
    FF_LAUNCH_CHECK(h, "rvk_convert");
    RvkStemArgs sa;
    sa.out = h->rvk_buf[0]; sa.w = h->rvk_stem_w; sa.n_img = n;
    for (int o = 0; o < 64; ++o) { sa.scale[o] = h->rvk_stem_scale[o]; sa.shift[o] = h->rvk_stem_shift[o]; }
    const int tiles = 14 * 7 * n;
    cudaError_t e = ffh::launch_smem(rvk_stem_kernel, dim3(std::min(tiles, h->num_sms * 4)), dim3(128), RVK_STEM_SMEM, st, false, h->rvk_tm_x4, sa);
    if (e != cudaSuccess) return fail(h, FF_ERR_CUDA, "launch of the stem failed: %s", cudaGetErrorString(e));
    ++h->launches;
    rvk_maxpool_kernel<<<(unsigned)(((size_t)n * 56 * 56 * 8 + 255) / 256), 256, 0, st>>>(h->rvk_buf[0], h->rvk_buf[1], n);
    FF_LAUNCH_CHECK(h, "rvk_maxpool");
  }
  if (tap_hit(1, h->rvk_buf[1], (int64_t)n * 56 * 56 * 64)) return FF_OK;
  int layer = 0;
  for (size_t i = 0; i < h->rvk_ops.size(); ++i) {
    const ff_cvit::RvkOp& op = h->rvk_ops[i];
    int rc = rvk_launch_op(h, op, n, st, KC_TC_CONV + std::min(layer, 4));
    if (rc) return rc;
    if (layer < 4 && (int)i == h->rvk_layer_end[layer]) {
      ++layer;
      if (tap_hit(1 + layer, h->rvk_buf[op.out_buf], (int64_t)n * op.out_hw * op.out_hw * op.cout)) return FF_OK;
    }
  }
  if (tap_hit(6, h->feat, (int64_t)n * PATCH)) return FF_OK;
  return FF_OK;
}

// The same trunk on the fp32 CUDA-core path (FF_COMPUTE_FP32): one thread per output element, `rvk_f32_chunk` crops at
// a time; featf = [n][49][512] fp32.
int rvk_features_fp32(ff_cvit* h, const void* x, int layout, int n, float* featf, cudaStream_t st) {
  const size_t crop_in_bytes = layout == FF_X_NHWC_U8 ? (size_t)224 * 224 * 3 : (size_t)224 * 224 * 3 * 4;
  auto blocks = [](size_t total) { return (unsigned)std::min<size_t>((total + 255) / 256, 1u << 30); };
  for (int c0 = 0; c0 < n; c0 += h->rvk_f32_chunk) {
    const int ns = std::min(h->rvk_f32_chunk, n - c0);
    const uint8_t* xin = reinterpret_cast<const uint8_t*>(x) + (size_t)c0 * crop_in_bytes;
    nhwc4_f32_kernel<<<blocks((size_t)ns * 224 * 224), 256, 0, st>>>(xin, layout == FF_X_NHWC_U8 ? 2 : 1, reinterpret_cast<float4*>(h->rvk_x4f), ns);
    FF_LAUNCH_CHECK(h, "nhwc4_f32");
    conv_fp32_kernel<<<blocks((size_t)ns * 112 * 112 * 64), 256, 0, st>>>(h->rvk_x4f, h->rvk_stem_wf, h->rvk_stem_scale_d, h->rvk_stem_shift_d,
                                                                         nullptr, h->rvk_f32[0], ns, 224, 112, 4, 64, 7, 2, 1);
    FF_LAUNCH_CHECK(h, "stem_fp32");
    maxpool3s2_f32_kernel<<<blocks((size_t)ns * 56 * 56 * 64), 256, 0, st>>>(h->rvk_f32[0], h->rvk_f32[1], ns, 112, 64);
    FF_LAUNCH_CHECK(h, "maxpool_fp32");
    for (const ff_cvit::RvkOp& op : h->rvk_ops) {
      float* out = op.out_buf < 0 ? featf + (size_t)c0 * PATCH : h->rvk_f32[op.out_buf];
      const int k = op.taps == 9 ? 3 : 1;
      conv_fp32_kernel<<<blocks((size_t)ns * op.out_hw * op.out_hw * op.cout), 256, 0, st>>>(
          h->rvk_f32[op.in_buf], op.wf, op.scale, op.shift, op.resid >= 0 ? h->rvk_f32[op.resid] : nullptr, out, ns, op.in_hw, op.out_hw,
          op.cin, op.cout, k, op.stride, op.act);
      FF_LAUNCH_CHECK(h, op.name.c_str());
    }
  }
  return FF_OK;
}

// kan_head.3 = KAN([2048, 64, 2]) on the ReLU'd hidden vector (ResVitKan.py:302-307, kan.py:90-206)
int kan_head(ff_cvit* h, int n, float* logits, cudaStream_t st) {
  ProfScope ps(h, st, KC_SMALL);
  kan_l0_kernel<<<dim3(KAN_CHUNKS, (n + KAN_SG - 1) / KAN_SG), 256, 0, st>>>(h->hid, h->kan_w0, h->kan_g0, h->kan_part, n, h->cap);
  FF_LAUNCH_CHECK(h, "kan_l0");
  kan_l1_kernel<<<(n + 7) / 8, 256, 0, st>>>(h->kan_part, h->kan_w1, h->kan_g1, logits, n, h->cap);
  FF_LAUNCH_CHECK(h, "kan_l1");
  return FF_OK;
}

}  // namespace ffe
