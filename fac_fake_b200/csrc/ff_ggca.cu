// ff_ggca.cu — host side of the `cvit_GGCA_ADD_DEConv_RepBn8` variant: DEConv folding, the extra
// BN-less Conv2d(128,128) and the GGCA gate around the CViT plan of ff_cvit.cu.
//
// Reference: /root/reference/CViT-main/model/cvit_GGCA_ADD_DEConv_RepBn8.py:353-455 (CViT), :329-351 (DEConv),
// :143-213 (GGCA), :22-60 (LinearNorm).
#include <cmath>
#include <cstring>

#include "ff_cvit.h"
#include "ff_rvk.cuh"

namespace ffe {

using namespace ff;

// cvit_GGCA_ADD_DEConv_RepBn8.py:361-423: (sequential, conv index, is DEConv, BN index or -1).  Entry 8 is the extra
// Conv2d(128,128) without BN / activation (features1.26); entry 9 the BN-less DEConv(128) + ReLU + pool (features1.27).
const GgcaPlan kGgcaPlan[18] = {
    {"features1", 0, false, 1},  {"features1", 3, true, 4},   {"features1", 6, true, 7},   {"features1", 10, false, 11},
    {"features1", 13, true, 14}, {"features1", 16, true, 17}, {"features1", 20, false, 21}, {"features1", 23, true, 24},
    {"features1", 26, false, -1}, {"features1", 27, true, -1}, {"features1", 30, false, 31}, {"features1", 33, true, 34},
    {"features1", 36, true, 37}, {"features1", 39, true, 40}, {"features2", 0, false, 1},  {"features2", 3, true, 4},
    {"features2", 6, true, 7},   {"features2", 9, true, 10},
};

// [cout][cin][3][3] kernel and bias of a plan entry.  A DEConv (:329-351) is folded exactly as its forward does:
// central difference (centre tap minus the tap sum, :218-235), horizontal / vertical difference built from Conv1d
// weights (:290-326), angular difference w - w[perm] (:238-255, theta = 1) and a plain 3x3 kernel; biases add.
int ggca_conv_weights(ff_cvit* h, const GgcaPlan& gp, int cin, int cout, std::vector<float>* w_out, std::vector<float>* b_out) {
  const std::string p = std::string(gp.seq) + "." + std::to_string(gp.conv_idx);
  if (!gp.de) {
    const auto* w = get_w(h, p + ".weight", {cout, cin, 3, 3});
    const auto* b = w ? get_w(h, p + ".bias", {cout}) : nullptr;
    if (!w || !b) return weight_rc(h);
    *w_out = *w;
    *b_out = *b;
    return FF_OK;
  }
  const std::vector<float>* ws[5];
  const std::vector<float>* bs[5];
  const char* wn[5] = {".conv1_1.conv.weight", ".conv1_2.conv.weight", ".conv1_3.conv.weight", ".conv1_4.conv.weight", ".conv1_5.weight"};
  const char* bnm[5] = {".conv1_1.conv.bias", ".conv1_2.conv.bias", ".conv1_3.conv.bias", ".conv1_4.conv.bias", ".conv1_5.bias"};
  for (int i = 0; i < 5; ++i) {
    ws[i] = (i == 1 || i == 2) ? get_w(h, p + wn[i], {cout, cin, 3}) : get_w(h, p + wn[i], {cout, cin, 3, 3});
    if (!ws[i]) return weight_rc(h);
    bs[i] = get_w(h, p + bnm[i], {cout});
    if (!bs[i]) return weight_rc(h);
  }
  const std::vector<float>&w1 = *ws[0], &w2 = *ws[1], &w3 = *ws[2], &w4 = *ws[3], &w5 = *ws[4];
  static const int perm[9] = {3, 0, 1, 6, 4, 2, 7, 8, 5};
  w_out->assign((size_t)cout * cin * 9, 0.0f);
  b_out->resize(cout);
  for (size_t oi = 0; oi < (size_t)cout * cin; ++oi) {
    const float* a1 = &w1[oi * 9];
    const float* a2 = &w2[oi * 3];
    const float* a3 = &w3[oi * 3];
    const float* a4 = &w4[oi * 9];
    const float* a5 = &w5[oi * 9];
    float sum1 = 0.0f;
    for (int t = 0; t < 9; ++t) sum1 += a1[t];
    for (int t = 0; t < 9; ++t) {
      float cd = a1[t];
      if (t == 4) cd = a1[4] - sum1;
      float hd = 0.0f, vd = 0.0f;
      if (t % 3 == 0) hd = a2[t / 3];                 // taps 0,3,6 = +w, taps 2,5,8 = -w
      else if (t % 3 == 2) hd = -a2[t / 3];
      if (t < 3) vd = a3[t];                          // taps 0,1,2 = +w, taps 6,7,8 = -w
      else if (t >= 6) vd = -a3[t - 6];
      const float ad = a4[t] - a4[perm[t]];
      (*w_out)[oi * 9 + t] = (((cd + hd) + vd) + ad) + a5[t];     // the reference's summation order: w1 + w2 + w3 + w4 + w5
    }
  }
  for (int o = 0; o < cout; ++o) (*b_out)[o] = ((((*bs[0])[o] + (*bs[1])[o]) + (*bs[2])[o]) + (*bs[3])[o]) + (*bs[4])[o];
  return FF_OK;
}

// The BN-less, activation-less Conv2d(128,128) of features1.26 (P -> bufR) as one rvk_conv2_kernel op, and the GGCA
// shared convs with their BatchNorm2d(8) folded into the first one.
int finalize_ggca_extras(ff_cvit* h) {
  int rc;
  std::vector<float> wsrc, bias;
  if ((rc = ggca_conv_weights(h, kGgcaPlan[8], 128, 128, &wsrc, &bias))) return rc;
  ff_cvit::RvkOp op;
  op.name = "features1.26";
  op.type = 1; op.cin = 128; op.cout = 128; op.taps = 9; op.stride = 1; op.in_hw = 56; op.out_hw = 56;
  op.act = 0; op.resid = -1; op.bn = 128; op.bw = 8; op.bh = 8; op.bi = 2;
  op.out_ptr = h->bufR;
  std::vector<float> wr((size_t)128 * 9 * 128), ones(128, 1.0f);
  for (int o = 0; o < 128; ++o)
    for (int ci = 0; ci < 128; ++ci)
      for (int t = 0; t < 9; ++t) wr[((size_t)o * 9 + t) * 128 + ci] = wsrc[((size_t)o * 128 + ci) * 9 + t];
  if ((rc = dev_upload(h, &op.scale, ones))) return rc;
  if ((rc = dev_upload(h, &op.shift, bias))) return rc;
  if (h->compute == FF_COMPUTE_FP32) {
    if ((rc = dev_upload(h, &op.wf, wr))) return rc;
  } else {
    if ((rc = dev_upload(h, &op.w, to_act16(h, wr)))) return rc;
    if ((rc = tmap_2d(h, &op.tmB, op.w, 9 * 128, 128, 64, 128))) return rc;
    if ((rc = tmap_4d(h, &op.tmA, conv_output_buffer(h, 7), 128, 56, 56, h->cap, 64, 8, 8, 2))) return rc;
    if ((rc = tmap_4d(h, &op.tmO, h->bufR, 128, 56, 56, h->cap, 64, 8, 8, 2))) return rc;
    op.tmR = op.tmO;
  }
  h->rvk_ops.clear();
  h->rvk_ops.push_back(op);
  // GGCA(512,7,7).shared_conv: Conv2d(128,8,1) + BatchNorm2d(8) + ReLU + Conv2d(8,128,1)  (:159-166)
  const auto* pw1 = get_w(h, "ggca.shared_conv.0.weight", {8, 128, 1, 1});
  const auto* pb1 = pw1 ? get_w(h, "ggca.shared_conv.0.bias", {8}) : nullptr;
  const auto* pg = pb1 ? get_w(h, "ggca.shared_conv.1.weight", {8}) : nullptr;
  const auto* pbe = pg ? get_w(h, "ggca.shared_conv.1.bias", {8}) : nullptr;
  const auto* pmu = pbe ? get_w(h, "ggca.shared_conv.1.running_mean", {8}) : nullptr;
  const auto* pvar = pmu ? get_w(h, "ggca.shared_conv.1.running_var", {8}) : nullptr;
  const auto* pw2 = pvar ? get_w(h, "ggca.shared_conv.3.weight", {128, 8, 1, 1}) : nullptr;
  const auto* pb2 = pw2 ? get_w(h, "ggca.shared_conv.3.bias", {128}) : nullptr;
  if (!pb2) return weight_rc(h);
  const std::vector<float>&w1 = *pw1, &b1 = *pb1, &g = *pg, &be = *pbe, &mu = *pmu, &var = *pvar, &w2 = *pw2, &b2 = *pb2;
  std::vector<float> fw1(8 * 128), fb1(8);
  for (int u = 0; u < 8; ++u) {
    const float s = g[u] / std::sqrt(var[u] + BN_EPS);
    for (int k = 0; k < 128; ++k) fw1[u * 128 + k] = w1[u * 128 + k] * s;
    fb1[u] = (b1[u] - mu[u]) * s + be[u];
  }
  if ((rc = dev_upload(h, &h->ggca_w1, fw1))) return rc;
  if ((rc = dev_upload(h, &h->ggca_b1, fb1))) return rc;
  if ((rc = dev_upload(h, &h->ggca_w2, w2))) return rc;
  if ((rc = dev_upload(h, &h->ggca_b2, b2))) return rc;
  return FF_OK;
}

// x = x * ggca(x)  (cvit_GGCA_ADD_DEConv_RepBn8.py:447-448), in place on the [n,7,7,512] feature map
int ggca_gate(ff_cvit* h, int n, cudaStream_t st) {
  ProfScope ps(h, st, KC_SMALL);
  if (h->act_f16) ggca_gate_kernel<1><<<n, 512, 0, st>>>(h->feat, h->ggca_w1, h->ggca_b1, h->ggca_w2, h->ggca_b2, n);
  else ggca_gate_kernel<0><<<n, 512, 0, st>>>(h->feat, h->ggca_w1, h->ggca_b1, h->ggca_w2, h->ggca_b2, n);
  FF_LAUNCH_CHECK(h, "ggca_gate");
  return FF_OK;
}

int ggca_gate_fp32(ff_cvit* h, float* featf, int n, cudaStream_t st) {
  ggca_gate_kernel<2><<<n, 512, 0, st>>>(featf, h->ggca_w1, h->ggca_b1, h->ggca_w2, h->ggca_b2, n);
  FF_LAUNCH_CHECK(h, "ggca_gate_fp32");
  return FF_OK;
}

}  // namespace ffe
