// ff_cvit.cu — the CViT engine of libfacfake.so: weight folding/layout, workspace, TMA descriptors, the launch
// schedule of the forward pass and the C-ABI declared in include/facfake.h.
//
// Reference path being replaced (all under /root/reference/CViT-main/):
//   model/cvit.py:80-179 (CViT), cvit_prediction.py:209-242 (model half of predict()), :258-281 (reduction).
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>

#include "ff_cvit.h"
#include "ff_fp32.cuh"
#include "ff_pre.cuh"
#include "ff_small.cuh"
#include "ff_tc.cuh"
#include "ff_ws.cuh"
#include "ff_c1.cuh"
#include "ff_c12.cuh"
#include "ff_xf.cuh"
#include "ff_ptc2.cuh"
#include "ff_ptcw.cuh"

namespace ffe {

using namespace ff;

const ConvPlan kConv[17] = {
    {3, 32, 224, false, 0},    {32, 32, 224, false, 3},   {32, 32, 224, true, 6},
    {32, 64, 112, false, 10},  {64, 64, 112, false, 13},  {64, 64, 112, true, 16},
    {64, 128, 56, false, 20},  {128, 128, 56, false, 23}, {128, 128, 56, true, 26},
    {128, 256, 28, false, 30}, {256, 256, 28, false, 33}, {256, 256, 28, false, 36}, {256, 256, 28, true, 39},
    {256, 512, 14, false, 43}, {512, 512, 14, false, 46}, {512, 512, 14, false, 49}, {512, 512, 14, true, 52},
};

namespace {
std::string g_create_error;
}

int fail(const ff_cvit* h, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (h) h->err = buf; else g_create_error = buf;
  return code;
}

void prof_mark(ff_cvit* h, cudaStream_t st, int cls, bool begin, bool coarse) {
  if (!h->profiling || (h->prof_coarse != coarse)) return;
  if (h->ev_used >= h->ev_pool.size()) {
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) return;
    h->ev_pool.push_back(e);
  }
  cudaEventRecord(h->ev_pool[h->ev_used++], st);
  if (begin) h->ev_class.push_back(cls);
}

std::vector<bf16> to_act16(const ff_cvit* h, const std::vector<float>& v) {
  return h->act_f16 ? ffh::to_f16_bits(v) : ffh::to_bf16(v);
}

// ------------------------------------------------------------------------------------------------ TMA descriptors
int tmap_2d(ff_cvit* h, CUtensorMap* m, const void* base, uint64_t inner, uint64_t rows, uint32_t box_inner,
            uint32_t box_rows) {
  cuuint64_t dims[2] = {inner, rows};
  cuuint64_t strides[1] = {inner * 2};
  cuuint32_t box[2] = {box_inner, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMapSwizzle sw = box_inner * 2 == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  CUresult r = ffh::encode_tiled()(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                                   CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(h, FF_ERR_CUDA, "cuTensorMapEncodeTiled(2d %llu x %llu) failed: %d",
                                     (unsigned long long)inner, (unsigned long long)rows, (int)r);
  return FF_OK;
}
int tmap_4d(ff_cvit* h, CUtensorMap* m, const void* base, int C, int W, int H, int N, int boxC, int bw, int bh, int bi,
            int estride) {
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  // with element strides the box is given in input-space extents: ceil(box/stride) elements are loaded per dim
  cuuint32_t box[4] = {(cuuint32_t)boxC, (cuuint32_t)(bw * estride), (cuuint32_t)(bh * estride), (cuuint32_t)bi};
  cuuint32_t estr[4] = {1, (cuuint32_t)estride, (cuuint32_t)estride, 1};
  CUtensorMapSwizzle sw = boxC * 2 == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  CUresult r = ffh::encode_tiled()(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                                   CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(h, FF_ERR_CUDA, "cuTensorMapEncodeTiled(4d C%d W%d H%d N%d) failed: %d", C, W, H, N, (int)r);
  return FF_OK;
}

// ------------------------------------------------------------------------------------------------ weights
const std::vector<float>* get_w(ff_cvit* h, const std::string& key, std::initializer_list<int64_t> shape) {
  auto it = h->host_w.find(key);
  if (it == h->host_w.end()) {
    fail(h, FF_ERR_STATE, "missing weight '%s'", key.c_str());
    return nullptr;
  }
  const auto& s = h->host_shape[key];
  if (s.size() != shape.size() || !std::equal(s.begin(), s.end(), shape.begin())) {
    fail(h, FF_ERR_SHAPE, "weight '%s' has the wrong shape", key.c_str());
    return nullptr;
  }
  h->host_used[key] = true;
  return &it->second;
}
int weight_rc(const ff_cvit* h) { return h->err.find("wrong shape") != std::string::npos ? FF_ERR_SHAPE : FF_ERR_STATE; }

int upload_linear(ff_cvit* h, LinearDev* L, const std::string& name, int out_f, int in_f, bool bias, int bn) {
  const auto* w = get_w(h, name + ".weight", {out_f, in_f});
  if (!w) return weight_rc(h);
  L->out_f = out_f;
  L->in_f = in_f;
  L->bn = bn;
  int rc;
  if (out_f >= 32) {
    if ((rc = dev_upload(h, &L->w, ffh::to_bf16(*w)))) return rc;
    if ((rc = tmap_2d(h, &L->tmB, L->w, in_f, out_f, 64, bn))) return rc;
  }
  if (h->compute == FF_COMPUTE_FP32 || out_f < 32)
    if ((rc = dev_upload(h, &L->wf, *w))) return rc;
  if (bias) {
    const auto* b = get_w(h, name + ".bias", {out_f});
    if (!b) return weight_rc(h);
    if ((rc = dev_upload(h, &L->b, *b))) return rc;
  }
  return FF_OK;
}

// LayerNorm folded into the linear that follows it (encoder kernel, ff_xf.cuh):
//   LN(x) . W^T + b = rstd * (x . W'^T - mean * c1) + c2,   W' = W * gamma (per input column), c1 = W' . 1, c2 = W . beta + b.
// c1 is summed over the bf16-rounded W' the tensor core multiplies, so the mean term cancels exactly.
int upload_folded_linear(ff_cvit* h, LinearDev* L, float** c1, float** c2, const std::string& name, const std::string& norm,
                         int out_f, int in_f, bool bias) {
  const auto* w = get_w(h, name + ".weight", {out_f, in_f});
  const auto* g = w ? get_w(h, norm + ".weight", {in_f}) : nullptr;
  const auto* be = g ? get_w(h, norm + ".bias", {in_f}) : nullptr;
  const auto* b = (be && bias) ? get_w(h, name + ".bias", {out_f}) : nullptr;
  if (!w || !g || !be || (bias && !b)) return weight_rc(h);
  std::vector<float> wg((size_t)out_f * in_f);
  for (int n = 0; n < out_f; ++n)
    for (int k = 0; k < in_f; ++k) wg[(size_t)n * in_f + k] = (*w)[(size_t)n * in_f + k] * (*g)[k];
  const std::vector<bf16> wq = ffh::to_bf16(wg);
  std::vector<float> v1(out_f), v2(out_f);
  for (int n = 0; n < out_f; ++n) {
    double s1 = 0.0, s2 = bias ? (double)(*b)[n] : 0.0;
    for (int k = 0; k < in_f; ++k) {
      s1 += (double)__bfloat162float(wq[(size_t)n * in_f + k]);
      s2 += (double)(*be)[k] * (double)(*w)[(size_t)n * in_f + k];
    }
    v1[n] = (float)s1;
    v2[n] = (float)s2;
  }
  L->out_f = out_f;
  L->in_f = in_f;
  L->bn = 64;
  int rc;
  if ((rc = dev_upload(h, &L->w, wq))) return rc;
  if ((rc = tmap_2d(h, &L->tmB, L->w, in_f, out_f, 64, 64))) return rc;
  if ((rc = dev_upload(h, c1, v1))) return rc;
  return dev_upload(h, c2, v2);
}

int upload_vec(ff_cvit* h, float** p, const std::string& key, int64_t n) {
  const auto* v = get_w(h, key, {n});
  if (!v) return weight_rc(h);
  return dev_upload(h, p, *v);
}

// Which buffer conv layer li (1..16; layer 0 is conv1) reads, per the ping-pong schedule in forward_pass().
const bf16* conv_input_buffer(const ff_cvit* h, int li, int set) {
  const bf16* A = set ? h->bufA2 : h->bufA;
  const bf16* B = set ? h->bufB2 : h->bufB;
  switch (li) {
    case 1: return A; case 2: return B; case 3: return A; case 4: return B; case 5: return A;
    case 6: return h->P; case 7: return h->Q; case 8: return h->kind == 2 ? h->bufR : h->P;
    case 9: return h->Q; case 10: return h->P; case 11: return h->Q; case 12: return h->P;
    case 13: return h->Q; case 14: return h->P; case 15: return h->Q; case 16: return h->P;
  }
  return nullptr;
}
bf16* conv_output_buffer(const ff_cvit* h, int li, int set) {
  bf16* A = set ? h->bufA2 : h->bufA;
  bf16* B = set ? h->bufB2 : h->bufB;
  switch (li) {
    case 0: return A; case 1: return B; case 2: return A; case 3: return B; case 4: return A;
    case 5: return h->P;
    case 6: return h->Q; case 7: return h->P; case 8: return h->Q;
    case 9: return h->P; case 10: return h->Q; case 11: return h->P; case 12: return h->Q;
    case 13: return h->P; case 14: return h->Q; case 15: return h->P; case 16: return h->feat;
  }
  return nullptr;
}

namespace {

void conv_tile_geometry(int hw, int* bw, int* bh, int* bi) {
  if (hw >= 112) { *bw = 16; *bh = 8; *bi = 1; }
  else if (hw == 56) { *bw = 8; *bh = 8; *bi = 2; }
  else if (hw == 28) { *bw = 4; *bh = 4; *bi = 8; }
  else { *bw = 2; *bh = 2; *bi = 32; }
}

// Kernel per feature layer li (0-based; li = 0 is conv1, fused into c12_kernel on the uint8 path):
//   1..3   ws2conv_kernel   (Cin = 32: pixel-pair GEMM, N = 2*Cout)            ff_ws.cuh
//   4, 5   ws2x_conv_kernel (Cin = 64: pixel-pair GEMM on a CTA pair, N = 128)  ff_ws.cuh
//   6..8   ptcw_conv_kernel (Cout = 128: filter = M operand, 256 pixels = N operand)  ff_ptcw.cuh
//   9..16  ptc2_conv_kernel (Cout >= 256: per-tap implicit GEMM on a CTA pair)   ff_ptc2.cuh
int build_conv_maps(ff_cvit* h) {
  for (int pass = 0; pass < 2; ++pass)
    for (int li = 1; li < (pass == 0 ? 17 : 6); ++li) {
      const ConvPlan& p = kConv[li];
      if (pass == 1) h->conv_alt[li] = h->conv[li];
      ConvLayerDev& L = pass == 0 ? h->conv[li] : h->conv_alt[li];
      const int set = pass;
      const int ncap = li <= 5 ? h->s12_cap : h->cap;
      int rc;
      L.ws2 = p.cin == 32;
      L.ws2x = p.cin == 64 && p.cout == 64;
      if (L.ws2) {
        if ((rc = tmap_4d(h, &L.tmA_ws2, conv_input_buffer(h, li, set), 64, p.hw / 2, p.hw, ncap, 64, 10, 18, 1))) return rc;
        if ((rc = tmap_2d(h, &L.tmW_ws2, L.w2, 384, (uint64_t)2 * p.cout, 64, 2 * p.cout))) return rc;
      } else if (L.ws2x) {
        if ((rc = tmap_4d(h, &L.tmA_ws2x, conv_input_buffer(h, li, set), 64, p.hw / 2, p.hw, 2 * ncap, 64, 10, 18, 1))) return rc;
        if ((rc = tmap_2d(h, &L.tmW_ws2x, L.w2x, 768, 128, 64, 64))) return rc;
      } else {
        L.bn = std::min(p.cout, 256);
        conv_tile_geometry(p.hw, &L.bw, &L.bh, &L.bi);
        L.pair2 = L.bn == 256;
        if (L.pair2 && (rc = tmap_4d(h, &L.tmA, conv_input_buffer(h, li, set), p.cin, p.hw, p.hw, ncap, 64, L.bw, L.bh, L.bi))) return rc;
        if (L.pair2) rc = tmap_2d(h, &L.tmB_half, L.w, (uint64_t)9 * p.cin, p.cout, 64, 128);
        else rc = tmap_2d(h, &L.tmB, L.w, (uint64_t)9 * p.cin, p.cout, 64, L.bn);
        if (rc) return rc;
        if (!L.pair2 && (rc = tmap_4d(h, &L.tmA_row, conv_input_buffer(h, li, set), p.cin, p.hw, p.hw, ncap, 64, L.bw + 2, L.bh, L.bi))) return rc;
      }
    }
  return FF_OK;
}

int xf_setup(ff_cvit* h);

int finalize(ff_cvit* h) {
  int rc;
  if (h->kind == 1) {
    if ((rc = finalize_rvk_features(h))) return rc;
  } else
  // ---- conv stack: fold bias + eval BN into (scale, shift); weights -> [cout][kh][kw][cin]
  for (int li = 0; li < 17; ++li) {
    const ConvPlan& p = kConv[li];
    std::vector<float> wsrc, bias_v, scale(p.cout), shift(p.cout), wr((size_t)p.cout * 9 * p.cin);
    std::string bn_key;
    if (h->kind == 2) {
      // plan entry of layer li: the extra BN-less conv (entry 8) sits between layers 8 and 9 (0-based li 7 and 8)
      const GgcaPlan& gp = kGgcaPlan[li < 8 ? li : li + 1];
      if ((rc = ggca_conv_weights(h, gp, p.cin, p.cout, &wsrc, &bias_v))) return rc;
      if (gp.bn_idx >= 0) bn_key = std::string(gp.seq) + "." + std::to_string(gp.bn_idx);
    } else {
      const std::string c = "features." + std::to_string(p.conv_idx);
      bn_key = "features." + std::to_string(p.conv_idx + 1);
      const auto* w = get_w(h, c + ".weight", {p.cout, p.cin, 3, 3});
      const auto* bias = w ? get_w(h, c + ".bias", {p.cout}) : nullptr;
      if (!w || !bias) return weight_rc(h);
      wsrc = *w;
      bias_v = *bias;
    }
    if (!bn_key.empty()) {
      const auto* g = get_w(h, bn_key + ".weight", {p.cout});
      const auto* be = g ? get_w(h, bn_key + ".bias", {p.cout}) : nullptr;
      const auto* mu = be ? get_w(h, bn_key + ".running_mean", {p.cout}) : nullptr;
      const auto* var = mu ? get_w(h, bn_key + ".running_var", {p.cout}) : nullptr;
      if (!g || !be || !mu || !var) return weight_rc(h);
      for (int o = 0; o < p.cout; ++o) {
        const float s = (*g)[o] / std::sqrt((*var)[o] + BN_EPS);
        scale[o] = s;
        shift[o] = (bias_v[o] - (*mu)[o]) * s + (*be)[o];
      }
    } else {
      for (int o = 0; o < p.cout; ++o) { scale[o] = 1.0f; shift[o] = bias_v[o]; }     // BN-less DEConv (features1.27)
    }
    for (int o = 0; o < p.cout; ++o)
      for (int ci = 0; ci < p.cin; ++ci)
        for (int t = 0; t < 9; ++t) wr[((size_t)o * 9 + t) * p.cin + ci] = wsrc[((size_t)o * p.cin + ci) * 9 + t];
    ConvLayerDev& L = h->conv[li];
    if (p.cout <= 64)
      for (int o = 0; o < p.cout; ++o) { L.epi.scale[o] = scale[o]; L.epi.shift[o] = shift[o]; }
    if ((rc = dev_upload(h, &L.scale, scale))) return rc;
    if ((rc = dev_upload(h, &L.shift, shift))) return rc;
    if (li == 0) {
      std::vector<float> w32(32 * 64, 0.0f);     // fp32-input kernel: k = kh*16 + kw*4 + cin (ff_c1.cuh)
      for (int o = 0; o < 32; ++o) {
        for (int k = 0; k < 27; ++k) {
          const int tap = k / 3, c = k % 3, kh = tap / 3, kw = tap % 3;
          w32[o * 64 + kh * 16 + kw * 4 + c] = wr[(size_t)o * 27 + k];
        }
        h->c1_scale[o] = scale[o];
        h->c1_shift[o] = shift[o];
      }
      if (h->compute == FF_COMPUTE_BF16) {
        if ((rc = dev_upload(h, &h->c1_w, to_act16(h, w32)))) return rc;
        // pair-expanded filter: B[kh][(p,co)][(q,c)] = W[co][kh][q-p][c], zero unless 0 <= q-p <= 2 and c < 3
        std::vector<float> wp(3 * 64 * 16, 0.0f);
        for (int kh = 0; kh < 3; ++kh)
          for (int pp = 0; pp < 2; ++pp)
            for (int o = 0; o < 32; ++o)
              for (int q = 0; q < 4; ++q) {
                const int kw = q - pp;
                if (kw < 0 || kw > 2) continue;
                for (int c = 0; c < 3; ++c) wp[(kh * 64 + pp * 32 + o) * 16 + q * 4 + c] = wr[(size_t)o * 27 + (kh * 3 + kw) * 3 + c];
              }
        if ((rc = dev_upload(h, &h->c1_wp, to_act16(h, wp)))) return rc;
        // The uint8 kernels normalise with ONE fma per channel, u*na + nb.  It must reproduce, bit for bit after the
        // rounding to the 16-bit operand type, the fp32 arithmetic of cvit_prediction.py:41-45,214-215 for all 768 codes.
        const float mean[3] = {0.485f, 0.456f, 0.406f}, sd[3] = {0.229f, 0.224f, 0.225f};
        for (int c = 0; c < 3; ++c) {
          h->c1_na[c] = 1.0f / (255.0f * sd[c]);
          h->c1_nb[c] = -mean[c] / sd[c];
          for (int u = 0; u < 256; ++u) {
            const std::vector<float> two = {std::fmaf((float)u, h->c1_na[c], h->c1_nb[c]), ((float)u / 255.0f - mean[c]) / sd[c]};
            const std::vector<bf16> r = to_act16(h, two);
            unsigned short b0, b1;
            memcpy(&b0, &r[0], 2);
            memcpy(&b1, &r[1], 2);
            // bf16: bit-identical for all 768 codes.  fp16 (11-bit mantissa): the two roundings of values 1e-7 apart may
            // land on neighbouring codes; accept one unit in the last place (2^-11 relative, below the type's own rounding)
            const int ulp = b0 > b1 ? b0 - b1 : b1 - b0;
            if (ulp > (h->act_f16 ? 1 : 0))
              return fail(h, FF_ERR_STATE, "internal: single-FMA normalisation differs from (u/255-mean)/std at code %d channel %d", u, c);
          }
        }
      }
    }
    if (h->compute == FF_COMPUTE_FP32) {
      if ((rc = dev_upload(h, &L.wf, wr))) return rc;
    } else if (li > 0) {
      if (p.cin == 32) {
        // pair-expanded filter: B[(p,co)][(kh,q,ci)] = W[co][kh][q-p][ci], zero unless 0 <= q-p <= 2 (ff_ws.cuh)
        std::vector<float> w2((size_t)2 * p.cout * 384, 0.0f);
        for (int pp = 0; pp < 2; ++pp)
          for (int o = 0; o < p.cout; ++o)
            for (int kh = 0; kh < 3; ++kh)
              for (int q = 0; q < 4; ++q) {
                const int kw = q - pp;
                if (kw < 0 || kw > 2) continue;
                for (int ci = 0; ci < 32; ++ci)
                  w2[((size_t)pp * p.cout + o) * 384 + kh * 128 + q * 32 + ci] = wr[((size_t)o * 9 + kh * 3 + kw) * 32 + ci];
              }
        if ((rc = dev_upload(h, &L.w2, to_act16(h, w2)))) return rc;
      } else if (p.cin == 64 && p.cout == 64) {
        // CTA-pair variant: B[(pp,co)][(cb,kh,q,ci)] = W[co][kh][q-pp][cb*32+ci]
        std::vector<float> w2((size_t)128 * 768, 0.0f);
        for (int pp = 0; pp < 2; ++pp)
          for (int o = 0; o < 64; ++o)
            for (int cb = 0; cb < 2; ++cb)
              for (int kh = 0; kh < 3; ++kh)
                for (int q = 0; q < 4; ++q) {
                  const int kw = q - pp;
                  if (kw < 0 || kw > 2) continue;
                  for (int ci = 0; ci < 32; ++ci)
                    w2[((size_t)pp * 64 + o) * 768 + cb * 384 + kh * 128 + q * 32 + ci] = wr[((size_t)o * 9 + kh * 3 + kw) * 64 + cb * 32 + ci];
                }
        if ((rc = dev_upload(h, &L.w2x, to_act16(h, w2)))) return rc;
      } else {
        if ((rc = dev_upload(h, &L.w, to_act16(h, wr)))) return rc;
      }
    }
  }
  // ---- embedding / tokens
  if ((rc = upload_linear(h, &h->embed, "patch_to_embedding", DIM, PATCH, true, 128))) return rc;
  {
    const auto* pos = get_w(h, "pos_embedding", {SLOTS, 1, DIM});
    const auto* cls = pos ? get_w(h, "cls_token", {1, 1, DIM}) : nullptr;
    if (!pos || !cls) return weight_rc(h);
    if ((rc = dev_upload(h, &h->pos, *pos))) return rc;
    if ((rc = dev_upload(h, &h->cls, *cls))) return rc;
  }
  // ---- transformer
  for (int l = 0; l < DEPTH; ++l) {
    const std::string p = "transformer.layers." + std::to_string(l);
    XfLayerDev& X = h->xf[l];
    if ((rc = upload_vec(h, &X.ln1_g, p + ".0.fn.norm.weight", DIM))) return rc;
    if ((rc = upload_vec(h, &X.ln1_b, p + ".0.fn.norm.bias", DIM))) return rc;
    // kind 2: the MLP branch is pre-normed by LinearNorm, which in eval() is its norm1 = LayerNorm(eps 1e-6)
    // (cvit_GGCA_ADD_DEConv_RepBn8.py:22-60); its RepBN / schedule buffers are not on the inference path
    const std::string ln2 = h->kind == 2 ? p + ".1.fn.norm.norm1" : p + ".1.fn.norm";
    if ((rc = upload_vec(h, &X.ln2_g, ln2 + ".weight", DIM))) return rc;
    if ((rc = upload_vec(h, &X.ln2_b, ln2 + ".bias", DIM))) return rc;
    if ((rc = upload_linear(h, &X.qkv, p + ".0.fn.fn.to_qkv", 3 * DIM, DIM, false, 64))) return rc;
    if ((rc = upload_linear(h, &X.out, p + ".0.fn.fn.to_out", DIM, DIM, true, 64))) return rc;
    if ((rc = upload_linear(h, &X.ff1, p + ".1.fn.fn.net.0", MLP, DIM, true, 64))) return rc;
    if ((rc = upload_linear(h, &X.ff2, p + ".1.fn.fn.net.2", DIM, MLP, true, 64))) return rc;
    if (h->compute == FF_COMPUTE_BF16) {
      if ((rc = upload_folded_linear(h, &X.qkv_f, &X.c1q, &X.c2q, p + ".0.fn.fn.to_qkv", p + ".0.fn.norm", 3 * DIM, DIM, false))) return rc;
      if ((rc = upload_folded_linear(h, &X.ff1_f, &X.c1f, &X.c2f, p + ".1.fn.fn.net.0", ln2, MLP, DIM, true))) return rc;
    }
  }
  if (h->kind == 1) {     // kan_head = Linear, Dropout, ReLU, KAN (ResVitKan.py:302-307); mlp_head is not on the forward path
    if ((rc = upload_linear(h, &h->head1, "kan_head.0", MLP, DIM, true, 64))) return rc;
  } else {
    if ((rc = upload_linear(h, &h->head1, "mlp_head.0", MLP, DIM, true, 64))) return rc;
    if ((rc = upload_linear(h, &h->head2, "mlp_head.2", 2, MLP, true, 64))) return rc;
  }

  if (h->kind == 2 && (rc = finalize_ggca_extras(h))) return rc;
  if (h->compute == FF_COMPUTE_BF16) {
    if (h->kind != 1 && (rc = build_conv_maps(h))) return rc;
    const int cap128 = (h->cap + 127) / 128 * 128;
    if ((rc = tmap_2d(h, &h->tm_feat, h->feat, PATCH, cap128, 64, 128))) return rc;
    if ((rc = tmap_2d(h, &h->tm_xn, h->xn, DIM, h->rows_cap, 64, 128))) return rc;
    if ((rc = tmap_2d(h, &h->tm_att, h->att, DIM, h->rows_cap, 64, 128))) return rc;
    if ((rc = tmap_2d(h, &h->tm_ffh, h->ffh_buf, MLP, h->rows_cap, 64, 128))) return rc;
    if ((rc = tmap_2d(h, &h->tm_cls, h->clsb, DIM, cap128, 64, 128))) return rc;
    if ((rc = xf_setup(h))) return rc;
  }
  h->unused_keys.clear();
  for (const auto& kv : h->host_w)
    if (!h->host_used.count(kv.first)) h->unused_keys += (h->unused_keys.empty() ? "" : ",") + kv.first;
  h->host_w.clear();
  h->host_shape.clear();
  h->host_used.clear();
  h->finalized = true;
  return FF_OK;
}

// ------------------------------------------------------------------------------------------------ encoder kernel
// One cooperative launch for the 6 transformer layers (ff_xf.cuh): groups of 16 CTAs, one group per 128-row token tile
// (groups loop over tiles when there are more tiles than co-resident groups).
int xf_setup(ff_cvit* h) {
  h->xf_ready = false;
  int coop = 0;
  FF_CUDA(h, cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, h->device));
  if (!coop) return FF_OK;
  std::vector<CUtensorMap> maps(3 + 4 * DEPTH);
  maps[0] = h->tm_xn;
  maps[1] = h->tm_att;
  maps[2] = h->tm_ffh;
  for (int l = 0; l < DEPTH; ++l) {
    maps[3 + 4 * l + 0] = h->xf[l].qkv_f.tmB;     // LayerNorm-folded weights (upload_folded_linear)
    maps[3 + 4 * l + 1] = h->xf[l].out.tmB;
    maps[3 + 4 * l + 2] = h->xf[l].ff1_f.tmB;
    maps[3 + 4 * l + 3] = h->xf[l].ff2.tmB;
  }
  h->xf_maps = maps;
  int rc;
  if ((rc = dev_alloc(h, &h->xf_sync, (size_t)2 * XF_MAX_GROUPS))) return rc;
  FF_CUDA(h, cudaMemset(h->xf_sync, 0, 2 * XF_MAX_GROUPS * sizeof(unsigned int)));   // the kernel re-arms them itself
  if ((rc = dev_alloc(h, &h->xf_stats, (size_t)2 * h->rows_cap * 64))) return rc;
  FF_CUDA(h, cudaMemset(h->xf_stats, 0, (size_t)2 * h->rows_cap * 64 * sizeof(float)));
  FF_CUDA(h, ffh::ensure_dyn_smem(reinterpret_cast<const void*>(xf_kernel), XF_SMEM_TOTAL));
  int per_sm = 0;
  FF_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, xf_kernel, XF_THREADS, XF_SMEM_TOTAL));
  h->xf_groups = std::min(per_sm * h->num_sms / XF_CS, (int)XF_MAX_GROUPS);
  h->xf_ready = h->xf_groups >= 1;
  return FF_OK;
}

// returns FF_OK, a negative error, or 1 = "cooperative launch refused, run the per-op launches"
int launch_xf(ff_cvit* h, cudaStream_t st, int n, int depth) {
  XfArgs a;
  memset(&a, 0, sizeof(a));
  a.x = h->x; a.xn = h->xn; a.qkv = h->qkvb; a.att = h->att; a.ffh = h->ffh_buf;
  static_assert(sizeof(XfArgs) <= 4096, "XfArgs must fit the kernel parameter space");
  memcpy(a.maps, h->xf_maps.data(), sizeof(a.maps));
  a.sync = h->xf_sync;
  a.stats = reinterpret_cast<float2*>(h->xf_stats);
  a.stats_stride = (long long)h->rows_cap * 32;
  a.rows = 2 * n; a.n_crops = n; a.depth = depth;
  a.eps1 = 1e-5f; a.eps2 = h->ln2_eps;
  for (int l = 0; l < DEPTH; ++l) {
    const XfLayerDev& X = h->xf[l];
    a.L[l] = XfLayerP{X.c1q, X.c2q, X.c1f, X.c2f, X.out.b, X.ff2.b};
  }
  const int tiles = (a.rows + 127) / 128;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(XF_CS * std::min(tiles, h->xf_groups));
  cfg.blockDim = dim3(XF_THREADS);
  cfg.dynamicSmemBytes = XF_SMEM_TOTAL;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;    // all CTAs co-resident: the group barriers spin on global counters
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  FF_CUDA(h, ffh::ensure_dyn_smem(reinterpret_cast<const void*>(xf_kernel), XF_SMEM_TOTAL));
  ProfScope ps(h, st, KC_GEMM_XF);
  cudaError_t e = cudaLaunchKernelEx(&cfg, xf_kernel, a);
  if (e == cudaErrorCooperativeLaunchTooLarge || e == cudaErrorNotSupported) {
    // the device cannot hold the groups any more (e.g. SMs reserved by another client): the per-op GPU launches take over
    cudaGetLastError();
    h->xf_ready = false;
    return 1;
  }
  if (e != cudaSuccess) return fail(h, FF_ERR_CUDA, "launch of the encoder kernel failed: %s", cudaGetErrorString(e));
  ++h->launches;
  return FF_OK;
}

// ------------------------------------------------------------------------------------------------ forward
int launch_gemm(ff_cvit* h, cudaStream_t st, const CUtensorMap& tmA, const LinearDev& L, int M, void* out, int ldo, int epi,
                int act, int splits, const char* what) {
  TcArgs a;
  memset(&a, 0, sizeof(a));
  a.M = M;
  a.N = L.out_f;
  a.ldo = ldo;
  a.kb_total = L.in_f / 64;
  a.kb_per_split = (a.kb_total + splits - 1) / splits;
  a.shift = (splits > 1) ? nullptr : L.b;     // split-K: bias is added by the consumer of the partial slabs
  a.split_stride = (splits > 1) ? (long long)h->cap * ldo : 0;
  a.out = out;
  a.epi = epi;
  a.act = act;
  const int zs = (a.kb_total + a.kb_per_split - 1) / a.kb_per_split;
  dim3 grid((M + 127) / 128, L.out_f / L.bn, zs);
  ProfScope ps(h, st, &L == &h->embed ? KC_GEMM_EMBED : (&L == &h->head1 ? KC_GEMM_HEAD : KC_GEMM_XF));
  cudaError_t e = L.bn == 128 ? launch_tc_gemm<128>(grid, st, tmA, L.tmB, a) : launch_tc_gemm<64>(grid, st, tmA, L.tmB, a);
  if (e != cudaSuccess) return fail(h, FF_ERR_CUDA, "launch of gemm %s failed: %s", what, cudaGetErrorString(e));
  ++h->launches;
  return FF_OK;
}

int forward_fp32(ff_cvit* h, const void* x, int layout, const int32_t* slot, int slot_base, int n, float* logits,
                 cudaStream_t st, DebugTap* tap);

int encode_u8_map(ff_cvit* h, CUtensorMap* tmX, const uint8_t* xin, int ns) {
  // per-launch 3-D map over the caller's uint8 crops viewed as [ns][224][672]; TMA needs 16-byte aligned coordinates
  cuuint64_t dims[3] = {672, 224, (cuuint64_t)ns};
  cuuint64_t strides[2] = {672, (cuuint64_t)224 * 672};
  cuuint32_t box[3] = {80, 18, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = ffh::encode_tiled()(tmX, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<uint8_t*>(xin), dims, strides, box, estr,
                                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(h, FF_ERR_CUDA, "cuTensorMapEncodeTiled(uint8 crops) failed: %d (is the input 16-byte aligned?)", (int)r);
  return FF_OK;
}

// One feature layer li >= 1 on n_img images of ping-pong set `set`.
int run_conv(ff_cvit* h, int li, int n_img, int img_off_out, int set, cudaStream_t st) {
  const ConvPlan& p = kConv[li];
  const ConvLayerDev& L = (set && li <= 5) ? h->conv_alt[li] : h->conv[li];
  TcArgs a;
  memset(&a, 0, sizeof(a));
  a.H = p.hw; a.W = p.hw;
  a.n_img = n_img;
  a.img_off_out = img_off_out;
  a.cout = p.cout;
  a.cin = p.cin;
  a.scale = L.scale; a.shift = L.shift;
  a.out = conv_output_buffer(h, li, set);
  ProfScope ps(h, st, KC_TC_CONV + li - 1);
  cudaError_t e;
  const bool f16 = h->act_f16;     // GGCA variant: fp16 activations / filters (same kernels, other operand-format template)
  if (L.ws2x) {
    a.tiles_w = p.hw / 16; a.tiles_h = p.hw / 16;
    a.out_blocked = (li == 4) ? 1 : 0;          // layer 5 feeds layer 6 (also a pair kernel); layer 6 writes plain NHWC
    const int tiles = a.tiles_w * a.tiles_h * n_img;
    const int g = std::min(2 * ((tiles + 1) / 2), h->num_sms & ~1);
    const WsEpi& epi = reinterpret_cast<const WsEpi&>(L.epi);
    if (f16) e = p.pool ? launch_ws2x<true, true>(g, st, L.tmA_ws2x, L.tmW_ws2x, a, epi) : launch_ws2x<false, true>(g, st, L.tmA_ws2x, L.tmW_ws2x, a, epi);
    else e = p.pool ? launch_ws2x<true>(g, st, L.tmA_ws2x, L.tmW_ws2x, a, epi) : launch_ws2x<false>(g, st, L.tmA_ws2x, L.tmW_ws2x, a, epi);
  } else if (L.ws2) {
    a.out_blocked = (li == 3) ? 1 : 0;          // layer 4 feeds the CTA-pair kernel of layer 5
    a.tiles_w = p.hw / 16; a.tiles_h = p.hw / 16;
    const int tiles = a.tiles_w * a.tiles_h * n_img;
    const WsEpi& epi = reinterpret_cast<const WsEpi&>(L.epi);
    if (p.cout == 32) {
      const int g = std::min(tiles, h->num_sms * 2);
      if (f16) e = p.pool ? launch_ws2<64, true, 2, true>(g, st, L.tmA_ws2, L.tmW_ws2, a, epi) : launch_ws2<64, false, 2, true>(g, st, L.tmA_ws2, L.tmW_ws2, a, epi);
      else e = p.pool ? launch_ws2<64, true, 2>(g, st, L.tmA_ws2, L.tmW_ws2, a, epi) : launch_ws2<64, false, 2>(g, st, L.tmA_ws2, L.tmW_ws2, a, epi);
    } else {
      e = f16 ? launch_ws2<128, false, 4, true>(std::min(tiles, h->num_sms), st, L.tmA_ws2, L.tmW_ws2, a, epi)
              : launch_ws2<128, false, 4>(std::min(tiles, h->num_sms), st, L.tmA_ws2, L.tmW_ws2, a, epi);
    }
  } else {
    a.tiles_w = p.hw / L.bw; a.tiles_h = p.hw / L.bh;
    a.lg_bw = ffh::ilog2(L.bw); a.lg_bh = ffh::ilog2(L.bh);
    a.kb_per_tap = p.cin / 64;
    a.kb_total = 9 * a.kb_per_tap;
    a.kb_per_split = a.kb_total;
    const int m_tiles = a.tiles_w * a.tiles_h * ((n_img + L.bi - 1) / L.bi);
    if (L.pair2) {         // one item = two pixel tiles x one 256-channel tile on a CTA pair
      const int items = ((m_tiles + 1) / 2) * (p.cout / 256);
      const int g2 = std::min(2 * items, h->num_sms & ~1);
      if (f16) e = p.pool ? launch_ptc2<true, true>(g2, st, L.tmA, L.tmB_half, a) : launch_ptc2<false, true>(g2, st, L.tmA, L.tmB_half, a);
      else e = p.pool ? launch_ptc2<true>(g2, st, L.tmA, L.tmB_half, a) : launch_ptc2<false>(g2, st, L.tmA, L.tmB_half, a);
    } else {               // Cout = 128 at 56 x 56: the filter tile is the M operand, two pixel sub-tiles (256 pixels) the N operand
      const int tiles = ((m_tiles + 1) / 2) * (p.cout / 128);
      const int g = std::min(tiles, h->num_sms);
      if (f16) e = p.pool ? launch_ptcw<true, true>(g, st, L.tmA_row, L.tmB, a) : launch_ptcw<false, true>(g, st, L.tmA_row, L.tmB, a);
      else e = p.pool ? launch_ptcw<true>(g, st, L.tmA_row, L.tmB, a) : launch_ptcw<false>(g, st, L.tmA_row, L.tmB, a);
    }
  }
  if (e != cudaSuccess) return fail(h, FF_ERR_CUDA, "launch of conv layer %d failed: %s", li + 1, cudaGetErrorString(e));
  ++h->launches;
  return FF_OK;
}

// One pass over n <= cap crops.  x points at the first crop of the pass.
int forward_pass(ff_cvit* h, const void* x, int layout, const int32_t* slot, int slot_base, int n, float* logits,
                 cudaStream_t st, DebugTap* tap) {
  if (h->compute == FF_COMPUTE_FP32) return forward_fp32(h, x, layout, slot, slot_base, n, logits, st, tap);
  const int stop = tap ? tap->stop_after : 0;
  auto tap_hit = [&](int step, const void* p, int64_t elems, bool is_16) {
    if (stop == step) { tap->ptr = p; tap->elems = elems; tap->is_16 = is_16; tap->hit = true; return true; }
    return false;
  };
  const size_t crop_in_bytes = layout == FF_X_NHWC_U8 ? (size_t)224 * 224 * 3 : (size_t)224 * 224 * 3 * 4;
  int rc;

  if (h->kind == 1) {
    rc = rvk_features(h, x, layout, slot_base, n, st, tap);
    if (rc || (tap && tap->hit)) return rc;
  } else {
    // ---- stages 1-2 in sub-passes of s12 crops
    prof_mark(h, st, 0, true, true);
    const int sub = stop ? std::min(n, h->s12_cap) : h->s12;
    if (stop && n > h->s12_cap) return fail(h, FF_ERR_BAD_ARG, "debug tap needs n <= %d", h->s12_cap);
    // consecutive sub-passes alternate between the caller's stream and aux_stream (own ping-pong buffers), so the
    // launch tail / prologue of one chain is filled by the other chain's kernels
    const bool dual = !stop && (!h->profiling || h->prof_coarse) && n > sub;
    cudaStream_t st_main = st;
    if (dual) {
      FF_CUDA(h, cudaEventRecord(h->ev_fork, st_main));
      FF_CUDA(h, cudaStreamWaitEvent(h->aux_stream, h->ev_fork, 0));
    }
    int sub_idx = 0;
    for (int s0 = 0; s0 < n; s0 += sub, ++sub_idx) {
      const int set = dual ? (sub_idx & 1) : 0;
      cudaStream_t sst = set ? h->aux_stream : st_main;
      bf16* bufA = set ? h->bufA2 : h->bufA;
      const int ns = std::min(sub, n - s0);
      const uint8_t* xin = reinterpret_cast<const uint8_t*>(x) + (size_t)s0 * crop_in_bytes;
      if (h->h2d_chunks_pending > 0) {   // input still streaming in: wait for the chunks covering [g0, g0+ns)
        const int g0 = slot_base + s0;   // slot_base == offset of this pass inside the whole batch
        const int c1 = std::min(h->h2d_chunks_pending - 1, (g0 + ns - 1) / h->h2d_chunk);
        FF_CUDA(h, cudaStreamWaitEvent(sst, h->h2d_ready[c1], 0));
      }
      // layers 1 + 2 in one kernel on the uint8 path unless layer 1's own output is asked for
      const bool fused12 = layout == FF_X_NHWC_U8 && stop != 1;
      if (fused12) {
        ProfScope ps(h, sst, KC_TC_CONV);
        CUtensorMap tmX;
        if ((rc = encode_u8_map(h, &tmX, xin, ns))) return rc;
        C12Args ca;
        ca.out = conv_output_buffer(h, 1, set); ca.w1 = h->c1_wp; ca.n_img = ns;
        for (int c = 0; c < 3; ++c) { ca.na[c] = h->c1_na[c]; ca.nb[c] = h->c1_nb[c]; }
        for (int o = 0; o < 32; ++o) {
          ca.scale1[o] = h->c1_scale[o]; ca.shift1[o] = h->c1_shift[o];
          ca.scale2[o] = h->conv[1].epi.scale[o]; ca.shift2[o] = h->conv[1].epi.shift[o];
        }
        const ConvLayerDev& L2 = set ? h->conv_alt[1] : h->conv[1];
        const int grid = std::min(16 * 14 * ns, h->num_sms * 2);
        cudaError_t e = h->act_f16 ? launch_c12<true>(grid, sst, tmX, L2.tmW_ws2, ca) : launch_c12<false>(grid, sst, tmX, L2.tmW_ws2, ca);
        if (e != cudaSuccess) return fail(h, FF_ERR_CUDA, "launch of the fused layer-1/2 kernel failed: %s", cudaGetErrorString(e));
        ++h->launches;
      } else {
        ProfScope ps(h, sst, KC_CONV1);
        cudaError_t e;
        if (layout == FF_X_NHWC_U8) {
          CUtensorMap tmX;
          if ((rc = encode_u8_map(h, &tmX, xin, ns))) return rc;
          C1PairArgs ca;
          ca.out = bufA; ca.w = h->c1_wp; ca.n_img = ns;
          for (int c = 0; c < 3; ++c) { ca.na[c] = h->c1_na[c]; ca.nb[c] = h->c1_nb[c]; }
          for (int o = 0; o < 32; ++o) { ca.scale[o] = h->c1_scale[o]; ca.shift[o] = h->c1_shift[o]; }
          e = h->act_f16 ? ffh::launch_k(conv1_pair_kernel<true>, dim3(std::min(196 * ns, h->num_sms * 8)), dim3(128), 0, sst, true, tmX, ca)
                         : ffh::launch_k(conv1_pair_kernel<false>, dim3(std::min(196 * ns, h->num_sms * 8)), dim3(128), 0, sst, true, tmX, ca);
        } else {
          C1Args ca;
          ca.x = reinterpret_cast<const float*>(xin); ca.out = bufA; ca.w = h->c1_w; ca.n_img = ns;
          for (int o = 0; o < 32; ++o) { ca.scale[o] = h->c1_scale[o]; ca.shift[o] = h->c1_shift[o]; }
          e = h->act_f16 ? ffh::launch_k(conv1_f32_kernel<true>, dim3(std::min(392 * ns, h->num_sms * 8)), dim3(128), 0, sst, true, ca)
                         : ffh::launch_k(conv1_f32_kernel<false>, dim3(std::min(392 * ns, h->num_sms * 8)), dim3(128), 0, sst, true, ca);
        }
        if (e != cudaSuccess) return fail(h, FF_ERR_CUDA, "launch of conv1 failed: %s", cudaGetErrorString(e));
        ++h->launches;
      }
      if (!fused12 && tap_hit(1, bufA, (int64_t)ns * 224 * 224 * 32, true)) return FF_OK;
      for (int li = 1; li <= 5; ++li) {
        if (!(fused12 && li == 1) && (rc = run_conv(h, li, ns, li == 5 ? s0 : 0, set, sst))) return rc;
        const ConvPlan& p = kConv[li];
        const int ohw = p.pool ? p.hw / 2 : p.hw;
        if (tap_hit(li + 1, conv_output_buffer(h, li, set), (int64_t)ns * ohw * ohw * p.cout, true)) {
          if (li == 3 || li == 4) tap->blocked_hw = ohw;
          return FF_OK;
        }
      }
    }
    if (dual) {
      FF_CUDA(h, cudaEventRecord(h->ev_join, h->aux_stream));
      FF_CUDA(h, cudaStreamWaitEvent(st_main, h->ev_join, 0));
    }
    prof_mark(h, st, 0, false, true);
    prof_mark(h, st, 1, true, true);
    // ---- stages 3-5 on the whole pass
    for (int li = 6; li < 17; ++li) {
      if ((rc = run_conv(h, li, n, 0, 0, st))) return rc;
      const ConvPlan& p = kConv[li];
      const int ohw = p.pool ? p.hw / 2 : p.hw;
      if (tap_hit(li + 1, conv_output_buffer(h, li), (int64_t)n * ohw * ohw * p.cout, true)) return FF_OK;
      if (h->kind == 2 && li == 7) {   // features1.26: Conv2d(128,128), no BN, no activation -> bufR (input of layer 9)
        if ((rc = rvk_launch_op(h, h->rvk_ops[0], n, st, KC_SMALL))) return rc;
        if (tap_hit(26, h->bufR, (int64_t)n * 56 * 56 * 128, true)) return FF_OK;
      }
    }
    if (h->kind == 2) {                // x = x * ggca(x)  (cvit_GGCA_ADD_DEConv_RepBn8.py:447-448), in place on feat
      if ((rc = ggca_gate(h, n, st))) return rc;
      if (tap_hit(27, h->feat, (int64_t)n * PATCH, true)) return FF_OK;
    }
    prof_mark(h, st, 1, false, true);
  }
  prof_mark(h, st, 2, true, true);
  // ---- patch embedding (split-K, deterministic slabs) + token assembly
  rc = launch_gemm(h, st, h->tm_feat, h->embed, n, h->emb, DIM, EPI_STORE_F32, ACT_NONE, EMBED_SPLITS, "patch_to_embedding");
  if (rc) return rc;
  { ProfScope ps(h, st, KC_SMALL); ffh::launch_k(tokens_kernel, dim3(n), dim3(256), 0, st, true, (const float*)h->emb, (int)EMBED_SPLITS, (long long)h->cap * DIM, (const float*)h->embed.b, (const float*)h->cls, (const float*)h->pos, slot, slot_base, h->x, n,
                                                   h->xf_ready ? h->xn : (bf16*)nullptr, reinterpret_cast<float2*>(h->xf_stats),
                                                   reinterpret_cast<float2*>(h->xf_stats) + (size_t)h->rows_cap * 32); }
  FF_LAUNCH_CHECK(h, "tokens");
  const int rows = 2 * n;
  if (tap_hit(18, h->x, (int64_t)rows * DIM, false)) return FF_OK;
  // ---- transformer: one cooperative launch for all layers (a debug tap inside the encoder shortens the depth)
  bool encoder_done = false;
  if (h->xf_ready) {
    const int depth = (stop >= 19 && stop < 19 + DEPTH) ? stop - 18 : DEPTH;
    rc = launch_xf(h, st, n, depth);
    if (rc < 0) return rc;
    if (rc == 0) {
      if (tap_hit(18 + depth, h->x, (int64_t)rows * DIM, false)) return FF_OK;
      encoder_done = true;
    }
  }
  // per-op launches: only when the cooperative launch is not available on this device / was refused at run time
  for (int l = 0; l < DEPTH && !encoder_done; ++l) {
    const XfLayerDev& X = h->xf[l];
    { ProfScope ps(h, st, KC_SMALL); ffh::launch_k(layernorm_kernel, dim3((rows + 7) / 8), dim3(256), 0, st, true, (const float*)h->x, (const float*)X.ln1_g, (const float*)X.ln1_b, h->xn, rows, 1e-5f); }
    FF_LAUNCH_CHECK(h, "layernorm1");
    if ((rc = launch_gemm(h, st, h->tm_xn, X.qkv, rows, h->qkvb, 3 * DIM, EPI_STORE_BF16, ACT_NONE, 1, "to_qkv"))) return rc;
    { ProfScope ps(h, st, KC_SMALL); ffh::launch_k(attention2_kernel, dim3((n * 8 + 7) / 8), dim3(256), 0, st, true, (const bf16*)h->qkvb, h->att, n); }
    FF_LAUNCH_CHECK(h, "attention2");
    if ((rc = launch_gemm(h, st, h->tm_att, X.out, rows, h->x, DIM, EPI_RESID_F32, ACT_NONE, 1, "to_out"))) return rc;
    { ProfScope ps(h, st, KC_SMALL); ffh::launch_k(layernorm_kernel, dim3((rows + 7) / 8), dim3(256), 0, st, true, (const float*)h->x, (const float*)X.ln2_g, (const float*)X.ln2_b, h->xn, rows, h->ln2_eps); }
    FF_LAUNCH_CHECK(h, "layernorm2");
    if ((rc = launch_gemm(h, st, h->tm_xn, X.ff1, rows, h->ffh_buf, MLP, EPI_STORE_BF16, ACT_GELU, 1, "ff1"))) return rc;
    if ((rc = launch_gemm(h, st, h->tm_ffh, X.ff2, rows, h->x, DIM, EPI_RESID_F32, ACT_NONE, 1, "ff2"))) return rc;
    if (tap_hit(19 + l, h->x, (int64_t)rows * DIM, false)) return FF_OK;
  }
  // ---- head
  { ProfScope ps(h, st, KC_SMALL); ffh::launch_k(cls_gather_kernel, dim3(n), dim3(256), 0, st, true, (const float*)h->x, h->clsb, n); }
  FF_LAUNCH_CHECK(h, "cls_gather");
  if ((rc = launch_gemm(h, st, h->tm_cls, h->head1, n, h->hid, MLP, EPI_STORE_F32, ACT_RELU, 1, "head.0"))) return rc;
  if (h->kind == 1) {
    if ((rc = kan_head(h, n, logits, st))) return rc;
  } else {
    { ProfScope ps(h, st, KC_SMALL); ffh::launch_k(head2_kernel, dim3((n + 7) / 8), dim3(256), 0, st, true, (const float*)h->hid, (const float*)h->head2.wf, (const float*)h->head2.b, logits, n); }
    FF_LAUNCH_CHECK(h, "head2");
  }
  prof_mark(h, st, 2, false, true);
  if (tap_hit(25, logits, (int64_t)n * 2, false)) return FF_OK;
  return FF_OK;
}

// ---- fp32 CUDA-core path (parity to 1e-4; reuses LN/attention-like kernels in fp32)
int forward_fp32(ff_cvit* h, const void* x, int layout, const int32_t* slot, int slot_base, int n, float* logits,
                 cudaStream_t st, DebugTap* tap) {
  const int stop = tap ? tap->stop_after : 0;
  auto tap_hit = [&](int step, const void* p, int64_t elems) {
    if (stop == step) { tap->ptr = p; tap->elems = elems; tap->is_16 = false; tap->hit = true; return true; }
    return false;
  };
  if (h->h2d_chunks_pending > 0) FF_CUDA(h, cudaStreamWaitEvent(st, h->h2d_ready[h->h2d_chunks_pending - 1], 0));
  // conv stack, NHWC fp32, ping-pong fA/fB; processed `chunk` crops at a time to bound the workspace
  const int chunk = h->s12_cap;
  const size_t crop_in_bytes = layout == FF_X_NHWC_U8 ? (size_t)224 * 224 * 3 : (size_t)224 * 224 * 3 * 4;
  float* featf = h->featf;                                  // tail of fB is reserved for [cap][25088]
  if (stop && n > chunk) return fail(h, FF_ERR_BAD_ARG, "debug tap needs n <= %d", chunk);
  if (stop && h->kind != 0 && stop != 25) return fail(h, FF_ERR_BAD_ARG, "the fp32 path of this variant taps the logits (25) only");
  if (h->kind == 1) {
    int rc1 = rvk_features_fp32(h, x, layout, n, featf, st);
    if (rc1) return rc1;
  } else
  for (int s0 = 0; s0 < n; s0 += chunk) {
    const int ns = std::min(chunk, n - s0);
    const uint8_t* xin = reinterpret_cast<const uint8_t*>(x) + (size_t)s0 * crop_in_bytes;
    float* cur = h->fA;
    float* nxt = h->fB;
    for (int li = 0; li < 17; ++li) {
      const ConvPlan& p = kConv[li];
      const ConvLayerDev& L = h->conv[li];
      float* dst = (li == 16) ? featf + (size_t)s0 * PATCH : nxt;
      const int ohw = p.pool ? p.hw / 2 : p.hw;
      const size_t total = (size_t)ns * ohw * ohw * p.cout;
      const int blocks = (int)std::min<size_t>((total + 255) / 256, 1u << 30);
      if (li == 0)
        conv3x3_fp32_kernel<<<blocks, 256, 0, st>>>(xin, layout == FF_X_NHWC_U8 ? 2 : 1, nullptr, L.wf, L.scale, L.shift, dst,
                                                   ns, p.hw, p.cin, p.cout, p.pool ? 1 : 0, 1);
      else
        conv3x3_fp32_kernel<<<blocks, 256, 0, st>>>(nullptr, 0, cur, L.wf, L.scale, L.shift, dst, ns, p.hw, p.cin, p.cout,
                                                   p.pool ? 1 : 0, 1);
      FF_LAUNCH_CHECK(h, "conv3x3_fp32");
      if (tap_hit(li + 1, dst, (int64_t)total)) return FF_OK;
      std::swap(cur, nxt);
      if (h->kind == 2 && li == 7) {     // features1.26: Conv2d(128,128) without BN / activation
        const ff_cvit::RvkOp& op = h->rvk_ops[0];
        conv3x3_fp32_kernel<<<blocks, 256, 0, st>>>(nullptr, 0, cur, op.wf, op.scale, op.shift, nxt, ns, 56, 128, 128, 0, 0);
        FF_LAUNCH_CHECK(h, "features1.26_fp32");
        std::swap(cur, nxt);
      }
    }
  }
  if (h->kind == 2) {                    // x = x * ggca(x)
    int rc2 = ggca_gate_fp32(h, featf, n, st);
    if (rc2) return rc2;
  }
  const int rows = 2 * n;
  float* xn = reinterpret_cast<float*>(h->fA);                  // [rows][1024]
  float* att = xn + (size_t)h->rows_cap * DIM;                  // [rows][1024]
  float* ffh = att + (size_t)h->rows_cap * DIM;                 // [rows][2048]
  auto lin = [&](const float* A, const LinearDev& L, int M, float* out, int act, int resid, const char* what) -> int {
    dim3 grid((L.out_f + 63) / 64, (M + 63) / 64);
    linear_fp32_kernel<<<grid, 256, 0, st>>>(A, L.wf, L.b, out, M, L.out_f, L.in_f, act, resid);
    FF_LAUNCH_CHECK(h, what);
    return FF_OK;
  };
  int rc;
  if ((rc = lin(featf, h->embed, n, h->emb, 0, 0, "embed_fp32"))) return rc;
  tokens_kernel<<<n, 256, 0, st>>>(h->emb, 1, 0, nullptr, h->cls, h->pos, slot, slot_base, h->x, n, nullptr, nullptr, nullptr);
  FF_LAUNCH_CHECK(h, "tokens");
  if (tap_hit(18, h->x, (int64_t)rows * DIM)) return FF_OK;
  for (int l = 0; l < DEPTH; ++l) {
    const XfLayerDev& X = h->xf[l];
    layernorm_f32_kernel<<<(rows + 7) / 8, 256, 0, st>>>(h->x, X.ln1_g, X.ln1_b, xn, rows, 1e-5f);
    FF_LAUNCH_CHECK(h, "layernorm_f32");
    if ((rc = lin(xn, X.qkv, rows, h->qkv, 0, 0, "qkv_fp32"))) return rc;
    attention2_f32_kernel<<<(n * 8 + 7) / 8, 256, 0, st>>>(h->qkv, att, n);
    FF_LAUNCH_CHECK(h, "attention2_f32");
    if ((rc = lin(att, X.out, rows, h->x, 0, 1, "out_fp32"))) return rc;
    layernorm_f32_kernel<<<(rows + 7) / 8, 256, 0, st>>>(h->x, X.ln2_g, X.ln2_b, xn, rows, h->ln2_eps);
    FF_LAUNCH_CHECK(h, "layernorm_f32");
    if ((rc = lin(xn, X.ff1, rows, ffh, 2, 0, "ff1_fp32"))) return rc;
    if ((rc = lin(ffh, X.ff2, rows, h->x, 0, 1, "ff2_fp32"))) return rc;
    if (tap_hit(19 + l, h->x, (int64_t)rows * DIM)) return FF_OK;
  }
  cls_gather_f32_kernel<<<n, 256, 0, st>>>(h->x, xn, n);
  FF_LAUNCH_CHECK(h, "cls_gather_f32");
  if ((rc = lin(xn, h->head1, n, h->hid, 1, 0, "head1_fp32"))) return rc;
  if (h->kind == 1) {
    if ((rc = kan_head(h, n, logits, st))) return rc;
  } else {
    head2_kernel<<<(n + 7) / 8, 256, 0, st>>>(h->hid, h->head2.wf, h->head2.b, logits, n);
    FF_LAUNCH_CHECK(h, "head2");
  }
  if (tap_hit(25, logits, (int64_t)n * 2)) return FF_OK;
  return FF_OK;
}

// caller holds h->mu and a DeviceGuard
int forward_all(ff_cvit* h, const void* x, int layout, const int32_t* slot, int n, float* logits, cudaStream_t st,
                DebugTap* tap) {
  if (!h->finalized) return fail(h, FF_ERR_STATE, "weights not finalized");
  if (n < 0 || (n > 0 && (!x || !logits))) return fail(h, FF_ERR_BAD_ARG, "bad forward arguments");
  if (layout != FF_X_NCHW_F32 && layout != FF_X_NHWC_U8) return fail(h, FF_ERR_BAD_ARG, "unknown x_layout %d", layout);
  // the workspace is shared by all calls on this handle: order this call after the previous one (any stream)
  FF_CUDA(h, cudaStreamWaitEvent(st, h->done_ev, 0));
  const size_t crop_in_bytes = layout == FF_X_NHWC_U8 ? (size_t)224 * 224 * 3 : (size_t)224 * 224 * 3 * 4;
  int rc = FF_OK;
  for (int p0 = 0; p0 < n && rc == FF_OK; p0 += h->cap) {
    const int np = std::min(h->cap, n - p0);
    rc = forward_pass(h, reinterpret_cast<const uint8_t*>(x) + (size_t)p0 * crop_in_bytes, layout,
                      slot ? slot + p0 : nullptr, p0, np, logits + (size_t)2 * p0, st, tap);
    if (tap && tap->hit) break;
  }
  cudaEventRecord(h->done_ev, st);
  return rc;
}

bool mode_ok(int mode) { return mode == FF_REDUCE_REFERENCE || mode == FF_REDUCE_SOFTMAX_MEAN || mode == FF_REDUCE_REFERENCE_PROBS; }

template <typename T>
int grow(ff_cvit* h, T** p, size_t* cap, size_t need) {
  if (need <= *cap) return FF_OK;
  if (*p) cudaFree(*p);
  *p = nullptr;
  *cap = 0;
  void* q = nullptr;
  FF_CUDA(h, cudaMalloc(&q, need * sizeof(T)));
  *p = reinterpret_cast<T*>(q);
  *cap = need;
  return FF_OK;
}

// caller holds h->mu and a DeviceGuard
int predict_locked(ff_cvit* h, const void* x, int x_layout, const int32_t* off_host, const int32_t* off_dev, int n_videos,
                   int mode, float* logits_out, float* scores, cudaStream_t st) {
  if (n_videos < 0 || (n_videos > 0 && (!off_host || !off_dev || !scores))) return fail(h, FF_ERR_BAD_ARG, "ff_cvit_predict: bad arguments");
  if (mode != FF_REDUCE_REFERENCE && mode != FF_REDUCE_SOFTMAX_MEAN) return fail(h, FF_ERR_BAD_ARG, "unknown reduction mode %d", mode);
  if (n_videos == 0) return FF_OK;
  for (int v = 0; v < n_videos; ++v)
    if (off_host[v + 1] < off_host[v]) return fail(h, FF_ERR_BAD_ARG, "video_offsets must be non-decreasing");
  if (off_host[0] != 0) return fail(h, FF_ERR_BAD_ARG, "video_offsets[0] must be 0");
  const int n = off_host[n_videos];
  int rc;
  if ((rc = grow(h, &h->slot_buf, &h->slot_cap, (size_t)std::max(n, 1)))) return rc;
  float* lg = logits_out;
  if (!lg) {
    if ((rc = grow(h, &h->logit_buf, &h->logit_cap, (size_t)std::max(n, 1) * 2))) return rc;
    lg = h->logit_buf;
  }
  if (n > 0) {
    slots_from_offsets_kernel<<<n_videos, 64, 0, st>>>(off_dev, n_videos, h->slot_buf);
    FF_LAUNCH_CHECK(h, "slots_from_offsets");
    if ((rc = forward_all(h, x, x_layout, h->slot_buf, n, lg, st, nullptr))) return rc;
  }
  video_reduce_kernel<<<(n_videos + 7) / 8, 256, 0, st>>>(lg, off_dev, n_videos, mode, MAX_FRAMES, scores);
  FF_LAUNCH_CHECK(h, "video_reduce");
  return FF_OK;
}

int create_impl(ff_cvit_t** out, int device, int max_crops, int compute_dtype, int kind) {
  if (!out || max_crops <= 0) return fail(nullptr, FF_ERR_BAD_ARG, "ff_cvit_create: bad arguments");
  if (compute_dtype != FF_COMPUTE_BF16 && compute_dtype != FF_COMPUTE_FP32)
    return fail(nullptr, FF_ERR_BAD_ARG, "ff_cvit_create: unknown compute_dtype %d", compute_dtype);
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev <= 0)
    return fail(nullptr, FF_ERR_CUDA, "no CUDA device (%s): libfacfake has no CPU fallback", cudaGetErrorString(e));
  if (device < 0 || device >= ndev) return fail(nullptr, FF_ERR_BAD_ARG, "device %d out of range (%d devices)", device, ndev);
  cudaDeviceProp prop;
  if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess)
    return fail(nullptr, FF_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
  if (prop.major != 10)
    return fail(nullptr, FF_ERR_CUDA, "device %d is sm_%d%d; libfacfake is built for sm_100a (B200) only", device, prop.major, prop.minor);
  ffh::DeviceGuard guard(device);
  if (guard.status != cudaSuccess) return fail(nullptr, FF_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(guard.status));
  if (!ffh::encode_tiled()) return fail(nullptr, FF_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
  ff_cvit* h = new ff_cvit();
  h->device = device;
  h->compute = compute_dtype;
  h->kind = kind;
  h->act_f16 = kind == 2 && compute_dtype == FF_COMPUTE_BF16;
  h->ln2_eps = kind == 2 ? 1e-6f : 1e-5f;
  h->cap = (max_crops + 31) / 32 * 32;
  h->rows_cap = (2 * h->cap + 127) / 128 * 128;
  h->s12_cap = 256;
  h->s12 = std::min(64, h->cap);
  cudaDeviceGetAttribute(&h->num_sms, cudaDevAttrMultiProcessorCount, device);
  int rc = FF_OK;
  const int cap128 = (h->cap + 127) / 128 * 128;
  do {
    if (cudaEventCreateWithFlags(&h->done_ev, cudaEventDisableTiming) != cudaSuccess) { rc = fail(h, FF_ERR_CUDA, "event create failed"); break; }
    if ((rc = dev_alloc(h, &h->emb, (size_t)EMBED_SPLITS * h->cap * DIM))) break;
    if ((rc = dev_alloc(h, &h->x, (size_t)h->rows_cap * DIM))) break;
    if ((rc = dev_alloc(h, &h->qkv, (size_t)h->rows_cap * 3 * DIM))) break;
    if ((rc = dev_alloc(h, &h->hid, (size_t)cap128 * MLP))) break;
    if (kind == 1) {
      constexpr size_t kRvkActElems = (size_t)112 * 112 * 64;      // largest activation per crop (stem / layer1 output)
      if (compute_dtype == FF_COMPUTE_BF16) {
        for (int i = 0; i < 5 && rc == FF_OK; ++i) rc = dev_alloc(h, &h->rvk_buf[i], (size_t)h->cap * kRvkActElems);
        if (rc) break;
        if ((rc = dev_alloc(h, &h->rvk_x4, (size_t)h->cap * 224 * 224 * 4))) break;
      } else {
        h->rvk_f32_chunk = std::min(h->cap, 32);
        for (int i = 0; i < 5 && rc == FF_OK; ++i) rc = dev_alloc(h, &h->rvk_f32[i], (size_t)h->rvk_f32_chunk * kRvkActElems);
        if (rc) break;
        if ((rc = dev_alloc(h, &h->rvk_x4f, (size_t)h->rvk_f32_chunk * 224 * 224 * 4))) break;
      }
      if ((rc = dev_alloc(h, &h->kan_part, (size_t)KAN_PART_CHUNKS * h->cap * 64))) break;
    }
    if (kind == 2 && compute_dtype == FF_COMPUTE_BF16 && (rc = dev_alloc(h, &h->bufR, (size_t)h->cap * 56 * 56 * 128))) break;
    if (compute_dtype == FF_COMPUTE_BF16 && kind != 1) {
      if ((rc = dev_alloc(h, &h->bufA, (size_t)h->s12_cap * 224 * 224 * 32))) break;
      if ((rc = dev_alloc(h, &h->bufB, (size_t)h->s12_cap * 224 * 224 * 32))) break;
      if ((rc = dev_alloc(h, &h->bufA2, (size_t)h->s12_cap * 224 * 224 * 32))) break;
      if ((rc = dev_alloc(h, &h->bufB2, (size_t)h->s12_cap * 224 * 224 * 32))) break;
      if (cudaStreamCreateWithFlags(&h->aux_stream, cudaStreamNonBlocking) != cudaSuccess ||
          cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
          cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming) != cudaSuccess) { rc = fail(h, FF_ERR_CUDA, "aux stream/event create failed"); break; }
      if ((rc = dev_alloc(h, &h->P, (size_t)h->cap * 56 * 56 * 128))) break;
      if ((rc = dev_alloc(h, &h->Q, (size_t)h->cap * 56 * 56 * 128))) break;
    }
    if (compute_dtype == FF_COMPUTE_BF16) {
      if ((rc = dev_alloc(h, &h->feat, (size_t)cap128 * PATCH))) break;
      if ((rc = dev_alloc(h, &h->xn, (size_t)h->rows_cap * DIM))) break;
      if ((rc = dev_alloc(h, &h->att, (size_t)h->rows_cap * DIM))) break;
      if ((rc = dev_alloc(h, &h->ffh_buf, (size_t)h->rows_cap * MLP))) break;
      if ((rc = dev_alloc(h, &h->qkvb, (size_t)h->rows_cap * 3 * DIM))) break;
      if ((rc = dev_alloc(h, &h->clsb, (size_t)cap128 * DIM))) break;
      // rows beyond the valid ones are read by TMA (results masked): keep them finite
      cudaMemset(h->feat, 0, (size_t)cap128 * PATCH * 2);
      cudaMemset(h->xn, 0, (size_t)h->rows_cap * DIM * 2);
      cudaMemset(h->att, 0, (size_t)h->rows_cap * DIM * 2);
      cudaMemset(h->ffh_buf, 0, (size_t)h->rows_cap * MLP * 2);
      cudaMemset(h->clsb, 0, (size_t)cap128 * DIM * 2);
    } else {
      const size_t act = kind == 1 ? 0 : (size_t)h->s12_cap * 224 * 224 * 32;   // the ResNet trunk has its own fp32 buffers
      const size_t tail = std::max((size_t)h->rows_cap * DIM * 4, (size_t)1);
      if ((rc = dev_alloc(h, &h->fA, std::max(act, tail)))) break;
      if ((rc = dev_alloc(h, &h->fB, act + (size_t)h->cap * PATCH))) break;
      h->featf = h->fB + act;
    }
  } while (0);
  if (rc != FF_OK) {
    g_create_error = h->err;
    ff_cvit_destroy(h);
    return rc;
  }
  *out = h;
  return FF_OK;
}

}  // namespace
}  // namespace ffe

// ================================================================================================ C-ABI
using namespace ffe;

extern "C" {

const char* ff_last_error(const ff_cvit_t* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int ff_cvit_create(ff_cvit_t** out, int device, int max_crops, int compute_dtype) {
  return create_impl(out, device, max_crops, compute_dtype, 0);
}
int ff_resvitkan_create(ff_cvit_t** out, int device, int max_crops, int compute_dtype) {
  return create_impl(out, device, max_crops, compute_dtype, 1);
}
int ff_cvit_ggca_create(ff_cvit_t** out, int device, int max_crops, int compute_dtype) {
  return create_impl(out, device, max_crops, compute_dtype, 2);
}

void ff_cvit_destroy(ff_cvit_t* h) {
  if (!h) return;
  {
    ffh::DeviceGuard guard(h->device);
    cudaDeviceSynchronize();
    for (void* p : h->allocs) cudaFree(p);
    if (h->slot_buf) cudaFree(h->slot_buf);
    if (h->xin_buf) cudaFree(h->xin_buf);
    if (h->logit_buf) cudaFree(h->logit_buf);
    if (h->off_buf) cudaFree(h->off_buf);
    if (h->score_buf) cudaFree(h->score_buf);
    if (h->crop_desc) cudaFree(h->crop_desc);
    if (h->done_ev) cudaEventDestroy(h->done_ev);
    for (cudaEvent_t e : h->ev_pool) cudaEventDestroy(e);
    for (cudaEvent_t e : h->h2d_ready) cudaEventDestroy(e);
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    if (h->aux_stream) cudaStreamDestroy(h->aux_stream);
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    if (h->ev_join) cudaEventDestroy(h->ev_join);
  }
  delete h;
}

int ff_cvit_load_weight(ff_cvit_t* h, const char* key, const float* host_fp32, const int64_t* shape, int ndim) {
  if (!h) return FF_ERR_BAD_ARG;
  std::lock_guard<std::mutex> lk(h->mu);
  if (!key || (!host_fp32 && ndim > 0) || ndim < 0 || ndim > 4) return fail(h, FF_ERR_BAD_ARG, "ff_cvit_load_weight: bad arguments");
  if (h->finalized) return fail(h, FF_ERR_STATE, "weights already finalized");
  const std::string k(key);
  if (k.size() > 19 && k.compare(k.size() - 19, 19, "num_batches_tracked") == 0) return FF_OK;
  int64_t cnt = 1;
  std::vector<int64_t> shp;
  for (int i = 0; i < ndim; ++i) {
    if (shape[i] <= 0) return fail(h, FF_ERR_SHAPE, "weight '%s': non-positive dimension", key);
    cnt *= shape[i];
    shp.push_back(shape[i]);
  }
  if (cnt > (int64_t)1024 * 25088) return fail(h, FF_ERR_SHAPE, "weight '%s' too large for CViT", key);
  h->host_w[k].assign(host_fp32, host_fp32 + cnt);
  h->host_shape[k] = shp;
  return FF_OK;
}

int ff_cvit_finalize_weights(ff_cvit_t* h) {
  if (!h) return FF_ERR_BAD_ARG;
  std::lock_guard<std::mutex> lk(h->mu);
  if (h->finalized) return fail(h, FF_ERR_STATE, "weights already finalized");
  ffh::DeviceGuard guard(h->device);
  return finalize(h);
}

const char* ff_cvit_unused_keys(const ff_cvit_t* h) { return h ? h->unused_keys.c_str() : ""; }

int ff_cvit_forward(ff_cvit_t* h, const void* x, int x_layout, const int32_t* slot, int n, float* logits, void* stream) {
  if (!h) return FF_ERR_BAD_ARG;
  std::lock_guard<std::mutex> lk(h->mu);
  ffh::DeviceGuard guard(h->device);
  return forward_all(h, x, x_layout, slot, n, logits, reinterpret_cast<cudaStream_t>(stream), nullptr);
}

int ff_video_scores(ff_cvit_t* h, const float* logits, const int32_t* video_offsets, int n_videos, int mode, float* scores,
                    void* stream) {
  if (!h) return FF_ERR_BAD_ARG;
  std::lock_guard<std::mutex> lk(h->mu);
  if (n_videos < 0 || (n_videos > 0 && (!video_offsets || !scores))) return fail(h, FF_ERR_BAD_ARG, "ff_video_scores: bad arguments");
  if (!mode_ok(mode)) return fail(h, FF_ERR_BAD_ARG, "unknown reduction mode %d", mode);
  if (n_videos == 0) return FF_OK;
  ffh::DeviceGuard guard(h->device);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  video_reduce_kernel<<<(n_videos + 7) / 8, 256, 0, st>>>(logits, video_offsets, n_videos, mode, MAX_FRAMES, scores);
  FF_LAUNCH_CHECK(h, "video_reduce");
  return FF_OK;
}

int ff_cvit_predict(ff_cvit_t* h, const void* x, int x_layout, const int32_t* off_host, const int32_t* off_dev, int n_videos,
                    int mode, float* logits_out, float* scores, void* stream) {
  if (!h) return FF_ERR_BAD_ARG;
  std::lock_guard<std::mutex> lk(h->mu);
  ffh::DeviceGuard guard(h->device);
  return predict_locked(h, x, x_layout, off_host, off_dev, n_videos, mode, logits_out, scores, reinterpret_cast<cudaStream_t>(stream));
}

int ff_cvit_predict_host(ff_cvit_t* h, const uint8_t* x_host, const int32_t* off_host, int n_videos, int mode, float* scores_host,
                         void* stream) {
  if (!h) return FF_ERR_BAD_ARG;
  // One lock for staging, forward, reduction and the score copy: the staging buffers and the chunk events belong to the
  // handle, and the reference drives predict() from a thread pool (cvit_prediction.py:73-83).
  std::lock_guard<std::mutex> lk(h->mu);
  if (n_videos < 0 || (n_videos > 0 && (!off_host || !scores_host))) return fail(h, FF_ERR_BAD_ARG, "ff_cvit_predict_host: bad arguments");
  if (n_videos == 0) return FF_OK;
  const int n = off_host[n_videos];
  if (n < 0 || (n > 0 && !x_host)) return fail(h, FF_ERR_BAD_ARG, "ff_cvit_predict_host: bad arguments");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  ffh::DeviceGuard guard(h->device);
  int rc;
  if ((rc = grow(h, &h->xin_buf, &h->xin_cap, (size_t)std::max(n, 1) * 224 * 224 * 3))) return rc;
  if ((rc = grow(h, &h->off_buf, &h->off_cap, (size_t)n_videos + 1))) return rc;
  if ((rc = grow(h, &h->score_buf, &h->score_cap, (size_t)n_videos))) return rc;
  FF_CUDA(h, cudaStreamWaitEvent(st, h->done_ev, 0));
  FF_CUDA(h, cudaMemcpyAsync(h->off_buf, off_host, ((size_t)n_videos + 1) * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  // crops stream in on a second stream, one chunk at a time, overlapping the forward of earlier chunks
  if (!h->copy_stream) FF_CUDA(h, cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
  h->h2d_chunk = std::min(h->s12, 32);
  const int chunks = (n + h->h2d_chunk - 1) / h->h2d_chunk;
  while ((int)h->h2d_ready.size() < chunks) {
    cudaEvent_t e;
    FF_CUDA(h, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    h->h2d_ready.push_back(e);
  }
  FF_CUDA(h, cudaStreamWaitEvent(h->copy_stream, h->done_ev, 0));   // previous call has finished reading xin_buf
  const size_t crop_bytes = (size_t)224 * 224 * 3;
  for (int c = 0; c < chunks; ++c) {
    const int c0 = c * h->h2d_chunk, cn = std::min(h->h2d_chunk, n - c0);
    FF_CUDA(h, cudaMemcpyAsync(h->xin_buf + c0 * crop_bytes, x_host + c0 * crop_bytes, cn * crop_bytes, cudaMemcpyHostToDevice,
                               h->copy_stream));
    FF_CUDA(h, cudaEventRecord(h->h2d_ready[c], h->copy_stream));
  }
  h->h2d_chunks_pending = chunks;
  rc = predict_locked(h, h->xin_buf, FF_X_NHWC_U8, off_host, h->off_buf, n_videos, mode, nullptr, h->score_buf, st);
  h->h2d_chunks_pending = 0;
  if (rc) return rc;
  FF_CUDA(h, cudaMemcpyAsync(scores_host, h->score_buf, (size_t)n_videos * sizeof(float), cudaMemcpyDeviceToHost, st));
  FF_CUDA(h, cudaStreamSynchronize(st));
  return FF_OK;
}

int ff_preprocess_crops(ff_cvit_t* h, const uint8_t* const* crop_ptrs, const int32_t* hw, const int32_t* pitch, int n, int swap_rb,
                        uint8_t* out_u8, float* out_norm_nchw, void* stream) {
  if (!h) return FF_ERR_BAD_ARG;
  std::lock_guard<std::mutex> lk(h->mu);
  if (n < 0 || (n > 0 && (!crop_ptrs || !hw || !pitch || (!out_u8 && !out_norm_nchw)))) return fail(h, FF_ERR_BAD_ARG, "ff_preprocess_crops: bad arguments");
  if (n == 0) return FF_OK;
  std::vector<CropDesc> d(n);
  for (int i = 0; i < n; ++i) {
    const int ch = hw[2 * i], cw = hw[2 * i + 1];
    if (!crop_ptrs[i] || ch <= 0 || cw <= 0 || pitch[i] < cw * 3) return fail(h, FF_ERR_SHAPE, "crop %d: bad pointer/size/pitch", i);
    d[i] = make_crop_desc(crop_ptrs[i], ch, cw, pitch[i], 224);
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  ffh::DeviceGuard guard(h->device);
  int rc;
  CropDesc* dd = reinterpret_cast<CropDesc*>(h->crop_desc);
  if ((rc = grow(h, &dd, &h->crop_desc_cap, (size_t)n))) { h->crop_desc = dd; return rc; }
  h->crop_desc = dd;
  // pageable-host source: cudaMemcpyAsync stages the descriptors before returning, so `d` may die with this call;
  // the descriptor buffer is reused by the next call only after the stream-ordered kernel below has been enqueued
  FF_CUDA(h, cudaMemcpyAsync(dd, d.data(), sizeof(CropDesc) * n, cudaMemcpyHostToDevice, st));
  {
    ProfScope ps(h, st, KC_SMALL);
    preprocess_kernel<224><<<dim3(224 / 4, n), 256, 0, st>>>(dd, n, swap_rb, out_u8, out_norm_nchw);
  }
  FF_LAUNCH_CHECK(h, "preprocess");
  return FF_OK;
}

int64_t ff_cvit_launch_count(const ff_cvit_t* h) { return h ? h->launches : 0; }

int64_t ff_cvit_debug_activation(ff_cvit_t* h, const void* x, int x_layout, const int32_t* slot, int n, int stop_after,
                                 float* out_host, int64_t out_elems, void* stream) {
  if (!h) return FF_ERR_BAD_ARG;
  std::lock_guard<std::mutex> lk(h->mu);
  if (stop_after < 1 || stop_after > 27 || !out_host || n <= 0 || n > h->cap) return fail(h, FF_ERR_BAD_ARG, "ff_cvit_debug_activation: bad arguments");
  ffh::DeviceGuard guard(h->device);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  DebugTap tap;
  tap.stop_after = stop_after;
  float* lg = nullptr;
  if (cudaMalloc(&lg, (size_t)n * 2 * sizeof(float)) != cudaSuccess) return fail(h, FF_ERR_CUDA, "debug alloc failed");
  int rc = forward_all(h, x, x_layout, slot, n, lg, st, &tap);
  int64_t ret = rc;
  if (rc == FF_OK) {
    if (!tap.hit) ret = fail(h, FF_ERR_STATE, "debug tap %d not reached", stop_after);
    else if (tap.elems > out_elems) ret = fail(h, FF_ERR_BAD_ARG, "debug buffer too small: need %lld floats", (long long)tap.elems);
    else {
      cudaError_t e;
      if (tap.is_16) {
        float* tmp = nullptr;
        e = cudaMalloc(&tmp, (size_t)tap.elems * sizeof(float));
        if (e == cudaSuccess) {
          const unsigned short* src = reinterpret_cast<const unsigned short*>(tap.ptr);
          // every 16-bit tap is a conv-stack activation in the handle's 16-bit type, except the gated map (always bf16)
          const int f16 = (h->act_f16 && stop_after != 27) ? 1 : 0;
          if (tap.blocked_hw)
            unblock_act16_to_f32_kernel<<<(unsigned)((tap.elems + 255) / 256), 256, 0, st>>>(src, tmp, n, tap.blocked_hw, f16);
          else
            act16_to_f32_kernel<<<(unsigned)((tap.elems + 255) / 256), 256, 0, st>>>(src, tmp, (size_t)tap.elems, f16);
          e = cudaMemcpyAsync(out_host, tmp, (size_t)tap.elems * sizeof(float), cudaMemcpyDeviceToHost, st);
          if (e == cudaSuccess) e = cudaStreamSynchronize(st);
          cudaFree(tmp);
        }
      } else {
        e = cudaMemcpyAsync(out_host, tap.ptr, (size_t)tap.elems * sizeof(float), cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
      }
      ret = (e == cudaSuccess) ? tap.elems : fail(h, FF_ERR_CUDA, "debug copy failed: %s", cudaGetErrorString(e));
    }
  }
  cudaStreamSynchronize(st);
  cudaFree(lg);
  return ret;
}

int ff_cvit_set_profiling(ff_cvit_t* h, int enable) {
  if (!h) return FF_ERR_BAD_ARG;
  std::lock_guard<std::mutex> lk(h->mu);
  h->profiling = enable != 0;
  h->prof_coarse = enable == 2;
  h->ev_used = 0;
  h->ev_class.clear();
  for (int i = 0; i < KC_COUNT; ++i) { h->prof_ms[i] = 0; h->prof_launches[i] = 0; }
  return FF_OK;
}

int ff_cvit_get_profile(ff_cvit_t* h, double* ms_by_class, int64_t* launches_by_class) {
  if (!h || !ms_by_class || !launches_by_class) return FF_ERR_BAD_ARG;
  std::lock_guard<std::mutex> lk(h->mu);
  ffh::DeviceGuard guard(h->device);
  const size_t pairs = std::min(h->ev_class.size(), h->ev_used / 2);
  if (pairs > 0) FF_CUDA(h, cudaEventSynchronize(h->ev_pool[2 * pairs - 1]));
  for (size_t i = 0; i < pairs; ++i) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, h->ev_pool[2 * i], h->ev_pool[2 * i + 1]) == cudaSuccess) {
      h->prof_ms[h->ev_class[i]] += ms;
      h->prof_launches[h->ev_class[i]] += 1;
    }
  }
  h->ev_used = 0;
  h->ev_class.clear();
  for (int i = 0; i < KC_COUNT; ++i) { ms_by_class[i] = h->prof_ms[i]; launches_by_class[i] = h->prof_launches[i]; }
  return FF_OK;
}

int ff_cvit_set_tuning(ff_cvit_t* h, int stage12_sub_batch) {
  if (!h) return FF_ERR_BAD_ARG;
  std::lock_guard<std::mutex> lk(h->mu);
  if (stage12_sub_batch < 0 || stage12_sub_batch > h->s12_cap) return fail(h, FF_ERR_BAD_ARG, "stage12_sub_batch must be in [1,%d]", h->s12_cap);
  if (stage12_sub_batch > 0) h->s12 = std::min(stage12_sub_batch, h->cap);
  return FF_OK;
}

}  // extern "C"
