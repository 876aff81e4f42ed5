// ff_blaze.cu — BlazeFace face detector on the GPU: the network and the box decoding of
// /root/reference/CViT-main/helpers/blazeface.py (:8-43 BlazeBlock, :80-148 forward, :236-303 decode) behind the
// `ff_blazeface_*` entry points of include/facfake.h.  The blending NMS (:305-358) is data dependent and stays on the
// host in the mirror class, exactly where the reference runs it.
//
// The detector is 30 MFLOP per 128x128 tile and its detections are thresholded (score >= 0.75, IoU > 0.3), so the
// whole path is fp32 on the CUDA cores with the reference's operation order; the cost that matters is launches:
// 1 (5x5 stride-2 stem) + 16 (one fused kernel per BlazeBlock: depthwise 3x3 -> pointwise 1x1 -> + max-pooled /
// channel-padded residual -> ReLU) + 1 (four 1x1 heads, anchor-major) + 1 (decode + sigmoid) per batch of tiles.
// Activations are NHWC fp32; no tensor cores: K is 24..96 and the tiles are tiny, HBM/L2 traffic and launch count
// bound it.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/facfake.h"
#include "ff_host.h"
#include "ff_pre.cuh"

namespace {

struct BlockPlan { int cin, cout, stride, hw_in; };
// backbone1.2..12 and backbone2.0..4 (blazeface.py:86-107); hw_in = input spatial size
const BlockPlan kBlocks[16] = {
    {24, 24, 1, 64}, {24, 28, 1, 64}, {28, 32, 2, 64}, {32, 36, 1, 32}, {36, 42, 1, 32}, {42, 48, 2, 32},
    {48, 56, 1, 16}, {56, 64, 1, 16}, {64, 72, 1, 16}, {72, 80, 1, 16}, {80, 88, 1, 16},
    {88, 96, 2, 16}, {96, 96, 1, 8},  {96, 96, 1, 8},  {96, 96, 1, 8},  {96, 96, 1, 8},
};
constexpr int NUM_ANCHORS = 896;
constexpr size_t ACT_ELEMS = (size_t)64 * 64 * 28;   // largest activation per tile

std::string g_blaze_create_error;

}  // namespace

struct ff_blazeface {
  int device = 0;
  int cap = 0;
  bool finalized = false;
  std::string err;
  std::mutex mu;
  std::map<std::string, std::vector<float>> host_w;
  std::map<std::string, std::vector<int64_t>> host_shape;
  std::vector<void*> allocs;
  float *stem_w = nullptr, *stem_b = nullptr;             // [5][5][3][24] (tap-major, cout fastest), [24]
  float *dw_w[16] = {}, *dw_b[16] = {};                   // [9][cin], [cin]
  float *pw_w[16] = {}, *pw_b[16] = {};                   // [cin][cout] (transposed: coalesced over cout), [cout]
  float *head8_w = nullptr, *head8_b = nullptr;           // [88][34]: 2 classifier + 32 regressor outputs per cell
  float *head16_w = nullptr, *head16_b = nullptr;         // [96][102]: 6 + 96
  float* anchors = nullptr;                               // [896][4]
  float *act_a = nullptr, *act_b = nullptr, *feat8 = nullptr;   // ping-pong activations; 16x16x88 map kept for the heads
  float *raw_boxes = nullptr, *raw_scores = nullptr;      // [cap][896][16], [cap][896]
  int64_t launches = 0;
  void* tile_desc = nullptr; size_t tile_desc_cap = 0;    // crop descriptors of ff_blazeface_tile_frames
};

namespace {

int bfail(ff_blazeface* h, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (h) h->err = buf; else g_blaze_create_error = buf;
  return code;
}

#define BZ_CUDA(h, call)                                                                                          \
  do {                                                                                                            \
    cudaError_t e_ = (call);                                                                                      \
    if (e_ != cudaSuccess) return bfail(h, FF_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_));          \
  } while (0)

// ---- stem: x/127.5 - 1, zero pad (1,2,1,2) of the PREPROCESSED image, Conv2d(3,24,5,stride 2) + ReLU (:113-116,162-164)
// CTA = 16x16 output pixels of one tile: the 35x35x3 input patch is preprocessed once into shared memory (zero =
// padding), the 75x24 filter sits in shared memory too; one thread = one pixel x all 24 output channels.
__global__ void __launch_bounds__(256)
blaze_stem_kernel(const uint8_t* __restrict__ tiles, const float* __restrict__ w, const float* __restrict__ b,
                  float* __restrict__ out, int n) {
  __shared__ float s_in[35 * 35 * 3];
  __shared__ __align__(16) float s_w[75 * 24];
  const size_t img = blockIdx.z;
  const int ty0 = blockIdx.y * 16, tx0 = blockIdx.x * 16;
  const uint8_t* src = tiles + img * 128 * 128 * 3;
  for (int i = threadIdx.x; i < 75 * 24; i += 256) s_w[i] = w[i];
  for (int i = threadIdx.x; i < 35 * 35 * 3; i += 256) {
    const int c = i % 3, px = (i / 3) % 35, py = i / (3 * 35);
    const int iy = 2 * ty0 + py - 1, ix = 2 * tx0 + px - 1;
    float v = 0.0f;
    if (iy >= 0 && iy < 128 && ix >= 0 && ix < 128) v = __fdiv_rn((float)src[((size_t)iy * 128 + ix) * 3 + c], 127.5f) - 1.0f;
    s_in[i] = v;
  }
  __syncthreads();
  const int ly = threadIdx.x >> 4, lx = threadIdx.x & 15;
  float acc[24];
#pragma unroll
  for (int co = 0; co < 24; ++co) acc[co] = b[co];
  for (int kh = 0; kh < 5; ++kh) {
#pragma unroll
    for (int kw = 0; kw < 5; ++kw) {
      const float* ip = &s_in[((2 * ly + kh) * 35 + 2 * lx + kw) * 3];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float v = ip[c];
        const float4* wp = reinterpret_cast<const float4*>(&s_w[((kh * 5 + kw) * 3 + c) * 24]);
#pragma unroll
        for (int q = 0; q < 6; ++q) {
          const float4 w4 = wp[q];
          acc[4 * q] = fmaf(v, w4.x, acc[4 * q]);
          acc[4 * q + 1] = fmaf(v, w4.y, acc[4 * q + 1]);
          acc[4 * q + 2] = fmaf(v, w4.z, acc[4 * q + 2]);
          acc[4 * q + 3] = fmaf(v, w4.w, acc[4 * q + 3]);
        }
      }
    }
  }
  float4* o = reinterpret_cast<float4*>(out + ((img * 64 + ty0 + ly) * 64 + tx0 + lx) * 24);
#pragma unroll
  for (int q = 0; q < 6; ++q)
    o[q] = make_float4(fmaxf(acc[4 * q], 0.0f), fmaxf(acc[4 * q + 1], 0.0f), fmaxf(acc[4 * q + 2], 0.0f), fmaxf(acc[4 * q + 3], 0.0f));
}

// Pointwise 1x1 micro-tile: one thread accumulates 8 consecutive pixels x 4 consecutive output channels (32 FMAs per
// input channel for two 16-byte shared-memory loads and four cached weight loads).  `dw` points at s_dw[0][first pixel],
// `pitch` is the row pitch of s_dw in floats (a multiple of 4: 16-byte aligned rows).
__device__ __forceinline__ void blaze_pw_8x4(const float* __restrict__ dw, int pitch, const float* __restrict__ pw_w, int cin, int cout,
                                             int co, int nco, float (&acc)[8][4]) {
  for (int c = 0; c < cin; ++c) {
    const float4 a0 = *reinterpret_cast<const float4*>(dw + c * pitch);
    const float4 a1 = *reinterpret_cast<const float4*>(dw + c * pitch + 4);
    const float* wp = pw_w + c * cout + co;
    float w[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) w[j] = j < nco ? wp[j] : 0.0f;
    const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
    for (int k = 0; k < 8; ++k)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[k][j] = fmaf(a[k], w[j], acc[k][j]);
  }
}

// ---- one BlazeBlock (:8-43).  CTA = PIX output pixels of one tile x all channels.
//   phase 1: depthwise 3x3 (stride 1: pad 1; stride 2: zero pad right/bottom by 2, no other padding) -> smem [PIX][cin]
//   phase 2: pointwise 1x1 + bias + residual (stride 2: 2x2 max-pool of x; channels >= cin see zeros) -> ReLU
constexpr int BLAZE_PG = 8;                        // pixels per thread in the pointwise phase
// pixels per CTA: large maps have few channels, so the [cin][PIX+1] staging stays under 32 KB everywhere
inline int blaze_pix(int hw_out) { return hw_out >= 64 ? 256 : (hw_out >= 32 ? 128 : 64); }
template <int BLAZE_PIX>
__global__ void __launch_bounds__(256)
blaze_block_kernel(const float* __restrict__ x, float* __restrict__ out, const float* __restrict__ dw_w,
                   const float* __restrict__ dw_b, const float* __restrict__ pw_w, const float* __restrict__ pw_b,
                   int cin, int cout, int stride, int hw_in) {
  extern __shared__ __align__(16) float s_dw_raw[];  // [cin][BLAZE_PIX + 4]: the pointwise phase reads 8 consecutive pixels
  float (*s_dw)[BLAZE_PIX + 4] = reinterpret_cast<float (*)[BLAZE_PIX + 4]>(s_dw_raw);   // rows 16-byte aligned
  const int hw_out = hw_in / stride;
  const int npix = hw_out * hw_out;
  const int pix0 = blockIdx.x * BLAZE_PIX;
  const size_t img = blockIdx.y;
  const float* xin = x + img * (size_t)hw_in * hw_in * cin;
  // depthwise: item = (4 consecutive output pixels of a row, channel), channel fastest (coalesced NHWC reads); the three
  // input rows x (6 | 9) columns and the 9 taps are loaded once for the 4 outputs
  for (int i = threadIdx.x; i < (BLAZE_PIX / 4) * cin; i += blockDim.x) {
    const int q = i / cin, c = i - q * cin;
    const int pix = pix0 + q * 4;                    // hw_out is a multiple of 4: the quad never wraps
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    if (pix < npix) {
      const int oy = pix / hw_out, ox = pix - oy * hw_out;
      float w[9];
#pragma unroll
      for (int t = 0; t < 9; ++t) w[t] = dw_w[t * cin + c];
      const float bias = dw_b[c];
#pragma unroll
      for (int k = 0; k < 4; ++k) acc[k] = bias;
      const int nc = stride == 2 ? 9 : 6;            // input columns touched by the quad
      const int ix0 = stride == 2 ? 2 * ox : ox - 1, iy0 = stride == 2 ? 2 * oy : oy - 1;
#pragma unroll
      for (int kh = 0; kh < 3; ++kh) {
        const int iy = iy0 + kh;
        if (iy < 0 || iy >= hw_in) continue;
        float v[9];
#pragma unroll
        for (int j = 0; j < 9; ++j) {
          const int ix = ix0 + j;
          v[j] = (j < nc && ix >= 0 && ix < hw_in) ? xin[((size_t)iy * hw_in + ix) * cin + c] : 0.0f;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
          for (int kw = 0; kw < 3; ++kw) {
            // stride 1: column k + kw; stride 2: column 2k + kw  (selected without dynamic indexing)
            const float xv = stride == 2 ? v[2 * k + kw] : v[k + kw];
            acc[k] = fmaf(xv, w[kh * 3 + kw], acc[k]);
          }
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) s_dw[c][q * 4 + k] = acc[k];
  }
  __syncthreads();
  float* o = out + img * (size_t)npix * cout;
  // item = (pixel group of 8, group of 4 output channels), channel group fastest
  const int cog = (cout + 3) / 4;
  for (int i = threadIdx.x; i < (BLAZE_PIX / BLAZE_PG) * cog; i += blockDim.x) {
    const int pg = i / cog, co = (i - pg * cog) * 4;
    const int nco = min(4, cout - co);
    float acc[8][4];
#pragma unroll
    for (int k = 0; k < 8; ++k)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[k][j] = j < nco ? pw_b[co + j] : 0.0f;
    blaze_pw_8x4(&s_dw[0][pg * BLAZE_PG], BLAZE_PIX + 4, pw_w, cin, cout, co, nco, acc);
#pragma unroll
    for (int k = 0; k < BLAZE_PG; ++k) {
      const int pix = pix0 + pg * BLAZE_PG + k;
      if (pix >= npix) continue;
      const int oy = pix / hw_out, ox = pix - oy * hw_out;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (j >= nco) break;
        float res = 0.0f;
        if (co + j < cin) {
          if (stride == 2) {
            const float* q = xin + ((size_t)(2 * oy) * hw_in + 2 * ox) * cin + co + j;
            res = fmaxf(fmaxf(q[0], q[cin]), fmaxf(q[(size_t)hw_in * cin], q[(size_t)hw_in * cin + cin]));
          } else {
            res = xin[((size_t)oy * hw_in + ox) * cin + co + j];
          }
        }
        o[(size_t)pix * cout + co + j] = fmaxf(acc[k][j] + res, 0.0f);
      }
    }
  }
}

// ---- a CHAIN of stride-1 BlazeBlocks at one resolution in a single launch: the 16x16 (blocks 6..10, 48 -> 88 channels)
// and 8x8 (blocks 12..15, 96 channels) maps of one tile fit in shared memory, so the activation makes one round trip to
// global memory per chain instead of one per block (the ncu launch list had 1.45 of 4.0 ms in these nine launches).
// One CTA = one tile.  Per block: depthwise 3x3 from the resident map into s_dw[c][p]; pointwise 1x1 + residual + ReLU
// written back IN PLACE (element (p, co) is read and written by the one thread that owns it; the channel count only
// grows, the row pitch CP is that of the widest block).
struct BlazeChain {
  int nblk;
  int cin[5], cout[5];
  const float *dw_w[5], *dw_b[5], *pw_w[5], *pw_b[5];
};
template <int HW, int CP, int THREADS>
__global__ void __launch_bounds__(THREADS)
blaze_chain_kernel(const float* __restrict__ x, float* __restrict__ out, const BlazeChain ch) {
  constexpr int P = HW * HW;
  extern __shared__ float s_chain[];
  float (*s_act)[CP + 1] = reinterpret_cast<float (*)[CP + 1]>(s_chain);                    // [P][CP + 1]
  float (*s_dw)[P + 4] = reinterpret_cast<float (*)[P + 4]>(s_chain + ((P * (CP + 1) + 3) & ~3));   // [<= CP][P + 4], rows 16-byte aligned
  const size_t img = blockIdx.x;
  const int c0 = ch.cin[0];
  const float* xin = x + img * (size_t)P * c0;
  for (int i = threadIdx.x; i < P * c0; i += THREADS) s_act[i / c0][i % c0] = xin[i];
  __syncthreads();
  for (int b = 0; b < ch.nblk; ++b) {
    const int cin = ch.cin[b], cout = ch.cout[b];
    const float* dw_w = ch.dw_w[b];
    const float* dw_b = ch.dw_b[b];
    const float* pw_w = ch.pw_w[b];
    const float* pw_b = ch.pw_b[b];
    for (int i = threadIdx.x; i < (P / 4) * cin; i += THREADS) {      // item = (4 consecutive pixels of a row, channel)
      const int q = i / cin, c = i - q * cin;
      const int p = q * 4;
      const int oy = p / HW, ox = p - oy * HW;
      float w[9];
#pragma unroll
      for (int t = 0; t < 9; ++t) w[t] = dw_w[t * cin + c];
      float acc[4];
      const float bias = dw_b[c];
#pragma unroll
      for (int k = 0; k < 4; ++k) acc[k] = bias;
#pragma unroll
      for (int kh = 0; kh < 3; ++kh) {
        const int iy = oy + kh - 1;
        if (iy < 0 || iy >= HW) continue;
        float v[6];
#pragma unroll
        for (int j = 0; j < 6; ++j) {
          const int ix = ox - 1 + j;
          v[j] = (ix >= 0 && ix < HW) ? s_act[iy * HW + ix][c] : 0.0f;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
          for (int kw = 0; kw < 3; ++kw) acc[k] = fmaf(v[k + kw], w[kh * 3 + kw], acc[k]);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) s_dw[c][p + k] = acc[k];
    }
    __syncthreads();
    const int cog = (cout + 3) / 4;
    for (int i = threadIdx.x; i < (P / BLAZE_PG) * cog; i += THREADS) {
      const int pg = i / cog, co = (i - pg * cog) * 4;
      const int nco = min(4, cout - co);
      float acc[8][4];
#pragma unroll
      for (int k = 0; k < 8; ++k)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[k][j] = j < nco ? pw_b[co + j] : 0.0f;
      blaze_pw_8x4(&s_dw[0][pg * BLAZE_PG], P + 4, pw_w, cin, cout, co, nco, acc);
#pragma unroll
      for (int k = 0; k < BLAZE_PG; ++k) {
        const int p = pg * BLAZE_PG + k;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (j >= nco) break;
          const float res = co + j < cin ? s_act[p][co + j] : 0.0f;
          s_act[p][co + j] = fmaxf(acc[k][j] + res, 0.0f);
        }
      }
    }
    __syncthreads();
  }
  const int cl = ch.cout[ch.nblk - 1];
  float* o = out + img * (size_t)P * cl;
  for (int i = threadIdx.x; i < P * cl; i += THREADS) o[i] = s_act[i / cl][i % cl];
}

// ---- heads (:118-148): classifier_8 / regressor_8 on the 16x16x88 map, classifier_16 / regressor_16 on 8x8x96,
// written anchor-major: anchor = cell * A + a (A = 2 resp. 6, the second grid after the first 512 anchors).
__global__ void __launch_bounds__(256)
blaze_heads_kernel(const float* __restrict__ f8, const float* __restrict__ f16, const float* __restrict__ w8,
                   const float* __restrict__ b8, const float* __restrict__ w16, const float* __restrict__ b16,
                   float* __restrict__ raw_boxes, float* __restrict__ raw_scores) {
  const size_t img = blockIdx.y;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  constexpr int N8 = 256 * 34, N16 = 64 * 102;
  if (idx >= N8 + N16) return;
  const bool first = idx < N8;
  const int j = first ? idx : idx - N8;
  const int nout = first ? 34 : 102, cin = first ? 88 : 96, na = first ? 2 : 6;
  const int cell = j / nout, o = j - cell * nout;
  const float* f = (first ? f8 + img * 256 * 88 : f16 + img * 64 * 96) + (size_t)cell * cin;
  const float* w = (first ? w8 : w16) + o;
  float acc = (first ? b8 : b16)[o];
  for (int c = 0; c < cin; ++c) acc = fmaf(f[c], w[c * nout], acc);
  const int anchor0 = (first ? 0 : 512) + cell * na;
  if (o < na) raw_scores[img * NUM_ANCHORS + anchor0 + o] = acc;
  else {
    const int r = o - na;                       // regressor channel = a * 16 + k
    raw_boxes[(img * NUM_ANCHORS + anchor0 + r / 16) * 16 + (r & 15)] = acc;
  }
}

// ---- _decode_boxes + clamp / sigmoid (:262-303): one thread per anchor -> [n][896][17]
__global__ void __launch_bounds__(256)
blaze_decode_kernel(const float* __restrict__ raw_boxes, const float* __restrict__ raw_scores,
                    const float* __restrict__ anchors, float* __restrict__ det, int n) {
  const size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (idx >= (size_t)n * NUM_ANCHORS) return;
  const int a = (int)(idx % NUM_ANCHORS);
  const float ax = anchors[a * 4], ay = anchors[a * 4 + 1], aw = anchors[a * 4 + 2], ah = anchors[a * 4 + 3];
  const float* r = raw_boxes + idx * 16;
  float* d = det + idx * 17;
  const float xc = __fadd_rn(__fmul_rn(__fdiv_rn(r[0], 128.0f), aw), ax);
  const float yc = __fadd_rn(__fmul_rn(__fdiv_rn(r[1], 128.0f), ah), ay);
  const float w = __fmul_rn(__fdiv_rn(r[2], 128.0f), aw);
  const float hh = __fmul_rn(__fdiv_rn(r[3], 128.0f), ah);
  d[0] = __fsub_rn(yc, __fdiv_rn(hh, 2.0f));
  d[1] = __fsub_rn(xc, __fdiv_rn(w, 2.0f));
  d[2] = __fadd_rn(yc, __fdiv_rn(hh, 2.0f));
  d[3] = __fadd_rn(xc, __fdiv_rn(w, 2.0f));
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    d[4 + 2 * k] = __fadd_rn(__fmul_rn(__fdiv_rn(r[4 + 2 * k], 128.0f), aw), ax);
    d[5 + 2 * k] = __fadd_rn(__fmul_rn(__fdiv_rn(r[5 + 2 * k], 128.0f), ah), ay);
  }
  const float s = fminf(fmaxf(raw_scores[idx], -100.0f), 100.0f);
  d[16] = 1.0f / (1.0f + expf(-s));
}

// ---- blending NMS on the device (blazeface.py:305-358) for predict_on_batch(apply_nms=True): one warp per tile.
// Candidates = anchors with score >= min_score (the mask of :262); they are ranked by score (ties: lower anchor index
// first), then greedily: the best remaining detection absorbs every remaining one with IoU > thr (itself included) into
// their score-weighted mean, the merged score being the mean score — exactly the reference's loop.  Up to 64
// candidates and 16 faces per tile are handled here; a tile with more sets its count to -1 and the mirror falls back to
// the host loop for that tile.
constexpr int NMS_MAX_CAND = 64, NMS_MAX_FACES = 16;
__global__ void __launch_bounds__(128)
blaze_nms_kernel(const float* __restrict__ det, const int* __restrict__ list_offsets, int n, float min_score, float iou_thr,
                 float* __restrict__ faces, int* __restrict__ counts) {
  __shared__ int s_idx[4][NMS_MAX_CAND];
  __shared__ float s_score[4][NMS_MAX_CAND];
  __shared__ float s_box[4][NMS_MAX_CAND][4];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x * 4 + w;
  if (tile >= n) return;
  // dense mode: the tile's 896 decoded anchors, masked by score;  list mode (`nms(list)`, blazeface.py:225-234): rows
  // list_offsets[tile] .. list_offsets[tile+1] of a packed [total,17] array, every one a candidate
  const float* d = det + (list_offsets ? (size_t)list_offsets[tile] * 17 : (size_t)tile * NUM_ANCHORS * 17);
  const int rows = list_offsets ? list_offsets[tile + 1] - list_offsets[tile] : NUM_ANCHORS;
  // 1. candidates in anchor order
  int ncand = 0;
  bool overflow = false;
  for (int a0 = 0; a0 < rows; a0 += 32) {
    const int a = a0 + lane;
    const float sc = a < rows ? d[a * 17 + 16] : 0.f;
    const bool keep = a < rows && (list_offsets != nullptr || sc >= min_score);
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    const int pos = ncand + __popc(m & ((1u << lane) - 1u));
    if (keep && pos < NMS_MAX_CAND) { s_idx[w][pos] = a; s_score[w][pos] = sc; }
    ncand += __popc(m);
  }
  if (ncand > NMS_MAX_CAND) overflow = true;
  __syncwarp();
  if (overflow) { if (lane == 0) counts[tile] = -1; return; }
  // 2. rank by score, descending (stable in the anchor index): rank[i] = #{j: s_j > s_i or (s_j == s_i and j < i)}
  int my_rank[2] = {-1, -1};
  for (int h = 0; h < 2; ++h) {
    const int i = lane + 32 * h;
    if (i < ncand) {
      const float si = s_score[w][i];
      int r = 0;
      for (int j = 0; j < ncand; ++j) { const float sj = s_score[w][j]; r += (sj > si || (sj == si && j < i)) ? 1 : 0; }
      my_rank[h] = r;
    }
  }
  int my_anchor[2]; float my_sc[2];
  for (int h = 0; h < 2; ++h) { const int i = lane + 32 * h; my_anchor[h] = i < ncand ? s_idx[w][i] : 0; my_sc[h] = i < ncand ? s_score[w][i] : 0.f; }
  __syncwarp();
  for (int h = 0; h < 2; ++h)
    if (my_rank[h] >= 0) { s_idx[w][my_rank[h]] = my_anchor[h]; s_score[w][my_rank[h]] = my_sc[h]; }
  __syncwarp();
  for (int h = 0; h < 2; ++h) {
    const int i = lane + 32 * h;
    if (i < ncand) {
      const float* b = d + s_idx[w][i] * 17;
      s_box[w][i][0] = b[0]; s_box[w][i][1] = b[1]; s_box[w][i][2] = b[2]; s_box[w][i][3] = b[3];
    }
  }
  __syncwarp();
  // 3. greedy blending; alive flags live in two 32-bit masks
  unsigned alive[2] = {ncand >= 32 ? 0xffffffffu : ((1u << ncand) - 1u), ncand > 32 ? ((ncand >= 64) ? 0xffffffffu : ((1u << (ncand - 32)) - 1u)) : 0u};
  int nfaces = 0;
  float* out = faces + (size_t)tile * NMS_MAX_FACES * 17;
  while ((alive[0] | alive[1]) != 0u) {
    const int first = alive[0] ? __ffs(alive[0]) - 1 : 32 + __ffs(alive[1]) - 1;     // sorted: lowest index = best score
    const float fy0 = s_box[w][first][0], fx0 = s_box[w][first][1], fy1 = s_box[w][first][2], fx1 = s_box[w][first][3];
    const float farea = (fy1 - fy0) * (fx1 - fx0);
    unsigned ov[2];
    for (int h = 0; h < 2; ++h) {
      const int i = lane + 32 * h;
      bool o = false;
      if (i < ncand && ((alive[h] >> lane) & 1u)) {
        const float ih = fmaxf(fminf(fy1, s_box[w][i][2]) - fmaxf(fy0, s_box[w][i][0]), 0.f);
        const float iw = fmaxf(fminf(fx1, s_box[w][i][3]) - fmaxf(fx0, s_box[w][i][1]), 0.f);
        const float inter = ih * iw;
        const float area = (s_box[w][i][2] - s_box[w][i][0]) * (s_box[w][i][3] - s_box[w][i][1]);
        o = inter / (farea + area - inter) > iou_thr;
      }
      ov[h] = __ballot_sync(0xffffffffu, o);
    }
    ov[first >> 5] |= 1u << (first & 31);            // a degenerate (zero-area) best box must still be consumed
    const int cnt = __popc(ov[0]) + __popc(ov[1]);
    if (nfaces < NMS_MAX_FACES) {
      if (cnt > 1) {
        // lane k < 17 accumulates coordinate k (k = 16: the score sum) over the overlapping set in rank order
        if (lane < 17) {
          float num = 0.f, tot = 0.f;
          for (int h = 0; h < 2; ++h)
            for (unsigned m = ov[h]; m; m &= m - 1) {
              const int i = 32 * h + __ffs(m) - 1;
              const float sc = s_score[w][i];
              tot += sc;
              if (lane < 16) num += d[s_idx[w][i] * 17 + lane] * sc;
            }
          out[nfaces * 17 + lane] = lane < 16 ? num / tot : tot / (float)cnt;
        }
      } else if (lane < 17) {
        out[nfaces * 17 + lane] = d[s_idx[w][first] * 17 + lane];
      }
    }
    ++nfaces;
    alive[0] &= ~ov[0];
    alive[1] &= ~ov[1];
  }
  if (lane == 0) counts[tile] = nfaces <= NMS_MAX_FACES ? nfaces : -1;
}


// ---- per-FRAME detections for the reference's FaceExtractor (helpers_face_extract_1.py:87-124): the dense detections
// of a frame's T tiles (T = 3 overlapping square windows of a landscape frame, 1 of a portrait one, :185-205) are masked
// (score >= min_score, blazeface.py:262), mapped back to frame coordinates — `_resize_detections` (:207-233: tile
// coordinate * 128 * split/128) then `_untile_detections` (:235-272: + the window's x / y offset) — merged by the blending
// NMS of blazeface.py:305-358 over the whole frame, and expanded by `_add_margin_to_detections` (:274-294: 20 % of the
// box height, twice that above, clamped to the frame) into the integer crop rectangle `_crop_faces` cuts (:296-312).
// One warp per frame; same candidate / face limits and the same fall-back contract (count = -1) as blaze_nms_kernel.
struct FrameGeom {
  int tiles;            // T
  float scale;          // split_size / 128 as fp32 (what `detection * target - 0) * scale` multiplies by)
  int x_step;           // window t starts at (t * x_step, 0): num_v = 1 in the reference
  int frame_w, frame_h;
};
__global__ void __launch_bounds__(128)
blaze_frame_nms_kernel(const float* __restrict__ det, int n_frames, FrameGeom g, float min_score, float iou_thr, float margin,
                       float* __restrict__ faces, int* __restrict__ boxes, int* __restrict__ counts) {
  __shared__ int s_idx[4][NMS_MAX_CAND];            // tile * 896 + anchor
  __shared__ float s_score[4][NMS_MAX_CAND];
  __shared__ float s_box[4][NMS_MAX_CAND][4];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int frame = blockIdx.x * 4 + w;
  if (frame >= n_frames) return;
  const float* d = det + (size_t)frame * g.tiles * NUM_ANCHORS * 17;
  // coordinate k of detection (tile t, anchor a) in frame coordinates, in the reference's fp32 order of operations
  auto coord = [&](int ta, int k) {
    const int t = ta / NUM_ANCHORS;
    const float v = __fmul_rn(__fsub_rn(__fmul_rn(d[(size_t)ta * 17 + k], 128.0f), 0.0f), g.scale);
    const bool is_x = k < 4 ? (k & 1) : !(k & 1);           // box: ymin xmin ymax xmax; keypoints: x y pairs
    return __fadd_rn(v, is_x ? (float)(t * g.x_step) : 0.0f);     // num_v = 1: every window starts at y = 0 (:248-268)
  };
  int ncand = 0;
  for (int a0 = 0; a0 < g.tiles * NUM_ANCHORS; a0 += 32) {
    const int a = a0 + lane;
    const float sc = a < g.tiles * NUM_ANCHORS ? d[(size_t)a * 17 + 16] : -1.0f;
    const bool keep = sc >= min_score;
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    const int pos = ncand + __popc(m & ((1u << lane) - 1u));
    if (keep && pos < NMS_MAX_CAND) { s_idx[w][pos] = a; s_score[w][pos] = sc; }
    ncand += __popc(m);
  }
  __syncwarp();
  if (ncand > NMS_MAX_CAND) { if (lane == 0) counts[frame] = -1; return; }
  int my_rank[2] = {-1, -1};
  for (int h = 0; h < 2; ++h) {
    const int i = lane + 32 * h;
    if (i < ncand) {
      const float si = s_score[w][i];
      int r = 0;
      for (int j = 0; j < ncand; ++j) { const float sj = s_score[w][j]; r += (sj > si || (sj == si && j < i)) ? 1 : 0; }
      my_rank[h] = r;
    }
  }
  int my_anchor[2]; float my_sc[2];
  for (int h = 0; h < 2; ++h) { const int i = lane + 32 * h; my_anchor[h] = i < ncand ? s_idx[w][i] : 0; my_sc[h] = i < ncand ? s_score[w][i] : 0.f; }
  __syncwarp();
  for (int h = 0; h < 2; ++h)
    if (my_rank[h] >= 0) { s_idx[w][my_rank[h]] = my_anchor[h]; s_score[w][my_rank[h]] = my_sc[h]; }
  __syncwarp();
  for (int h = 0; h < 2; ++h) {
    const int i = lane + 32 * h;
    if (i < ncand)
      for (int k = 0; k < 4; ++k) s_box[w][i][k] = coord(s_idx[w][i], k);
  }
  __syncwarp();
  unsigned alive[2] = {ncand >= 32 ? 0xffffffffu : ((1u << ncand) - 1u), ncand > 32 ? ((ncand >= 64) ? 0xffffffffu : ((1u << (ncand - 32)) - 1u)) : 0u};
  int nfaces = 0;
  float* out = faces + (size_t)frame * NMS_MAX_FACES * 17;
  int* bout = boxes + (size_t)frame * NMS_MAX_FACES * 4;
  while ((alive[0] | alive[1]) != 0u) {
    const int first = alive[0] ? __ffs(alive[0]) - 1 : 32 + __ffs(alive[1]) - 1;
    const float fy0 = s_box[w][first][0], fx0 = s_box[w][first][1], fy1 = s_box[w][first][2], fx1 = s_box[w][first][3];
    const float farea = (fy1 - fy0) * (fx1 - fx0);
    unsigned ov[2];
    for (int h = 0; h < 2; ++h) {
      const int i = lane + 32 * h;
      bool o = false;
      if (i < ncand && ((alive[h] >> lane) & 1u)) {
        const float ih = fmaxf(fminf(fy1, s_box[w][i][2]) - fmaxf(fy0, s_box[w][i][0]), 0.f);
        const float iw = fmaxf(fminf(fx1, s_box[w][i][3]) - fmaxf(fx0, s_box[w][i][1]), 0.f);
        const float inter = ih * iw;
        const float area = (s_box[w][i][2] - s_box[w][i][0]) * (s_box[w][i][3] - s_box[w][i][1]);
        o = inter / (farea + area - inter) > iou_thr;
      }
      ov[h] = __ballot_sync(0xffffffffu, o);
    }
    ov[first >> 5] |= 1u << (first & 31);
    const int cnt = __popc(ov[0]) + __popc(ov[1]);
    if (nfaces < NMS_MAX_FACES) {
      float val = 0.f;
      if (lane < 17) {
        if (cnt > 1) {
          float num = 0.f, tot = 0.f;
          for (int h = 0; h < 2; ++h)
            for (unsigned m = ov[h]; m; m &= m - 1) {
              const int i = 32 * h + __ffs(m) - 1;
              const float sc = s_score[w][i];
              tot += sc;
              if (lane < 16) num += coord(s_idx[w][i], lane) * sc;
            }
          val = lane < 16 ? num / tot : tot / (float)cnt;
        } else {
          val = lane < 16 ? coord(s_idx[w][first], lane) : s_score[w][first];
        }
        out[nfaces * 17 + lane] = val;
      }
      // margin + integer crop rectangle (ymin, xmin, ymax, xmax): lanes 0..3 hold the merged box
      const float ymin = __shfl_sync(0xffffffffu, val, 0), xmin = __shfl_sync(0xffffffffu, val, 1);
      const float ymax = __shfl_sync(0xffffffffu, val, 2), xmax = __shfl_sync(0xffffffffu, val, 3);
      if (lane == 0) {
        const float off = rintf(__fmul_rn(margin, __fsub_rn(ymax, ymin)));          // torch.round = half to even
        bout[nfaces * 4 + 0] = (int)fmaxf(__fsub_rn(ymin, __fmul_rn(off, 2.0f)), 0.0f);
        bout[nfaces * 4 + 1] = (int)fmaxf(__fsub_rn(xmin, off), 0.0f);
        bout[nfaces * 4 + 2] = (int)fminf(__fadd_rn(ymax, off), (float)g.frame_h);
        bout[nfaces * 4 + 3] = (int)fminf(__fadd_rn(xmax, off), (float)g.frame_w);
      }
    }
    ++nfaces;
    alive[0] &= ~ov[0];
    alive[1] &= ~ov[1];
  }
  if (lane == 0) counts[frame] = nfaces <= NMS_MAX_FACES ? nfaces : -1;
}

template <typename T>
int balloc(ff_blazeface* h, T** p, size_t count) {
  void* q = nullptr;
  BZ_CUDA(h, cudaMalloc(&q, std::max<size_t>(count, 1) * sizeof(T)));
  h->allocs.push_back(q);
  *p = reinterpret_cast<T*>(q);
  return FF_OK;
}
int bupload(ff_blazeface* h, float** p, const std::vector<float>& v) {
  int rc = balloc(h, p, v.size());
  if (rc) return rc;
  BZ_CUDA(h, cudaMemcpy(*p, v.data(), v.size() * sizeof(float), cudaMemcpyHostToDevice));
  return FF_OK;
}
const std::vector<float>* bget(ff_blazeface* h, const std::string& key, std::initializer_list<int64_t> shape) {
  auto it = h->host_w.find(key);
  if (it == h->host_w.end()) { bfail(h, FF_ERR_STATE, "missing weight '%s'", key.c_str()); return nullptr; }
  const auto& s = h->host_shape[key];
  if (s.size() != shape.size() || !std::equal(s.begin(), s.end(), shape.begin())) {
    bfail(h, FF_ERR_SHAPE, "weight '%s' has the wrong shape", key.c_str());
    return nullptr;
  }
  return &it->second;
}

int bfinalize(ff_blazeface* h) {
  auto bad = [&]() { return h->err.find("shape") != std::string::npos ? FF_ERR_SHAPE : FF_ERR_STATE; };
  int rc;
  {
    const auto* w = bget(h, "backbone1.0.weight", {24, 3, 5, 5});
    const auto* b = bget(h, "backbone1.0.bias", {24});
    if (!w || !b) return bad();
    std::vector<float> wt(5 * 5 * 3 * 24);
    for (int co = 0; co < 24; ++co)
      for (int c = 0; c < 3; ++c)
        for (int t = 0; t < 25; ++t) wt[(t * 3 + c) * 24 + co] = (*w)[(co * 3 + c) * 25 + t];
    if ((rc = bupload(h, &h->stem_w, wt))) return rc;
    if ((rc = bupload(h, &h->stem_b, *b))) return rc;
  }
  for (int i = 0; i < 16; ++i) {
    const BlockPlan& p = kBlocks[i];
    const std::string key = i < 11 ? "backbone1." + std::to_string(i + 2) : "backbone2." + std::to_string(i - 11);
    const auto* dw = bget(h, key + ".convs.0.weight", {p.cin, 1, 3, 3});
    const auto* db = bget(h, key + ".convs.0.bias", {p.cin});
    const auto* pw = bget(h, key + ".convs.1.weight", {p.cout, p.cin, 1, 1});
    const auto* pb = bget(h, key + ".convs.1.bias", {p.cout});
    if (!dw || !db || !pw || !pb) return bad();
    std::vector<float> dwt(9 * p.cin), pwt((size_t)p.cin * p.cout);
    for (int c = 0; c < p.cin; ++c)
      for (int t = 0; t < 9; ++t) dwt[t * p.cin + c] = (*dw)[c * 9 + t];
    for (int co = 0; co < p.cout; ++co)
      for (int c = 0; c < p.cin; ++c) pwt[(size_t)c * p.cout + co] = (*pw)[(size_t)co * p.cin + c];
    if ((rc = bupload(h, &h->dw_w[i], dwt))) return rc;
    if ((rc = bupload(h, &h->dw_b[i], *db))) return rc;
    if ((rc = bupload(h, &h->pw_w[i], pwt))) return rc;
    if ((rc = bupload(h, &h->pw_b[i], *pb))) return rc;
  }
  for (int g = 0; g < 2; ++g) {
    const int cin = g == 0 ? 88 : 96, na = g == 0 ? 2 : 6, nout = na * 17;
    const std::string sfx = g == 0 ? "_8" : "_16";
    const auto* cw = bget(h, "classifier" + sfx + ".weight", {na, cin, 1, 1});
    const auto* cb = bget(h, "classifier" + sfx + ".bias", {na});
    const auto* rw = bget(h, "regressor" + sfx + ".weight", {na * 16, cin, 1, 1});
    const auto* rb = bget(h, "regressor" + sfx + ".bias", {na * 16});
    if (!cw || !cb || !rw || !rb) return bad();
    std::vector<float> wt((size_t)cin * nout), bt(nout);
    for (int o = 0; o < nout; ++o) {
      const bool cls = o < na;
      bt[o] = cls ? (*cb)[o] : (*rb)[o - na];
      for (int c = 0; c < cin; ++c) wt[(size_t)c * nout + o] = cls ? (*cw)[(size_t)o * cin + c] : (*rw)[(size_t)(o - na) * cin + c];
    }
    if ((rc = bupload(h, g == 0 ? &h->head8_w : &h->head16_w, wt))) return rc;
    if ((rc = bupload(h, g == 0 ? &h->head8_b : &h->head16_b, bt))) return rc;
  }
  {
    const auto* a = bget(h, "anchors", {NUM_ANCHORS, 4});
    if (!a) return bad();
    if ((rc = bupload(h, &h->anchors, *a))) return rc;
  }
  h->host_w.clear();
  h->host_shape.clear();
  h->finalized = true;
  return FF_OK;
}

int bforward(ff_blazeface* h, const uint8_t* tiles, int n, float* det, cudaStream_t st) {
  {
    blaze_stem_kernel<<<dim3(4, 4, n), 256, 0, st>>>(tiles, h->stem_w, h->stem_b, h->act_a, n);
    BZ_CUDA(h, cudaGetLastError());
    ++h->launches;
  }
  float* cur = h->act_a;
  float* nxt = h->act_b;
  auto make_chain = [&](int first, int count) {
    BlazeChain ch;
    ch.nblk = count;
    for (int j = 0; j < count; ++j) {
      ch.cin[j] = kBlocks[first + j].cin; ch.cout[j] = kBlocks[first + j].cout;
      ch.dw_w[j] = h->dw_w[first + j]; ch.dw_b[j] = h->dw_b[first + j];
      ch.pw_w[j] = h->pw_w[first + j]; ch.pw_b[j] = h->pw_b[first + j];
    }
    return ch;
  };
  for (int i = 0; i < 16; ++i) {
    if (i == 6) {             // blocks 6..10: the whole 16x16 stage -> feat8
      constexpr int SM = (256 * 89 + 88 * 260 + 4) * 4;
      BZ_CUDA(h, ffh::ensure_dyn_smem(reinterpret_cast<const void*>(blaze_chain_kernel<16, 88, 512>), SM));
      blaze_chain_kernel<16, 88, 512><<<n, 512, SM, st>>>(cur, h->feat8, make_chain(6, 5));
      BZ_CUDA(h, cudaGetLastError());
      ++h->launches;
      cur = h->feat8;
      i = 10;
      continue;
    }
    if (i == 12) {            // blocks 12..15: the 8x8 stage after the stride-2 block 11
      constexpr int SM = (64 * 97 + 96 * 68 + 4) * 4;
      BZ_CUDA(h, ffh::ensure_dyn_smem(reinterpret_cast<const void*>(blaze_chain_kernel<8, 96, 256>), SM));
      blaze_chain_kernel<8, 96, 256><<<n, 256, SM, st>>>(cur, nxt, make_chain(12, 4));
      BZ_CUDA(h, cudaGetLastError());
      ++h->launches;
      cur = nxt;
      i = 15;
      continue;
    }
    const BlockPlan& p = kBlocks[i];
    const int hw_out = p.hw_in / p.stride;
    float* dst = (i == 10) ? h->feat8 : nxt;          // backbone1 output (16x16x88) feeds both backbone2 and the heads
    const int pix = blaze_pix(hw_out);
    dim3 grid((hw_out * hw_out + pix - 1) / pix, n);
    const size_t smem = (size_t)p.cin * (pix + 4) * sizeof(float);
    if (pix == 256) blaze_block_kernel<256><<<grid, 256, smem, st>>>(cur, dst, h->dw_w[i], h->dw_b[i], h->pw_w[i], h->pw_b[i], p.cin, p.cout, p.stride, p.hw_in);
    else if (pix == 128) blaze_block_kernel<128><<<grid, 256, smem, st>>>(cur, dst, h->dw_w[i], h->dw_b[i], h->pw_w[i], h->pw_b[i], p.cin, p.cout, p.stride, p.hw_in);
    else blaze_block_kernel<64><<<grid, 256, smem, st>>>(cur, dst, h->dw_w[i], h->dw_b[i], h->pw_w[i], h->pw_b[i], p.cin, p.cout, p.stride, p.hw_in);
    BZ_CUDA(h, cudaGetLastError());
    ++h->launches;
    if (i == 10) cur = h->feat8;
    else { cur = dst; nxt = (dst == h->act_a) ? h->act_b : h->act_a; }
  }
  {
    dim3 grid((256 * 34 + 64 * 102 + 255) / 256, n);
    blaze_heads_kernel<<<grid, 256, 0, st>>>(h->feat8, cur, h->head8_w, h->head8_b, h->head16_w, h->head16_b, h->raw_boxes, h->raw_scores);
    BZ_CUDA(h, cudaGetLastError());
    const size_t total = (size_t)n * NUM_ANCHORS;
    blaze_decode_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(h->raw_boxes, h->raw_scores, h->anchors, det, n);
    BZ_CUDA(h, cudaGetLastError());
    h->launches += 2;
  }
  return FF_OK;
}

}  // namespace

extern "C" {

int ff_blazeface_create(ff_blazeface_t** out, int device, int max_tiles) {
  if (!out || max_tiles <= 0) return bfail(nullptr, FF_ERR_BAD_ARG, "ff_blazeface_create: bad arguments");
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev <= 0)
    return bfail(nullptr, FF_ERR_CUDA, "no CUDA device (%s): libfacfake has no CPU fallback", cudaGetErrorString(e));
  if (device < 0 || device >= ndev) return bfail(nullptr, FF_ERR_BAD_ARG, "device %d out of range (%d devices)", device, ndev);
  ffh::DeviceGuard guard(device);
  if (guard.status != cudaSuccess) return bfail(nullptr, FF_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(guard.status));
  ff_blazeface* h = new ff_blazeface();
  h->device = device;
  h->cap = max_tiles;
  int rc = FF_OK;
  do {
    if ((rc = balloc(h, &h->act_a, (size_t)h->cap * ACT_ELEMS))) break;
    if ((rc = balloc(h, &h->act_b, (size_t)h->cap * ACT_ELEMS))) break;
    if ((rc = balloc(h, &h->feat8, (size_t)h->cap * 16 * 16 * 88))) break;
    if ((rc = balloc(h, &h->raw_boxes, (size_t)h->cap * NUM_ANCHORS * 16))) break;
    if ((rc = balloc(h, &h->raw_scores, (size_t)h->cap * NUM_ANCHORS))) break;
  } while (0);
  if (rc != FF_OK) {
    g_blaze_create_error = h->err;
    ff_blazeface_destroy(h);
    return rc;
  }
  *out = h;
  return FF_OK;
}

void ff_blazeface_destroy(ff_blazeface_t* h) {
  if (!h) return;
  {
    ffh::DeviceGuard guard(h->device);
    cudaDeviceSynchronize();
    for (void* p : h->allocs) cudaFree(p);
    if (h->tile_desc) cudaFree(h->tile_desc);
  }
  delete h;
}

const char* ff_blazeface_last_error(const ff_blazeface_t* h) { return h ? h->err.c_str() : g_blaze_create_error.c_str(); }

int ff_blazeface_load_weight(ff_blazeface_t* h, const char* key, const float* host_fp32, const int64_t* shape, int ndim) {
  if (!h || !key || !host_fp32 || ndim < 0 || ndim > 8 || (ndim > 0 && !shape)) return bfail(h, FF_ERR_BAD_ARG, "ff_blazeface_load_weight: bad arguments");
  std::lock_guard<std::mutex> lk(h->mu);
  if (h->finalized) return bfail(h, FF_ERR_STATE, "weights are already finalized");
  size_t count = 1;
  std::vector<int64_t> s(shape, shape + ndim);
  for (int64_t d : s) {
    if (d < 0) return bfail(h, FF_ERR_SHAPE, "negative dimension in '%s'", key);
    count *= (size_t)d;
  }
  h->host_w[key].assign(host_fp32, host_fp32 + count);
  h->host_shape[key] = s;
  return FF_OK;
}

int ff_blazeface_finalize(ff_blazeface_t* h) {
  if (!h) return FF_ERR_BAD_ARG;
  std::lock_guard<std::mutex> lk(h->mu);
  if (h->finalized) return FF_OK;
  ffh::DeviceGuard guard(h->device);
  return bfinalize(h);
}

int ff_blazeface_predict(ff_blazeface_t* h, const uint8_t* tiles, int n, float* detections, float* raw_boxes,
                         float* raw_scores, void* stream) {
  if (!h || n < 0 || (n > 0 && (!tiles || !detections))) return bfail(h, FF_ERR_BAD_ARG, "ff_blazeface_predict: bad arguments");
  std::lock_guard<std::mutex> lk(h->mu);
  if (!h->finalized) return bfail(h, FF_ERR_STATE, "ff_blazeface_finalize() has not been called");
  ffh::DeviceGuard guard(h->device);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  for (int s0 = 0; s0 < n; s0 += h->cap) {
    const int ns = std::min(h->cap, n - s0);
    int rc = bforward(h, tiles + (size_t)s0 * 128 * 128 * 3, ns, detections + (size_t)s0 * NUM_ANCHORS * 17, st);
    if (rc) return rc;
    if (raw_boxes) BZ_CUDA(h, cudaMemcpyAsync(raw_boxes + (size_t)s0 * NUM_ANCHORS * 16, h->raw_boxes, (size_t)ns * NUM_ANCHORS * 16 * sizeof(float), cudaMemcpyDeviceToDevice, st));
    if (raw_scores) BZ_CUDA(h, cudaMemcpyAsync(raw_scores + (size_t)s0 * NUM_ANCHORS, h->raw_scores, (size_t)ns * NUM_ANCHORS * sizeof(float), cudaMemcpyDeviceToDevice, st));
  }
  return FF_OK;
}

int ff_blazeface_nms(ff_blazeface_t* h, const float* detections, int n, float min_score, float iou_threshold, float* faces,
                     int32_t* counts, void* stream) {
  if (!h || n < 0 || (n > 0 && (!detections || !faces || !counts))) return bfail(h, FF_ERR_BAD_ARG, "ff_blazeface_nms: bad arguments");
  std::lock_guard<std::mutex> lk(h->mu);
  ffh::DeviceGuard guard(h->device);
  if (n == 0) return FF_OK;
  blaze_nms_kernel<<<(n + 3) / 4, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(detections, nullptr, n, min_score, iou_threshold, faces, counts);
  BZ_CUDA(h, cudaGetLastError());
  ++h->launches;
  return FF_OK;
}

int ff_blazeface_nms_lists(ff_blazeface_t* h, const float* detections, const int32_t* offsets, int n, float iou_threshold,
                           float* faces, int32_t* counts, void* stream) {
  if (!h || n < 0 || (n > 0 && (!detections || !offsets || !faces || !counts)))
    return bfail(h, FF_ERR_BAD_ARG, "ff_blazeface_nms_lists: bad arguments");
  std::lock_guard<std::mutex> lk(h->mu);
  ffh::DeviceGuard guard(h->device);
  if (n == 0) return FF_OK;
  blaze_nms_kernel<<<(n + 3) / 4, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(detections, offsets, n, 0.f, iou_threshold, faces, counts);
  BZ_CUDA(h, cudaGetLastError());
  ++h->launches;
  return FF_OK;
}

int ff_blazeface_tile_frames(ff_blazeface_t* h, const uint8_t* frames, int n_frames, int frame_h, int frame_w, uint8_t* tiles,
                             void* stream) {
  if (!h || n_frames < 0 || frame_h <= 0 || frame_w <= 0 || (n_frames > 0 && (!frames || !tiles)))
    return bfail(h, FF_ERR_BAD_ARG, "ff_blazeface_tile_frames: bad arguments");
  std::lock_guard<std::mutex> lk(h->mu);
  ffh::DeviceGuard guard(h->device);
  if (n_frames == 0) return FF_OK;
  // helpers_face_extract_1.py:185-205: square windows of side min(H, W); three of them, (W - side) / 2 apart, for a
  // landscape frame, one for a portrait frame; each resized to 128 x 128 with cv2.INTER_AREA
  const int split = std::min(frame_h, frame_w), x_step = (frame_w - split) / 2, T = frame_w > frame_h ? 3 : 1;
  const int n = n_frames * T;
  std::vector<ff::CropDesc> d(n);
  for (int f = 0; f < n_frames; ++f)
    for (int t = 0; t < T; ++t)
      d[f * T + t] = ff::make_crop_desc(frames + ((size_t)f * frame_h * frame_w + (size_t)t * x_step) * 3, split, split, frame_w * 3, 128);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if ((size_t)n > h->tile_desc_cap) {
    if (h->tile_desc) cudaFree(h->tile_desc);
    h->tile_desc = nullptr; h->tile_desc_cap = 0;
    BZ_CUDA(h, cudaMalloc(&h->tile_desc, sizeof(ff::CropDesc) * n));
    h->tile_desc_cap = n;
  }
  BZ_CUDA(h, cudaMemcpyAsync(h->tile_desc, d.data(), sizeof(ff::CropDesc) * n, cudaMemcpyHostToDevice, st));
  ff::preprocess_kernel<128><<<dim3(128 / 4, n), 256, 0, st>>>(reinterpret_cast<const ff::CropDesc*>(h->tile_desc), n, 0, tiles, nullptr);
  BZ_CUDA(h, cudaGetLastError());
  ++h->launches;
  return FF_OK;
}

int ff_blazeface_frame_faces(ff_blazeface_t* h, const float* detections, int n_frames, int frame_h, int frame_w, float min_score,
                             float iou_threshold, float margin, float* faces, int32_t* boxes, int32_t* counts, void* stream) {
  if (!h || n_frames < 0 || frame_h <= 0 || frame_w <= 0 || (n_frames > 0 && (!detections || !faces || !boxes || !counts)))
    return bfail(h, FF_ERR_BAD_ARG, "ff_blazeface_frame_faces: bad arguments");
  std::lock_guard<std::mutex> lk(h->mu);
  ffh::DeviceGuard guard(h->device);
  if (n_frames == 0) return FF_OK;
  FrameGeom g;
  const int split = std::min(frame_h, frame_w);
  g.tiles = frame_w > frame_h ? 3 : 1;
  g.scale = (float)((double)split / 128.0);
  g.x_step = (frame_w - split) / 2;
  g.frame_w = frame_w; g.frame_h = frame_h;
  blaze_frame_nms_kernel<<<(n_frames + 3) / 4, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(detections, n_frames, g, min_score, iou_threshold,
                                                                                               margin, faces, boxes, counts);
  BZ_CUDA(h, cudaGetLastError());
  ++h->launches;
  return FF_OK;
}

int64_t ff_blazeface_launch_count(const ff_blazeface_t* h) { return h ? h->launches : 0; }

}  // extern "C"
