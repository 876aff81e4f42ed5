// ff_s3d.cuh — S3D clip classifier on the GPU (SURVEY.md §8f-2;
// /root/reference/sx_exp_deepfakedetect-master/S3D/model.py:6-342; both SRM_net == 'no' and the SRM high-pass front-end
// SRM_net == 'yes', SRM/HPF.py:11-37).
//
// Activations are bf16 [clip][frame][h][w][channel] (NDHWC) with the TRUE channel count of each tensor.  Every
// convolution is one launch of rvk_conv2_kernel (ff_rvk.cuh: persistent tcgen05 implicit GEMM, TMA-store epilogue):
//   * 1x1x1 (BasicConv3d)        "flat": one GEMM row per voxel of the pass
//   * (1,3,3) spatial conv       the ResNet 3x3 path with clip*frame as the image index
//   * (k,1,1) temporal conv      4-D map (channel, h*w, frame, clip): the k taps shift the frame coordinate, the TMA
//                                zero-fill outside [0,T) is the temporal padding, the stem's temporal stride is the
//                                map's elementStride on that axis
//   * the stem's (1,7,7)/2 conv  rvk_stem_kernel (no im2col, ff_rvk.cuh) over all frames
// Channel counts that are not multiples of 64 (16, 24, 48, 112, 208, 480, 528, ...) cost nothing extra in memory: on
// the K side the TMA zero-fills the tail of the last 64-channel block (filters are zero-padded), on the N side each
// conv stores through a tensor-map VIEW of its slice of the Inception concat, so the hardware clips the padded
// columns and the four branches write the concat in place.  Eval-mode BatchNorm3d (eps 1e-3) + ReLU are folded into
// the producing launch.  Max-pools and the head (avg-pool (2,7,7), 1x1x1 fc, temporal mean) are small CUDA-core kernels.
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/facfake.h"
#include "ff_host.h"
#include "ff_rvk.cuh"
#include "ff_small.cuh"

using namespace ff;
using ffh::bf16;
using ffh::ilog2;
using ffh::launch_k;
using ffh::to_bf16;

namespace {

// name -> (cin, branch0, branch1 mid/out, branch2 mid/out, branch3)   model.py:84-342
struct S3dMixedPlan { int base_idx; int cin, b0, m1, o1, m2, o2, b3; };
const S3dMixedPlan kS3dMixed[9] = {
    {5, 192, 64, 96, 128, 16, 32, 32},    {6, 256, 128, 128, 192, 32, 96, 64},  {8, 480, 192, 96, 208, 16, 48, 64},
    {9, 512, 160, 112, 224, 24, 64, 64},  {10, 512, 128, 128, 256, 24, 64, 64}, {11, 512, 112, 144, 288, 32, 64, 64},
    {12, 528, 256, 160, 320, 32, 128, 128}, {14, 832, 256, 160, 320, 32, 128, 128}, {15, 832, 384, 192, 384, 48, 128, 128},
};
constexpr float S3D_BN_EPS = 1e-3f;

}  // namespace

struct ff_s3d {
  int device = 0, cap = 0, frames = 0, num_class = 1, num_sms = 148;
  int t1 = 0, t2 = 0, t3 = 0;                    // frames after the stem / after Mixed_3's pool / after Mixed_4's pool
  bool finalized = false;
  std::string err;
  std::mutex mu;
  std::map<std::string, std::vector<float>> host_w;
  std::map<std::string, std::vector<int64_t>> host_shape;
  std::vector<void*> allocs;
  enum { OP_CONV = 0, OP_POOL = 1 };
  enum { FLAT = 0, SPATIAL = 1, TEMPORAL = 2 };
  struct Op {
    int kind = OP_CONV;
    std::string name;
    // conv
    int mode = FLAT, cin = 0, cout = 0, cout_pad = 0, bn = 64, taps = 1, stride = 1;
    int hw = 0, t_in = 0, t_out = 0;             // output spatial size, frames per clip before / after
    int bw = 128, bh = 1, bi = 1;
    bf16* w = nullptr;
    float *scale = nullptr, *shift = nullptr;
    CUtensorMap tmA, tmB, tmO;
    // pool
    const bf16* pin = nullptr;
    bf16* pout = nullptr;
    int c = 0, hw_in = 0, kt = 1, ks = 1, st = 1, ss = 1, pt = 0, ps = 0;
    int tap_after = -1;                          // base.N index whose output this op completes (debug taps)
    const bf16* tap_ptr = nullptr;
    int tap_c = 0;
  };
  std::vector<Op> ops;
  bf16* buf[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};   // 0/1 = block in/out, 2..4 = branch temporaries
  bf16* x4 = nullptr;                            // bf16 NHWC4 frames
  bf16* stem_w = nullptr;
  float stem_scale[64], stem_shift[64];
  CUtensorMap tm_x4;
  float *fc_w = nullptr, *fc_b = nullptr;
  const bf16* final_feat = nullptr;              // [clip][t3][7][7][1024]
  // SRM_net == 'yes' (model.py:38-39, SRM/HPF.py:11-37): x4 holds fp16, hpf = Conv3d(3,30,(1,5,5)) output [frames][224][224][32] bf16
  int srm = 0;
  bf16* hpf = nullptr;
  bf16* hpf_w = nullptr;                         // pair-expanded 5x5 filter [5 kh][64 (p,co)][8 px x 4 ch] fp16 bits
  CUtensorMap tm_x4_hpf;
  int64_t launches = 0;
};

namespace {

std::string g_s3d_create_error;

int sfail(ff_s3d* h, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (h) h->err = buf; else g_s3d_create_error = buf;
  return code;
}
#define S3_CUDA(h, call)                                                                                        \
  do {                                                                                                          \
    cudaError_t e_ = (call);                                                                                    \
    if (e_ != cudaSuccess) return sfail(h, FF_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_));        \
  } while (0)

template <typename T>
int salloc(ff_s3d* h, T** p, size_t count) {
  void* q = nullptr;
  S3_CUDA(h, cudaMalloc(&q, std::max<size_t>(count, 1) * sizeof(T)));
  h->allocs.push_back(q);
  *p = reinterpret_cast<T*>(q);
  return FF_OK;
}
template <typename T>
int supload(ff_s3d* h, T** p, const std::vector<T>& v) {
  int rc = salloc(h, p, v.size());
  if (rc) return rc;
  S3_CUDA(h, cudaMemcpy(*p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  return FF_OK;
}
const std::vector<float>* sget(ff_s3d* h, const std::string& key, std::initializer_list<int64_t> shape) {
  auto it = h->host_w.find(key);
  if (it == h->host_w.end()) { sfail(h, FF_ERR_STATE, "missing weight '%s'", key.c_str()); return nullptr; }
  const auto& s = h->host_shape[key];
  if (s.size() != shape.size() || !std::equal(s.begin(), s.end(), shape.begin())) {
    sfail(h, FF_ERR_SHAPE, "weight '%s' has the wrong shape", key.c_str());
    return nullptr;
  }
  return &it->second;
}

// 4-D bf16 tensor map (channel, w, h, n) over a VIEW: `c` valid channels out of rows of `pitch_c` elements.
int s3d_tmap(ff_s3d* h, CUtensorMap* m, const void* base, int c, int pitch_c, long long w, int hh, int n, int bw, int bh,
             int bi, int es_w = 1, int es_h = 1) {
  cuuint64_t dims[4] = {(cuuint64_t)c, (cuuint64_t)w, (cuuint64_t)hh, (cuuint64_t)n};
  cuuint64_t strides[3] = {(cuuint64_t)pitch_c * 2, (cuuint64_t)w * pitch_c * 2, (cuuint64_t)hh * w * pitch_c * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)(bw * es_w), (cuuint32_t)(bh * es_h), (cuuint32_t)bi};
  cuuint32_t estr[4] = {1, (cuuint32_t)es_w, (cuuint32_t)es_h, 1};
  CUresult r = ffh::encode_tiled()(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return sfail(h, FF_ERR_CUDA, "cuTensorMapEncodeTiled(C%d/%d W%lld H%d N%d box %dx%dx%d) failed: %d", c, pitch_c, w, hh, n, bw, bh, bi, (int)r);
  return FF_OK;
}

int s3d_fold_bn(ff_s3d* h, const std::string& bn, int c, int c_pad, std::vector<float>* scale, std::vector<float>* shift) {
  const auto* g = sget(h, bn + ".weight", {c});
  const auto* be = sget(h, bn + ".bias", {c});
  const auto* mu = sget(h, bn + ".running_mean", {c});
  const auto* var = sget(h, bn + ".running_var", {c});
  if (!g || !be || !mu || !var) return h->err.find("shape") != std::string::npos ? FF_ERR_SHAPE : FF_ERR_STATE;
  scale->assign(c_pad, 0.0f);
  shift->assign(c_pad, 0.0f);
  for (int o = 0; o < c; ++o) {
    const float s = (*g)[o] / std::sqrt((*var)[o] + S3D_BN_EPS);
    (*scale)[o] = s;
    (*shift)[o] = (*be)[o] - (*mu)[o] * s;
  }
  return FF_OK;
}

// One convolution + BN + ReLU.  mode FLAT: 1x1x1; SPATIAL: (1,3,3) pad 1; TEMPORAL: (k,1,1), k = 3 (stride 1) or 7
// (stride 2).  `out` / `out_pitch` / `out_off` describe the destination tensor and the channel slice written.
int s3d_add_conv(ff_s3d* h, const std::string& conv_key, const std::string& bn_key, int mode, int cin, int cout, int k, int stride,
                 int hw, int t_in, const bf16* in, bf16* out, int out_pitch, int out_off) {
  ff_s3d::Op op;
  op.kind = ff_s3d::OP_CONV;
  op.name = conv_key;
  op.mode = mode; op.cin = cin; op.cout = cout; op.stride = stride; op.hw = hw;
  op.t_in = t_in;
  op.t_out = (mode == ff_s3d::TEMPORAL && stride == 2) ? (t_in - 1) / 2 + 1 : t_in;
  op.taps = mode == ff_s3d::FLAT ? 1 : (mode == ff_s3d::SPATIAL ? 9 : k);
  // N = 64 MMAs run at half the tensor rate (one M=128,K=16 tcgen05.mma costs ~64 cycles for any N <= 128), so a
  // 128-wide tile with up to 50 % padded columns is never slower than two 64-wide ones
  const int pad64 = (cout + 63) / 64 * 64, pad128 = (cout + 127) / 128 * 128;
  op.bn = cout > 64 ? 128 : 64;
  op.cout_pad = op.bn == 128 ? pad128 : pad64;
  const int kb_per_tap = (cin + 63) / 64, kpad = kb_per_tap * 64;
  std::vector<int64_t> wshape;
  if (mode == ff_s3d::FLAT) wshape = {cout, cin, 1, 1, 1};
  else if (mode == ff_s3d::SPATIAL) wshape = {cout, cin, 1, 3, 3};
  else wshape = {cout, cin, k, 1, 1};
  auto it = h->host_w.find(conv_key);
  if (it == h->host_w.end()) return sfail(h, FF_ERR_STATE, "missing weight '%s'", conv_key.c_str());
  if (h->host_shape[conv_key] != wshape) return sfail(h, FF_ERR_SHAPE, "weight '%s' has the wrong shape", conv_key.c_str());
  const std::vector<float>& w = it->second;
  std::vector<float> wr((size_t)op.cout_pad * op.taps * kpad, 0.0f), scale, shift;
  for (int o = 0; o < cout; ++o)
    for (int ci = 0; ci < cin; ++ci)
      for (int t = 0; t < op.taps; ++t) wr[((size_t)o * op.taps + t) * kpad + ci] = w[((size_t)o * cin + ci) * op.taps + t];
  int rc = s3d_fold_bn(h, bn_key, cout, op.cout_pad, &scale, &shift);
  if (rc) return rc;
  if ((rc = supload(h, &op.w, to_bf16(wr)))) return rc;
  if ((rc = supload(h, &op.scale, scale))) return rc;
  if ((rc = supload(h, &op.shift, shift))) return rc;
  {
    cuuint64_t dims[2] = {(cuuint64_t)op.taps * kpad, (cuuint64_t)op.cout_pad};
    cuuint64_t strides[1] = {(cuuint64_t)op.taps * kpad * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)op.bn};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = ffh::encode_tiled()(&op.tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, op.w, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return sfail(h, FF_ERR_CUDA, "cuTensorMapEncodeTiled(filter %s) failed: %d", conv_key.c_str(), (int)r);
  }
  const int cap_frames_in = h->cap * t_in, cap_frames_out = h->cap * op.t_out;
  bf16* out_view = out + out_off;
  if (mode == ff_s3d::FLAT) {
    op.bw = 128; op.bh = 1; op.bi = 1;
    const long long rows = (long long)cap_frames_in * hw * hw;
    if ((rc = s3d_tmap(h, &op.tmA, in, cin, cin, rows, 1, 1, 128, 1, 1))) return rc;
    if ((rc = s3d_tmap(h, &op.tmO, out_view, cout, out_pitch, rows, 1, 1, 128, 1, 1))) return rc;
  } else if (mode == ff_s3d::SPATIAL) {
    rvk_tile_geometry(hw, &op.bw, &op.bh, &op.bi);
    if ((rc = s3d_tmap(h, &op.tmA, in, cin, cin, hw, hw, cap_frames_in, op.bw, op.bh, op.bi))) return rc;
    if ((rc = s3d_tmap(h, &op.tmO, out_view, cout, out_pitch, hw, hw, cap_frames_out, op.bw, op.bh, op.bi))) return rc;
  } else {
    const int px = hw * hw;
    int lg = 0;
    while ((1 << lg) < px && lg < 7) ++lg;
    op.bw = 1 << lg; op.bh = 128 >> lg; op.bi = 1;
    if ((rc = s3d_tmap(h, &op.tmA, in, cin, cin, px, t_in, h->cap, op.bw, op.bh, 1, 1, stride))) return rc;
    if ((rc = s3d_tmap(h, &op.tmO, out_view, cout, out_pitch, px, op.t_out, h->cap, op.bw, op.bh, 1))) return rc;
  }
  h->ops.push_back(op);
  return FF_OK;
}

void s3d_add_pool(ff_s3d* h, const bf16* in, bf16* out, int c, int hw_in, int t_in, int kt, int ks, int st, int ss, int pt, int ps) {
  ff_s3d::Op op;
  op.kind = ff_s3d::OP_POOL;
  op.name = "maxpool";
  op.pin = in; op.pout = out; op.c = c; op.hw_in = hw_in; op.t_in = t_in;
  op.kt = kt; op.ks = ks; op.st = st; op.ss = ss; op.pt = pt; op.ps = ps;
  op.t_out = (t_in + 2 * pt - kt) / st + 1;
  op.hw = (hw_in + 2 * ps - ks) / ss + 1;
  h->ops.push_back(op);
}

// ---- MaxPool3d on bf16 NDHWC (model.py:19,22,25,31 and branch3 of every Mixed block).  One thread = 8 channels x 4
// consecutive output columns: the (time, row) window is reduced once per INPUT column and the 3*SS + KS column maxima
// are shared by the four outputs (a 3x3x3 stride-1 pool reads 54 vectors per 4 outputs instead of 108).
template <int KS, int SS>
__global__ void __launch_bounds__(256)
s3d_maxpool_kernel(const bf16* __restrict__ in, bf16* __restrict__ out, int n, int t_in, int hw_in, int c8, int t_out, int hw_out,
                   int kt, int st, int pt, int ps) {
  constexpr int NC = 3 * SS + KS;
  const int wg = (hw_out + 3) / 4;
  const size_t total = (size_t)n * t_out * hw_out * wg * c8;
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int q = (int)(i % c8);
  size_t r = i / c8;
  const int ow0 = (int)(r % wg) * 4; r /= wg;
  const int oh = (int)(r % hw_out); r /= hw_out;
  const int ot = (int)(r % t_out);
  const size_t b = r / t_out;
  auto mx = [](uint32_t a0, uint32_t b0) {
    __nv_bfloat162 r2 = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&a0), *reinterpret_cast<__nv_bfloat162*>(&b0));
    return *reinterpret_cast<uint32_t*>(&r2);
  };
  uint4 col[NC];
#pragma unroll
  for (int j = 0; j < NC; ++j) col[j] = make_uint4(0xff80ff80u, 0xff80ff80u, 0xff80ff80u, 0xff80ff80u);   // bf16 -inf pairs
  const int iw0 = ow0 * SS - ps;
  for (int dt = 0; dt < kt; ++dt) {
    const int it = ot * st - pt + dt;
    if (it < 0 || it >= t_in) continue;
#pragma unroll
    for (int dy = 0; dy < KS; ++dy) {
      const int ih = oh * SS - ps + dy;
      if (ih < 0 || ih >= hw_in) continue;
      const bf16* row = in + (((b * t_in + it) * hw_in + ih) * (size_t)hw_in) * c8 * 8 + q * 8;
#pragma unroll
      for (int j = 0; j < NC; ++j) {
        const int iw = iw0 + j;
        if (iw < 0 || iw >= hw_in) continue;
        const uint4 v = *reinterpret_cast<const uint4*>(row + (size_t)iw * c8 * 8);
        col[j].x = mx(col[j].x, v.x); col[j].y = mx(col[j].y, v.y); col[j].z = mx(col[j].z, v.z); col[j].w = mx(col[j].w, v.w);
      }
    }
  }
  bf16* o = out + ((((b * t_out + ot) * hw_out + oh) * (size_t)hw_out + ow0) * c8 + q) * 8;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (ow0 + k >= hw_out) break;
    uint4 m = col[k * SS];
#pragma unroll
    for (int j = 1; j < KS; ++j) {
      const uint4 v = col[k * SS + j];
      m.x = mx(m.x, v.x); m.y = mx(m.y, v.y); m.z = mx(m.z, v.z); m.w = mx(m.w, v.w);
    }
    *reinterpret_cast<uint4*>(o + (size_t)k * c8 * 8) = m;
  }
}

// ---- fp32 NCDHW clip [b,3,T,224,224] (the reference module's input) -> bf16 NHWC4 frames
template <bool F16>
__global__ void __launch_bounds__(256)
s3d_convert_ncdhw_kernel(const float* __restrict__ x, bf16* __restrict__ out, int n, int t) {
  const size_t total = (size_t)n * t * 224 * 224;
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= total) return;
  const size_t pix = i % (224 * 224);
  const size_t f = (i / (224 * 224)) % t, b = i / ((size_t)224 * 224 * t);
  const float* p = x + ((b * 3) * t + f) * 224 * 224 + pix;
  const size_t cs = (size_t)t * 224 * 224;
  reinterpret_cast<uint2*>(out)[i] = make_uint2(ff::pack16x2<F16>(p[0], p[cs]), ff::pack16x2<F16>(p[2 * cs], 0.0f));
}

// ---- SRM high-pass front-end: Conv3d(3, 30, (1,5,5), padding (0,2,2), bias=False) on every frame (SRM/HPF.py:11-37;
// the 30 SRM residual filters arrive as `SRM.hpf.weight` of the checkpoint).  No im2col, same scheme as rvk_stem_kernel /
// conv1_pair_kernel: the frame is fp16 NHWC4 (raw 0..255 are exact in fp16; the filter taps keep 11 mantissa bits, which
// matters for residual filters whose taps sum to zero), TMA brings a 20-row x 24-pixel patch, non-swizzled K-major
// descriptors read it as overlapping windows: accumulator row = a PAIR of output pixels (rows 16 bytes = 2 pixels
// apart), K window of a filter row = the 8 pixels starting two left of the pair (64 bytes = two K=16 steps), N = 2 x 32
// with the pair-expanded filter B[kh][(p,co)][(q,c)] = W[co][c][kh][q-p] (0 <= q-p <= 4).  Tile = 16 x 16 output pixels,
// 10 tcgen05.mma (kind::f16, M=128, N=64) per tile.  Output bf16 NHWC with 32 channels (30 + 2 zero) = the A operand of
// the 7x7 stem convolution that follows; no BatchNorm / activation here (HPF.forward returns the raw filter output).
struct HpfArgs {
  __nv_bfloat16* out;            // [frames,224,224,32] bf16
  const __nv_bfloat16* w;        // [5 kh][64][32] fp16 bit patterns
  int n_img;
};
constexpr int HPF_RING = 3, HPF_PROW = 192, HPF_PROWS = 20, HPF_PSLOT = 3840;
constexpr int HPF_SMEM = HPF_RING * HPF_PSLOT + 5 * 4096 + 128;
__global__ void __launch_bounds__(128, 4)
s3d_hpf_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ HpfArgs a) {
  constexpr int OW = 224, TW = 16, TH = 16, TILES_W = OW / TW, TILES_H = OW / TH, TILES = TILES_W * TILES_H;
  constexpr int RING = HPF_RING, PROW = HPF_PROW, PBYTES = HPF_PROWS * HPF_PROW, PSLOT = HPF_PSLOT;
  extern __shared__ uint8_t hpf_smem_raw[];
  uint8_t* sm = hpf_smem_raw + ((128u - (smem_u32(hpf_smem_raw) & 127u)) & 127u);
  uint8_t (*s_patch)[PSLOT] = reinterpret_cast<uint8_t (*)[PSLOT]>(sm);
  uint8_t* sB = sm + RING * PSLOT;
  __shared__ __align__(8) uint64_t s_bar[1 + RING];
  __shared__ uint32_t s_tmem;
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t bar_mma = smem_u32(&s_bar[0]);
  const uint32_t bar_raw = smem_u32(&s_bar[1]);
  // filter -> core-matrix layout: (n, 16-byte chunk c) at ((n/8)*4 + c)*128 + (n%8)*16
  for (int i = tid; i < 5 * 64 * 4; i += 128) {
    const int kh = i / 256, rem = i % 256, n = rem >> 2, c = rem & 3;
    *reinterpret_cast<uint4*>(sB + kh * 4096 + ((n >> 3) * 4 + c) * 128 + (n & 7) * 16) = reinterpret_cast<const uint4*>(a.w)[i];
  }
  if (tid == 0) {
    tma_prefetch_desc(&tmX);
    mbar_init(bar_mma, 1);
    for (int s = 0; s < RING; ++s) mbar_init(bar_raw + 8 * s, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc<64>(smem_u32(&s_tmem));
  fence_proxy_async_smem();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = s_tmem;
  if (tid == 0) pdl_trigger();
  pdl_wait();
  const uint32_t sB_addr = smem_u32(sB);
  constexpr uint32_t idesc = make_idesc_f16(128, 64);
  const int num_tiles = TILES * a.n_img;
  const int hl = tid >> 3, jl = tid & 7;
  auto issue = [&](int t, int slot) {
    const int n = t / TILES;
    const int rem = t - n * TILES;
    const int th = rem / TILES_W, tw = rem - th * TILES_W;
    mbar_arrive_expect_tx(bar_raw + 8 * slot, PBYTES);
    // patch = pixels w0-2 .. w0+21 (96 fp16 = 192 B; the start is 16-byte aligned), rows h0-2 .. h0+17; zero fill = padding
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
            smem_u32(&s_patch[slot][0])),
        "l"(reinterpret_cast<uint64_t>(&tmX)), "r"(bar_raw + 8 * slot), "r"((tw * TW - 2) * 4), "r"(th * TH - 2), "r"(n)
        : "memory");
  };
  if (tid == 0)
    for (int s = 0; s < RING - 1; ++s) {
      const int t = blockIdx.x + s * gridDim.x;
      if (t < num_tiles) issue(t, s);
    }
  int it = 0;
  for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
    const int n = t / TILES;
    const int rem = t - n * TILES;
    const int th = rem / TILES_W, tw = rem - th * TILES_W;
    const int slot = it % RING;
    if (tid == 0) {
      const int tn = t + (RING - 1) * gridDim.x;        // slot (it+2)%3 was consumed by tile it-1 (all threads passed bar_mma)
      if (tn < num_tiles) issue(tn, (it + RING - 1) % RING);
      mbar_wait(bar_raw + 8 * slot, (it / RING) & 1);
      tcgen05_fence_after();
      const uint32_t patch = smem_u32(&s_patch[slot][0]);
#pragma unroll
      for (int kh = 0; kh < 5; ++kh) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          // A: row (h_l, pair j_l) at patch + (h_l + kh)*192 + j_l*16; K chunks 16 B apart; groups (h_l) one patch row apart
          const uint64_t ad = make_kmajor_desc_noswz(patch + kh * PROW + 32 * j, 16, PROW);
          const uint64_t bd = make_kmajor_desc_noswz(sB_addr + kh * 4096 + 2 * j * 128, 128, 512);
          umma_bf16_ss(tmem, ad, bd, idesc, (kh > 0 || j > 0) ? 1u : 0u);
        }
      }
      umma_commit(bar_mma);
    }
    mbar_wait(bar_mma, it & 1);
    tcgen05_fence_after();
    // thread = pixel pair (h_l, j_l): columns 0..31 = pixel 2j, 32..63 = pixel 2j+1 -> 128 contiguous bytes of NHWC32
    __nv_bfloat16* o = a.out + ((static_cast<size_t>(n) * OW + (th * TH + hl)) * OW + (tw * TW + 2 * jl)) * 32;
#pragma unroll
    for (int p = 0; p < 2; ++p) {
      uint32_t v[32];
      tmem_ld_32x32(tmem + (static_cast<uint32_t>(warp * 32) << 16) + p * 32, v);
      tmem_ld_wait();
      uint32_t pk[16];
#pragma unroll
      for (int c = 0; c < 32; c += 2) pk[c >> 1] = pack_bf16x2(__uint_as_float(v[c]), __uint_as_float(v[c + 1]));
      st_global_v8(o + p * 32, pk);
      st_global_v8(o + p * 32 + 16, pk + 8);
    }
    tcgen05_fence_before();
    __syncthreads();               // every thread has read TMEM and passed bar_mma before the next tile's MMAs / TMA reuse
  }
  __syncthreads();
  if (warp == 0) { tcgen05_fence_after(); tmem_dealloc<64>(tmem); }
}

// ---- head (model.py:40-46): avg_pool3d((2,7,7), stride 1) -> 1x1x1 conv fc (+bias) -> mean over time.  One block per clip.
__global__ void __launch_bounds__(256)
s3d_head_kernel(const bf16* __restrict__ feat, const float* __restrict__ fc_w, const float* __restrict__ fc_b,
                float* __restrict__ logits, int t3, int num_class) {
  __shared__ float s_sum[8][1024];               // per-frame spatial sums (t3 <= 8)
  __shared__ float s_red[8];
  const int b = blockIdx.x;
  const bf16* f = feat + (size_t)b * t3 * 49 * 1024;
  for (int i = threadIdx.x; i < t3 * 1024; i += 256) {
    const int t = i >> 10, c = i & 1023;
    float s = 0.0f;
    for (int p = 0; p < 49; ++p) s += __bfloat162float(f[((size_t)t * 49 + p) * 1024 + c]);
    s_sum[t][c] = s;
  }
  __syncthreads();
  for (int k = 0; k < num_class; ++k) {
    float total = 0.0f;                          // sum over the t3-1 windows of the window's dot product
    for (int t = 0; t + 1 < t3; ++t) {
      float part = 0.0f;
      for (int c = threadIdx.x; c < 1024; c += 256) part = fmaf((s_sum[t][c] + s_sum[t + 1][c]) * (1.0f / 98.0f), fc_w[k * 1024 + c], part);
      part = ff::warp_sum(part);
      if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = part;
      __syncthreads();
      if (threadIdx.x == 0) {
        float d = fc_b[k];
        for (int w = 0; w < 8; ++w) d += s_red[w];
        total += d;
      }
      __syncthreads();
    }
    if (threadIdx.x == 0) logits[b * num_class + k] = total / (float)(t3 - 1);
  }
}

int s3d_finalize(ff_s3d* h) {
  auto bad = [&]() { return h->err.find("shape") != std::string::npos ? FF_ERR_SHAPE : FF_ERR_STATE; };
  int rc;
  const int T = h->frames;
  h->t1 = (T - 1) / 2 + 1;                       // stem temporal conv: k 7, stride 2, pad 3
  h->t2 = (h->t1 - 1) / 2 + 1;                   // MaxPool3d(3, 2, 1)
  h->t3 = h->t2 / 2;                             // MaxPool3d(2, 2, 0)
  if (h->t3 < 2 || h->t3 > 8) return sfail(h, FF_ERR_BAD_ARG, "frames per clip must give 2..8 frames at the head, i.e. T = 16..71 (got %d from T = %d)", h->t3, T);
  // ---- stem spatial conv (1,7,7)/2, 3 -> 64: rvk_stem_kernel's [kh][cout][8 px][4 ch] layout, kw = px - 1
  if (!h->srm) {
    const auto* w = sget(h, "base.0.conv_s.weight", {64, 3, 1, 7, 7});
    if (!w) return bad();
    std::vector<float> ws((size_t)7 * 64 * 32, 0.0f), scale, shift;
    for (int kh = 0; kh < 7; ++kh)
      for (int o = 0; o < 64; ++o)
        for (int kw = 0; kw < 7; ++kw)
          for (int c = 0; c < 3; ++c) ws[((size_t)kh * 64 + o) * 32 + (kw + 1) * 4 + c] = (*w)[(((size_t)o * 3 + c) * 7 + kh) * 7 + kw];
    if ((rc = supload(h, &h->stem_w, to_bf16(ws)))) return rc;
    if ((rc = s3d_fold_bn(h, "base.0.bn_s", 64, 64, &scale, &shift))) return rc;
    for (int o = 0; o < 64; ++o) { h->stem_scale[o] = scale[o]; h->stem_shift[o] = shift[o]; }
    cuuint64_t dims[3] = {896, 224, (cuuint64_t)h->cap * T};
    cuuint64_t strides[2] = {896 * 2, (cuuint64_t)224 * 896 * 2};
    cuuint32_t box[3] = {96, 37, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = ffh::encode_tiled()(&h->tm_x4, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, h->x4, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return sfail(h, FF_ERR_CUDA, "cuTensorMapEncodeTiled(stem input) failed: %d", (int)r);
  }
  h->ops.clear();
  bf16 *A = h->buf[0], *B = h->buf[1], *t1b = h->buf[2], *t2b = h->buf[3], *t3b = h->buf[4];
  if (h->srm) {
    // ---- SRM front-end (model.py:38-39): HPF = Conv3d(3,30,(1,5,5)) -> s3d_hpf_kernel, then base.0.conv_s = (1,7,7)/2 on
    //      30 channels as an implicit GEMM of the persistent kernel.  The hpf tensor has 32 channels per pixel, so two
    //      horizontally adjacent pixels are one 64-channel k-block: output pixel x needs input pixels 2x-3 .. 2x+3 = the
    //      pixel pairs x-2 (second pixel only) .. x+1 — 7 x 4 = 28 full k-blocks instead of 49 half-empty ones, and the
    //      pair axis is read with unit stride (only the row axis keeps elementStride 2)
    const auto* hw_ = sget(h, "SRM.hpf.weight", {30, 3, 1, 5, 5});
    if (!hw_) return bad();
    std::vector<float> wp((size_t)5 * 64 * 32, 0.0f);     // B[kh][(p,co)][(q,c)] = W[co][c][kh][q-p]
    for (int kh = 0; kh < 5; ++kh)
      for (int p = 0; p < 2; ++p)
        for (int o = 0; o < 30; ++o)
          for (int kw = 0; kw < 5; ++kw)
            for (int c = 0; c < 3; ++c)
              wp[((size_t)kh * 64 + p * 32 + o) * 32 + (p + kw) * 4 + c] = (*hw_)[(((size_t)o * 3 + c) * 5 + kh) * 5 + kw];
    if ((rc = supload(h, &h->hpf_w, ffh::to_f16_bits(wp)))) return rc;
    {
      cuuint64_t dims[3] = {896, 224, (cuuint64_t)h->cap * T};
      cuuint64_t strides[2] = {896 * 2, (cuuint64_t)224 * 896 * 2};
      cuuint32_t box[3] = {96, (cuuint32_t)HPF_PROWS, 1};
      cuuint32_t estr[3] = {1, 1, 1};
      CUresult r = ffh::encode_tiled()(&h->tm_x4_hpf, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, h->x4, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                       CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) return sfail(h, FF_ERR_CUDA, "cuTensorMapEncodeTiled(HPF input) failed: %d", (int)r);
    }
    const auto* w = sget(h, "base.0.conv_s.weight", {64, 30, 1, 7, 7});
    if (!w) return bad();
    ff_s3d::Op op;
    op.kind = ff_s3d::OP_CONV;
    op.name = "base.0.conv_s.weight";
    op.mode = ff_s3d::SPATIAL; op.cin = 64; op.cout = 64; op.cout_pad = 64; op.bn = 64; op.taps = 28; op.stride = 2;
    op.hw = 112; op.t_in = T; op.t_out = T;
    std::vector<float> wr((size_t)64 * 28 * 64, 0.0f), scale, shift;      // [cout][kh][pair tap dp = -2..1][pixel of the pair][32 ch]
    for (int o = 0; o < 64; ++o)
      for (int kh = 0; kh < 7; ++kh)
        for (int dpi = 0; dpi < 4; ++dpi)
          for (int half = 0; half < 2; ++half) {
            const int kw = 2 * (dpi - 2) + half + 3;                         // input pixel 2x + 2dp + half = 2x - 3 + kw
            if (kw < 0 || kw > 6) continue;
            for (int ci = 0; ci < 30; ++ci)
              wr[((size_t)o * 28 + kh * 4 + dpi) * 64 + half * 32 + ci] = (*w)[(((size_t)o * 30 + ci) * 7 + kh) * 7 + kw];
          }
    if ((rc = s3d_fold_bn(h, "base.0.bn_s", 64, 64, &scale, &shift))) return rc;
    if ((rc = supload(h, &op.w, to_bf16(wr)))) return rc;
    if ((rc = supload(h, &op.scale, scale))) return rc;
    if ((rc = supload(h, &op.shift, shift))) return rc;
    {
      cuuint64_t dims[2] = {(cuuint64_t)28 * 64, 64};
      cuuint64_t strides[1] = {(cuuint64_t)28 * 64 * 2};
      cuuint32_t box[2] = {64, 64};
      cuuint32_t estr[2] = {1, 1};
      CUresult r = ffh::encode_tiled()(&op.tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, op.w, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                       CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) return sfail(h, FF_ERR_CUDA, "cuTensorMapEncodeTiled(SRM stem filter) failed: %d", (int)r);
    }
    rvk_tile_geometry(112, &op.bw, &op.bh, &op.bi);
    if ((rc = s3d_tmap(h, &op.tmA, h->hpf, 64, 64, 112, 224, h->cap * T, op.bw, op.bh, op.bi, 1, 2))) return rc;
    if ((rc = s3d_tmap(h, &op.tmO, A, 64, 64, 112, 112, h->cap * T, op.bw, op.bh, op.bi))) return rc;
    h->ops.push_back(op);
  }
  auto mark = [&](int base_idx, const bf16* p, int c) { h->ops.back().tap_after = base_idx; h->ops.back().tap_ptr = p; h->ops.back().tap_c = c; };
  // base.0: the spatial half runs in rvk_stem_kernel (frames -> A [.,112,112,64]); temporal (7,1,1)/2 -> B
  if ((rc = s3d_add_conv(h, "base.0.conv_t.weight", "base.0.bn_t", ff_s3d::TEMPORAL, 64, 64, 7, 2, 112, T, A, B, 64, 0))) return rc;
  mark(0, B, 64);
  s3d_add_pool(h, B, A, 64, 112, h->t1, 1, 3, 1, 2, 0, 1);                              // base.1 -> A [.,56,56,64]
  mark(1, A, 64);
  if ((rc = s3d_add_conv(h, "base.2.conv.weight", "base.2.bn", ff_s3d::FLAT, 64, 64, 1, 1, 56, h->t1, A, B, 64, 0))) return rc;
  mark(2, B, 64);
  if ((rc = s3d_add_conv(h, "base.3.conv_s.weight", "base.3.bn_s", ff_s3d::SPATIAL, 64, 192, 3, 1, 56, h->t1, B, A, 192, 0))) return rc;
  if ((rc = s3d_add_conv(h, "base.3.conv_t.weight", "base.3.bn_t", ff_s3d::TEMPORAL, 192, 192, 3, 1, 56, h->t1, A, B, 192, 0))) return rc;
  mark(3, B, 192);
  s3d_add_pool(h, B, A, 192, 56, h->t1, 1, 3, 1, 2, 0, 1);                              // base.4 -> A [.,28,28,192]
  mark(4, A, 192);
  bf16 *X = A, *Y = B;
  int hw = 28, tcur = h->t1;
  for (int mi = 0; mi < 9; ++mi) {
    const S3dMixedPlan& m = kS3dMixed[mi];
    const std::string p = "base." + std::to_string(m.base_idx);
    const int ctot = m.b0 + m.o1 + m.o2 + m.b3;
    if ((rc = s3d_add_conv(h, p + ".branch0.0.conv.weight", p + ".branch0.0.bn", ff_s3d::FLAT, m.cin, m.b0, 1, 1, hw, tcur, X, Y, ctot, 0))) return rc;
    const int mids[2] = {m.m1, m.m2}, outs[2] = {m.o1, m.o2}, offs[2] = {m.b0, m.b0 + m.o1};
    for (int br = 0; br < 2; ++br) {
      const std::string q = p + ".branch" + std::to_string(br + 1);
      if ((rc = s3d_add_conv(h, q + ".0.conv.weight", q + ".0.bn", ff_s3d::FLAT, m.cin, mids[br], 1, 1, hw, tcur, X, t1b, mids[br], 0))) return rc;
      if ((rc = s3d_add_conv(h, q + ".1.conv_s.weight", q + ".1.bn_s", ff_s3d::SPATIAL, mids[br], outs[br], 3, 1, hw, tcur, t1b, t2b, outs[br], 0))) return rc;
      if ((rc = s3d_add_conv(h, q + ".1.conv_t.weight", q + ".1.bn_t", ff_s3d::TEMPORAL, outs[br], outs[br], 3, 1, hw, tcur, t2b, Y, ctot, offs[br]))) return rc;
    }
    s3d_add_pool(h, X, t3b, m.cin, hw, tcur, 3, 3, 1, 1, 1, 1);
    if ((rc = s3d_add_conv(h, p + ".branch3.1.conv.weight", p + ".branch3.1.bn", ff_s3d::FLAT, m.cin, m.b3, 1, 1, hw, tcur, t3b, Y, ctot, m.b0 + m.o1 + m.o2))) return rc;
    mark(m.base_idx, Y, ctot);
    std::swap(X, Y);
    if (m.base_idx == 6) {                       // base.7: MaxPool3d(3, 2, 1)
      s3d_add_pool(h, X, Y, ctot, hw, tcur, 3, 3, 2, 2, 1, 1);
      hw = 14; tcur = h->t2;
      mark(7, Y, ctot);
      std::swap(X, Y);
    } else if (m.base_idx == 12) {               // base.13: MaxPool3d(2, 2, 0)
      s3d_add_pool(h, X, Y, ctot, hw, tcur, 2, 2, 2, 2, 0, 0);
      hw = 7; tcur = h->t3;
      mark(13, Y, ctot);
      std::swap(X, Y);
    }
  }
  h->final_feat = X;
  {
    const auto* w = sget(h, "fc.0.weight", {h->num_class, 1024, 1, 1, 1});
    const auto* b = sget(h, "fc.0.bias", {h->num_class});
    if (!w || !b) return bad();
    if ((rc = supload(h, &h->fc_w, *w))) return rc;
    if ((rc = supload(h, &h->fc_b, *b))) return rc;
  }
  h->host_w.clear();
  h->host_shape.clear();
  h->finalized = true;
  return FF_OK;
}

int s3d_launch_conv(ff_s3d* h, const ff_s3d::Op& op, int n, cudaStream_t st) {
  TcArgs a;
  memset(&a, 0, sizeof(a));
  a.scale = op.scale; a.shift = op.shift;
  a.kb_per_tap = (op.cin + 63) / 64;
  a.kb_total = op.taps * a.kb_per_tap;
  a.kb_per_split = a.kb_total;
  a.cin = op.cin;
  a.cout = op.cout_pad;
  a.taps = op.taps; a.stride = op.stride;
  a.conv_act = 0;
  int bi = op.bi;
  if (op.mode == ff_s3d::FLAT) {
    a.H = 1; a.W = n * op.t_in * op.hw * op.hw;
    a.tiles_w = (a.W + 127) / 128; a.tiles_h = 1;
    a.lg_bw = 7; a.lg_bh = 0;
    a.n_img = 1;
    bi = 1;
  } else if (op.mode == ff_s3d::SPATIAL) {
    a.H = op.hw; a.W = op.hw;
    a.tiles_w = (op.hw + op.bw - 1) / op.bw; a.tiles_h = (op.hw + op.bh - 1) / op.bh;
    a.lg_bw = ilog2(op.bw); a.lg_bh = ilog2(op.bh);
    a.n_img = n * op.t_in;
  } else {
    a.W = op.hw * op.hw; a.H = op.t_out;
    a.tiles_w = (a.W + op.bw - 1) / op.bw; a.tiles_h = (a.H + op.bh - 1) / op.bh;
    a.lg_bw = ilog2(op.bw); a.lg_bh = ilog2(op.bh);
    a.n_img = n;
  }
  const int m_tiles = a.tiles_w * a.tiles_h * ((a.n_img + bi - 1) / bi);
  const int tiles = ((m_tiles + 1) / 2) * (op.cout_pad / op.bn);
  const int grid = std::min(tiles, h->num_sms);
  cudaError_t e = op.bn == 128 ? launch_rvk_conv2<128, 3, false>(grid, st, op.tmA, op.tmB, op.tmO, op.tmO, a)
                               : launch_rvk_conv2<64, 4, false>(grid, st, op.tmA, op.tmB, op.tmO, op.tmO, a);
  if (e != cudaSuccess) return sfail(h, FF_ERR_CUDA, "launch of %s failed: %s", op.name.c_str(), cudaGetErrorString(e));
  ++h->launches;
  return FF_OK;
}

// One pass over n <= cap clips.  layout 0: fp32 NCDHW [n,3,T,224,224]; 2: uint8 [n,T,224,224,3] (raw 0..255 either way).
int s3d_forward(ff_s3d* h, const void* x, int layout, int n, float* logits, cudaStream_t st, int stop_after, const bf16** tap_ptr,
                int64_t* tap_elems) {
  const int T = h->frames, frames = n * T;
  const unsigned blocks = (unsigned)(((size_t)frames * 224 * 224 + 255) / 256);
  cudaError_t e;
  if (h->srm) {
    if (layout == FF_X_NHWC_U8) rvk_convert_kernel<2, true><<<(blocks + 3) / 4, 256, 0, st>>>(x, h->x4, frames, 1.f, 0.f, 1.f, 0.f, 1.f, 0.f);
    else s3d_convert_ncdhw_kernel<true><<<blocks, 256, 0, st>>>(reinterpret_cast<const float*>(x), h->x4, n, T);
    S3_CUDA(h, cudaGetLastError());
    HpfArgs ha;
    ha.out = h->hpf; ha.w = h->hpf_w; ha.n_img = frames;
    e = ffh::launch_smem(s3d_hpf_kernel, dim3(std::min(14 * 14 * frames, h->num_sms * 4)), dim3(128), HPF_SMEM, st, false, h->tm_x4_hpf, ha);
    if (e != cudaSuccess) return sfail(h, FF_ERR_CUDA, "launch of the SRM high-pass kernel failed: %s", cudaGetErrorString(e));
  } else {
    if (layout == FF_X_NHWC_U8) rvk_convert_kernel<2><<<(blocks + 3) / 4, 256, 0, st>>>(x, h->x4, frames, 1.f, 0.f, 1.f, 0.f, 1.f, 0.f);
    else s3d_convert_ncdhw_kernel<false><<<blocks, 256, 0, st>>>(reinterpret_cast<const float*>(x), h->x4, n, T);
    S3_CUDA(h, cudaGetLastError());
    RvkStemArgs sa;
    sa.out = h->buf[0]; sa.w = h->stem_w; sa.n_img = frames;
    for (int o = 0; o < 64; ++o) { sa.scale[o] = h->stem_scale[o]; sa.shift[o] = h->stem_shift[o]; }
    e = ffh::launch_smem(rvk_stem_kernel, dim3(std::min(14 * 7 * frames, h->num_sms * 4)), dim3(128), RVK_STEM_SMEM, st, false, h->tm_x4, sa);
    if (e != cudaSuccess) return sfail(h, FF_ERR_CUDA, "launch of the stem failed: %s", cudaGetErrorString(e));
  }
  h->launches += 2;
  for (const ff_s3d::Op& op : h->ops) {
    if (op.kind == ff_s3d::OP_CONV) {
      int rc = s3d_launch_conv(h, op, n, st);
      if (rc) return rc;
    } else {
      const size_t total = (size_t)n * op.t_out * op.hw * ((op.hw + 3) / 4) * (op.c / 8);
      const unsigned blocks = (unsigned)((total + 255) / 256);
      if (op.ks == 3 && op.ss == 1)
        s3d_maxpool_kernel<3, 1><<<blocks, 256, 0, st>>>(op.pin, op.pout, n, op.t_in, op.hw_in, op.c / 8, op.t_out, op.hw, op.kt, op.st, op.pt, op.ps);
      else if (op.ks == 3 && op.ss == 2)
        s3d_maxpool_kernel<3, 2><<<blocks, 256, 0, st>>>(op.pin, op.pout, n, op.t_in, op.hw_in, op.c / 8, op.t_out, op.hw, op.kt, op.st, op.pt, op.ps);
      else if (op.ks == 2 && op.ss == 2)
        s3d_maxpool_kernel<2, 2><<<blocks, 256, 0, st>>>(op.pin, op.pout, n, op.t_in, op.hw_in, op.c / 8, op.t_out, op.hw, op.kt, op.st, op.pt, op.ps);
      else
        return sfail(h, FF_ERR_STATE, "unsupported pool geometry %d/%d", op.ks, op.ss);
      S3_CUDA(h, cudaGetLastError());
      ++h->launches;
    }
    if (stop_after >= 0 && op.tap_after == stop_after) {
      *tap_ptr = op.tap_ptr;
      *tap_elems = (int64_t)n * op.t_out * op.hw * op.hw * op.tap_c;
      return FF_OK;
    }
  }
  s3d_head_kernel<<<n, 256, 0, st>>>(h->final_feat, h->fc_w, h->fc_b, logits, h->t3, h->num_class);
  S3_CUDA(h, cudaGetLastError());
  ++h->launches;
  return FF_OK;
}

}  // namespace

extern "C" {

int ff_s3d_create(ff_s3d_t** out, int device, int max_clips, int frames_per_clip, int num_class, int srm_net) {
  if (!out || max_clips <= 0 || frames_per_clip <= 0 || num_class <= 0 || num_class > 16 || (srm_net != 0 && srm_net != 1)) return sfail(nullptr, FF_ERR_BAD_ARG, "ff_s3d_create: bad arguments");
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev <= 0)
    return sfail(nullptr, FF_ERR_CUDA, "no CUDA device (%s): libfacfake has no CPU fallback", cudaGetErrorString(e));
  if (device < 0 || device >= ndev) return sfail(nullptr, FF_ERR_BAD_ARG, "device %d out of range (%d devices)", device, ndev);
  cudaDeviceProp prop;
  if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) return sfail(nullptr, FF_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
  if (prop.major != 10) return sfail(nullptr, FF_ERR_CUDA, "device %d is sm_%d%d; libfacfake is built for sm_100a (B200) only", device, prop.major, prop.minor);
  ffh::DeviceGuard guard(device);
  if (guard.status != cudaSuccess) return sfail(nullptr, FF_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(guard.status));
  if (!ffh::encode_tiled()) return sfail(nullptr, FF_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
  ff_s3d* h = new ff_s3d();
  h->device = device;
  h->cap = max_clips;
  h->frames = frames_per_clip;
  h->num_class = num_class;
  h->srm = srm_net;
  cudaDeviceGetAttribute(&h->num_sms, cudaDevAttrMultiProcessorCount, device);
  // largest tensor: the stem's spatial output [clips*T][112][112][64]
  const size_t act = (size_t)max_clips * frames_per_clip * 112 * 112 * 64;
  int rc = FF_OK;
  for (int i = 0; i < 5 && rc == FF_OK; ++i) {
    // buffers 2..4 never hold more than a post-stem tensor (half the frames, quarter of the pixels, <= 3x the channels)
    rc = salloc(h, &h->buf[i], i < 2 ? act : act / 2);
  }
  if (rc == FF_OK) rc = salloc(h, &h->x4, (size_t)max_clips * frames_per_clip * 224 * 224 * 4);
  if (rc == FF_OK && srm_net) rc = salloc(h, &h->hpf, (size_t)max_clips * frames_per_clip * 224 * 224 * 32);
  if (rc != FF_OK) {
    g_s3d_create_error = h->err;
    ff_s3d_destroy(h);
    return rc;
  }
  *out = h;
  return FF_OK;
}

void ff_s3d_destroy(ff_s3d_t* h) {
  if (!h) return;
  {
    ffh::DeviceGuard guard(h->device);
    cudaDeviceSynchronize();
    for (void* p : h->allocs) cudaFree(p);
  }
  delete h;
}

const char* ff_s3d_last_error(const ff_s3d_t* h) { return h ? h->err.c_str() : g_s3d_create_error.c_str(); }

int ff_s3d_load_weight(ff_s3d_t* h, const char* key, const float* host_fp32, const int64_t* shape, int ndim) {
  if (!h || !key || !host_fp32 || ndim < 0 || ndim > 8 || (ndim > 0 && !shape)) return sfail(h, FF_ERR_BAD_ARG, "ff_s3d_load_weight: bad arguments");
  std::lock_guard<std::mutex> lk(h->mu);
  if (h->finalized) return sfail(h, FF_ERR_STATE, "weights are already finalized");
  size_t count = 1;
  std::vector<int64_t> s(shape, shape + ndim);
  for (int64_t d : s) {
    if (d < 0) return sfail(h, FF_ERR_SHAPE, "negative dimension in '%s'", key);
    count *= (size_t)d;
  }
  h->host_w[key].assign(host_fp32, host_fp32 + count);
  h->host_shape[key] = s;
  return FF_OK;
}

int ff_s3d_finalize(ff_s3d_t* h) {
  if (!h) return FF_ERR_BAD_ARG;
  std::lock_guard<std::mutex> lk(h->mu);
  if (h->finalized) return FF_OK;
  ffh::DeviceGuard guard(h->device);
  return s3d_finalize(h);
}

int ff_s3d_forward(ff_s3d_t* h, const void* x, int x_layout, int n, float* logits, void* stream) {
  if (!h || n < 0 || (n > 0 && (!x || !logits))) return sfail(h, FF_ERR_BAD_ARG, "ff_s3d_forward: bad arguments");
  if (x_layout != FF_X_NCHW_F32 && x_layout != FF_X_NHWC_U8) return sfail(h, FF_ERR_BAD_ARG, "ff_s3d_forward: unknown layout %d", x_layout);
  std::lock_guard<std::mutex> lk(h->mu);
  if (!h->finalized) return sfail(h, FF_ERR_STATE, "ff_s3d_finalize() has not been called");
  ffh::DeviceGuard guard(h->device);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const size_t clip_bytes = (size_t)h->frames * 224 * 224 * 3 * (x_layout == FF_X_NHWC_U8 ? 1 : 4);
  for (int s0 = 0; s0 < n; s0 += h->cap) {
    const int ns = std::min(h->cap, n - s0);
    int rc = s3d_forward(h, reinterpret_cast<const uint8_t*>(x) + (size_t)s0 * clip_bytes, x_layout, ns, logits + (size_t)s0 * h->num_class, st, -1, nullptr, nullptr);
    if (rc) return rc;
  }
  return FF_OK;
}

int64_t ff_s3d_debug_activation(ff_s3d_t* h, const void* x, int x_layout, int n, int base_index, float* out_host, int64_t out_elems, void* stream) {
  if (!h || !x || !out_host || n <= 0 || n > h->cap || base_index < 0 || base_index > 15) return sfail(h, FF_ERR_BAD_ARG, "ff_s3d_debug_activation: bad arguments");
  std::lock_guard<std::mutex> lk(h->mu);
  if (!h->finalized) return sfail(h, FF_ERR_STATE, "ff_s3d_finalize() has not been called");
  ffh::DeviceGuard guard(h->device);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const bf16* ptr = nullptr;
  int64_t elems = 0;
  float* lg = nullptr;
  if (cudaMalloc(&lg, (size_t)n * h->num_class * sizeof(float)) != cudaSuccess) return sfail(h, FF_ERR_CUDA, "debug alloc failed");
  int rc = s3d_forward(h, x, x_layout, n, lg, st, base_index, &ptr, &elems);
  int64_t ret = rc;
  if (rc == FF_OK) {
    if (!ptr) ret = sfail(h, FF_ERR_STATE, "debug tap base.%d not reached", base_index);
    else if (elems > out_elems) ret = sfail(h, FF_ERR_BAD_ARG, "debug buffer too small: need %lld floats", (long long)elems);
    else {
      float* tmp = nullptr;
      cudaError_t e = cudaMalloc(&tmp, (size_t)elems * sizeof(float));
      if (e == cudaSuccess) {
        act16_to_f32_kernel<<<(unsigned)((elems + 255) / 256), 256, 0, st>>>(reinterpret_cast<const unsigned short*>(ptr), tmp, (size_t)elems, 0);
        e = cudaMemcpyAsync(out_host, tmp, (size_t)elems * sizeof(float), cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        cudaFree(tmp);
      }
      ret = (e == cudaSuccess) ? elems : sfail(h, FF_ERR_CUDA, "debug copy failed: %s", cudaGetErrorString(e));
    }
  }
  cudaStreamSynchronize(st);
  cudaFree(lg);
  return ret;
}

int64_t ff_s3d_launch_count(const ff_s3d_t* h) { return h ? h->launches : 0; }

}  // extern "C"
