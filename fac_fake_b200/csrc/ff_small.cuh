// ff_small.cuh — the non-tensor-core kernels of the CViT path (HBM- or latency-bound):
//   LayerNorm, 2-token attention, token assembly, cls gather, the 2048->2 head GEMV, the per-video score reduction
//   and the debug-tap conversions.
// Reference lines are /root/reference/CViT-main/{model/cvit.py, cvit_prediction.py}.
#pragma once
#include <cuda_fp16.h>
#include "ff_ptx.cuh"

namespace ff {

// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// LayerNorm(1024), affine (cvit.py:16,20: eps 1e-5; the LinearNorm of the GGCA variant uses 1e-6).  One warp per row;
// fp32 in, bf16 out (next GEMM's A operand).
static __global__ void __launch_bounds__(256)
layernorm_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                 __nv_bfloat16* __restrict__ y, int rows, float eps) {
  pdl_trigger();
  pdl_wait();
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float4* xr = reinterpret_cast<const float4*>(x + static_cast<size_t>(row) * 1024);
  float4 v[8];
  float s = 0.0f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    v[i] = xr[i * 32 + lane];
    s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  }
  const float mean = warp_sum(s) * (1.0f / 1024.0f);
  float q = 0.0f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
    q += (a * a + b * b) + (c * c + d * d);
  }
  const float rstd = rsqrtf(warp_sum(q) * (1.0f / 1024.0f) + eps);
  const float4* g4 = reinterpret_cast<const float4*>(gamma);
  const float4* b4 = reinterpret_cast<const float4*>(beta);
  uint2* yr = reinterpret_cast<uint2*>(y + static_cast<size_t>(row) * 1024);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float4 g = g4[i * 32 + lane], b = b4[i * 32 + lane];
    const float o0 = (v[i].x - mean) * rstd * g.x + b.x;
    const float o1 = (v[i].y - mean) * rstd * g.y + b.y;
    const float o2 = (v[i].z - mean) * rstd * g.z + b.z;
    const float o3 = (v[i].w - mean) * rstd * g.w + b.w;
    yr[i * 32 + lane] = make_uint2(pack_bf16x2(o0, o1), pack_bf16x2(o2, o3));
  }
}

// 2-token attention (cvit.py:43-60): one warp per (crop, head).  qkv fp32 [2n][3072] with feature index
// which*1024 + head*128 + d (cvit.py:46); scale = dim**-0.5 = 1/32 (cvit.py:38); out bf16 [2n][1024] '(h d)'.
static __global__ void __launch_bounds__(256)
attention2_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out, int n_crops) {
  pdl_trigger();
  pdl_wait();
  const int wid = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (wid >= n_crops * 8) return;
  const int b = wid >> 3, h = wid & 7;
  const __nv_bfloat16* r0 = qkv + static_cast<size_t>(2 * b) * 3072 + h * 128 + lane * 4;
  const __nv_bfloat16* r1 = r0 + 3072;
  auto ld4 = [](const __nv_bfloat16* p) {
    const uint2 u = *reinterpret_cast<const uint2*>(p);
    const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
    const float2 c = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
    return make_float4(a.x, a.y, c.x, c.y);
  };
  const float4 q0 = ld4(r0), q1 = ld4(r1);
  const float4 k0 = ld4(r0 + 1024), k1 = ld4(r1 + 1024);
  const float4 v0 = ld4(r0 + 2048), v1 = ld4(r1 + 2048);
  auto dot = [](const float4& a, const float4& c) { return (a.x * c.x + a.y * c.y) + (a.z * c.z + a.w * c.w); };
  const float s = 0.03125f;
  const float d00 = warp_sum(dot(q0, k0)) * s, d01 = warp_sum(dot(q0, k1)) * s;
  const float d10 = warp_sum(dot(q1, k0)) * s, d11 = warp_sum(dot(q1, k1)) * s;
  const float m0 = fmaxf(d00, d01), m1 = fmaxf(d10, d11);
  const float e00 = expf(d00 - m0), e01 = expf(d01 - m0), e10 = expf(d10 - m1), e11 = expf(d11 - m1);
  const float i0 = 1.0f / (e00 + e01), i1 = 1.0f / (e10 + e11);
  const float a00 = e00 * i0, a01 = e01 * i0, a10 = e10 * i1, a11 = e11 * i1;
  __nv_bfloat16* o0 = out + static_cast<size_t>(2 * b) * 1024 + h * 128 + lane * 4;
  *reinterpret_cast<uint2*>(o0) = make_uint2(pack_bf16x2(a00 * v0.x + a01 * v1.x, a00 * v0.y + a01 * v1.y),
                                              pack_bf16x2(a00 * v0.z + a01 * v1.z, a00 * v0.w + a01 * v1.w));
  *reinterpret_cast<uint2*>(o0 + 1024) = make_uint2(pack_bf16x2(a10 * v0.x + a11 * v1.x, a10 * v0.y + a11 * v1.y),
                                                     pack_bf16x2(a10 * v0.z + a11 * v1.z, a10 * v0.w + a11 * v1.w));
}

// Sum over the 32 lanes of a warp in xor-butterfly order: every lane ends with the same bits.  The encoder's row mean is
// defined as this sum of the row's 32 segment sums / 1024 wherever it is (re)computed (tokens_kernel, ff_xf.cuh).
__device__ __forceinline__ float warp_sum_bfly(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Token assembly (cvit.py:171-175): tok0 = cls + pos[slot], tok1 = (patch embedding + bias) + pos[slot].
// The patch embedding arrives as n_splits split-K partial slabs that are summed here in a fixed order.
// For the encoder kernel (xb != nullptr) it also writes what the first LayerNorm-folded GEMM reads: xb = bf16(x - mean(x))
// and, per 32-column segment, (sum, centred sum of squares) into stats_a; stats_b receives the same sums (its row mean is
// the shift xb was written with, ff_xf.cuh).
static __global__ void __launch_bounds__(256)
tokens_kernel(const float* __restrict__ emb, int n_splits, long long split_stride, const float* __restrict__ bias,
              const float* __restrict__ cls, const float* __restrict__ pos, const int* __restrict__ slot, int slot_base,
              float* __restrict__ x, int n, __nv_bfloat16* __restrict__ xb, float2* __restrict__ stats_a,
              float2* __restrict__ stats_b) {
  __shared__ float s_seg[2][32];
  pdl_trigger();
  pdl_wait();
  const int b = blockIdx.x;
  if (b >= n) return;
  const int s = slot ? slot[b] : ((slot_base + b) & 31);
  const int i = threadIdx.x;   // 256 threads x float4 = 1024
  const float4 p = reinterpret_cast<const float4*>(pos + static_cast<size_t>(s) * 1024)[i];
  const float4 c = reinterpret_cast<const float4*>(cls)[i];
  float4 e = reinterpret_cast<const float4*>(emb + static_cast<size_t>(b) * 1024)[i];
  for (int z = 1; z < n_splits; ++z) {   // split-K partial sums, fixed order => deterministic
    const float4 t = reinterpret_cast<const float4*>(emb + z * split_stride + static_cast<size_t>(b) * 1024)[i];
    e.x += t.x; e.y += t.y; e.z += t.z; e.w += t.w;
  }
  const float4 bb = bias ? reinterpret_cast<const float4*>(bias)[i] : make_float4(0.f, 0.f, 0.f, 0.f);
  float4* x0 = reinterpret_cast<float4*>(x + static_cast<size_t>(2 * b) * 1024);
  const float4 t0 = make_float4(c.x + p.x, c.y + p.y, c.z + p.z, c.w + p.w);
  const float4 t1 = make_float4((e.x + bb.x) + p.x, (e.y + bb.y) + p.y, (e.z + bb.z) + p.z, (e.w + bb.w) + p.w);
  x0[i] = t0;
  x0[256 + i] = t1;
  if (xb == nullptr) return;
  const float4 tok[2] = {t0, t1};
  const int seg = i >> 3;               // 8 threads x 4 columns = one 32-column segment
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    float sum = (tok[r].x + tok[r].y) + (tok[r].z + tok[r].w);
    sum += __shfl_xor_sync(0xffffffffu, sum, 1);
    sum += __shfl_xor_sync(0xffffffffu, sum, 2);
    sum += __shfl_xor_sync(0xffffffffu, sum, 4);
    const float mu = sum * (1.0f / 32.0f);
    const float d0 = tok[r].x - mu, d1 = tok[r].y - mu, d2 = tok[r].z - mu, d3 = tok[r].w - mu;
    float m2 = (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
    m2 += __shfl_xor_sync(0xffffffffu, m2, 1);
    m2 += __shfl_xor_sync(0xffffffffu, m2, 2);
    m2 += __shfl_xor_sync(0xffffffffu, m2, 4);
    if ((i & 7) == 0) {
      const size_t o = static_cast<size_t>(2 * b + r) * 32 + seg;
      stats_a[o] = make_float2(sum, m2);
      stats_b[o] = make_float2(sum, 0.0f);
      s_seg[r][seg] = sum;
    }
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const float mean = warp_sum_bfly(s_seg[r][i & 31]) * (1.0f / 1024.0f);
    reinterpret_cast<uint2*>(xb + static_cast<size_t>(2 * b + r) * 1024)[i] =
        make_uint2(pack_bf16x2(tok[r].x - mean, tok[r].y - mean), pack_bf16x2(tok[r].z - mean, tok[r].w - mean));
  }
}

// cls select (cvit.py:177): bf16 copy of token 0 of every crop -> A operand of mlp_head.0.
static __global__ void __launch_bounds__(256)
cls_gather_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, int n) {
  pdl_trigger();
  pdl_wait();
  const int b = blockIdx.x;
  if (b >= n) return;
  const float4 v = reinterpret_cast<const float4*>(x + static_cast<size_t>(2 * b) * 1024)[threadIdx.x];
  reinterpret_cast<uint2*>(out + static_cast<size_t>(b) * 1024)[threadIdx.x] =
      make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
}

// mlp_head.2: Linear(2048 -> 2) (cvit.py:164) — one warp per crop, fp32.
static __global__ void __launch_bounds__(256)
head2_kernel(const float* __restrict__ hid, const float* __restrict__ w, const float* __restrict__ bias,
             float* __restrict__ logits, int n) {
  pdl_trigger();
  pdl_wait();
  const int b = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (b >= n) return;
  const float4* h4 = reinterpret_cast<const float4*>(hid + static_cast<size_t>(b) * 2048);
  const float4* w0 = reinterpret_cast<const float4*>(w);
  const float4* w1 = reinterpret_cast<const float4*>(w + 2048);
  float s0 = 0.0f, s1 = 0.0f;
#pragma unroll 4
  for (int i = lane; i < 512; i += 32) {
    const float4 hv = h4[i], a = w0[i], c = w1[i];
    s0 += (hv.x * a.x + hv.y * a.y) + (hv.z * a.z + hv.w * a.w);
    s1 += (hv.x * c.x + hv.y * c.y) + (hv.z * c.z + hv.w * c.w);
  }
  s0 = warp_sum(s0);
  s1 = warp_sum(s1);
  if (lane == 0) {
    logits[2 * b] = s0 + bias[0];
    logits[2 * b + 1] = s1 + bias[1];
  }
}

// Per-video reduction (cvit_prediction.py:258-281): sigmoid per logit, mean over the video's frames,
// f if f > r else |1 - r|; <= 2 frames (or none) -> 0.5.  One warp per video.  Only the first `max_frames` frames of a
// video count: the reference evaluates the chunks [0:32],[32:64],[64:90] and drops the rest (cvit_prediction.py:224-238).
// mode 1: mean of softmax(logits)[0] (extra).  mode 2: the rows already are pred_sig outputs (probabilities).
static __global__ void __launch_bounds__(256)
video_reduce_kernel(const float* __restrict__ logits, const int* __restrict__ off, int n_videos, int mode, int max_frames,
                    float* __restrict__ scores) {
  pdl_trigger();
  pdl_wait();
  const int v = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (v >= n_videos) return;
  const int a = off[v];
  const int e = min(off[v + 1], a + max_frames);
  const int cnt = e - a;
  float f = 0.0f, r = 0.0f;
  for (int i = a + lane; i < e; i += 32) {
    const float2 z = *reinterpret_cast<const float2*>(logits + 2 * static_cast<size_t>(i));
    if (mode == 0) {
      f += 1.0f / (1.0f + expf(-z.x));
      r += 1.0f / (1.0f + expf(-z.y));
    } else if (mode == 2) {
      f += z.x;
      r += z.y;
    } else {
      f += 1.0f / (1.0f + expf(z.y - z.x));
    }
  }
  f = warp_sum(f);
  r = warp_sum(r);
  if (lane == 0) {
    float s = 0.5f;
    if (mode == 0 || mode == 2) {
      if (cnt > 2) {
        const float fc = f / static_cast<float>(cnt), rc = r / static_cast<float>(cnt);
        s = (fc > rc) ? fc : fabsf(1.0f - rc);
      }
    } else if (cnt > 0) {
      s = f / static_cast<float>(cnt);
    }
    scores[v] = s;
  }
}

// ---- debug-tap conversions: 16-bit activations (bf16, or fp16 when f16 != 0) -> fp32
__device__ __forceinline__ float act16_to_f32(unsigned short bits, int f16) {
  if (f16) { __half hv; memcpy(&hv, &bits, 2); return __half2float(hv); }
  __nv_bfloat16 bv; memcpy(&bv, &bits, 2); return __bfloat162float(bv);
}
static __global__ void act16_to_f32_kernel(const unsigned short* __restrict__ in, float* __restrict__ out, size_t n, int f16) {
  size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (i < n) out[i] = act16_to_f32(in[i], f16);
}
// channel-blocked [n][2][hw][hw][32] -> NHWC fp32 [n][hw][hw][64] (debug tap of feature layers 4, 5)
static __global__ void unblock_act16_to_f32_kernel(const unsigned short* __restrict__ in, float* __restrict__ out, int n, int hw, int f16) {
  const size_t total = static_cast<size_t>(n) * hw * hw * 64;
  size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (i >= total) return;
  const int c = static_cast<int>(i % 64);
  const size_t pix = i / 64;
  const size_t per = static_cast<size_t>(hw) * hw;
  const size_t img = pix / per, rem = pix % per;
  out[i] = act16_to_f32(in[((img * 2 + c / 32) * per + rem) * 32 + (c % 32)], f16);
}
// slot = (frame index within the video) % 32: the reference's [0:32],[32:64],[64:90] chunking (cvit_prediction.py:226-238)
static __global__ void slots_from_offsets_kernel(const int* __restrict__ off, int n_videos, int* __restrict__ slot) {
  const int v = blockIdx.x;
  if (v >= n_videos) return;
  const int a = off[v], e = off[v + 1];
  for (int i = a + threadIdx.x; i < e; i += blockDim.x) slot[i] = (i - a) & 31;
}

}  // namespace ff
