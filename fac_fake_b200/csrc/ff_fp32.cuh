// ff_fp32.cuh — FF_COMPUTE_FP32 path: plain fp32 CUDA-core kernels for the whole CViT forward.
// Slow by design (no tensor cores); it exists for the north-star's "1e-4 (fp32 path)" parity gate and as an
// on-device cross-check of the bf16/tcgen05 path.  Reference ops: /root/reference/CViT-main/model/cvit.py.
#pragma once
#include "ff_ptx.cuh"
#include "ff_small.cuh"

namespace ff {

// conv3x3 pad 1 + (scale, shift) + ReLU [+ 2x2 max-pool], NHWC fp32 -> NHWC fp32 (cvit.py:88-147).
// One thread per output element (channel fastest).  in_kind 0: NHWC fp32 `in`; 1: fp32 NCHW raw input (cin=3);
// 2: uint8 NHWC raw crops with (x/255-mean)/std fused (cvit_prediction.py:41-45,214-215).
static __global__ void __launch_bounds__(256)
conv3x3_fp32_kernel(const void* __restrict__ raw, int in_kind, const float* __restrict__ in, const float* __restrict__ w,
                    const float* __restrict__ scale, const float* __restrict__ shift, float* __restrict__ out, int n_img,
                    int hw, int cin, int cout, int pool, int relu) {
  const int ohw = pool ? hw / 2 : hw;
  const size_t total = static_cast<size_t>(n_img) * ohw * ohw * cout;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int co = static_cast<int>(idx % cout);
    size_t t = idx / cout;
    const int ow = static_cast<int>(t % ohw); t /= ohw;
    const int oh = static_cast<int>(t % ohw);
    const int n = static_cast<int>(t / ohw);
    const float* wr = w + static_cast<size_t>(co) * 9 * cin;
    float best = relu ? 0.0f : -INFINITY;   // ReLU output is >= 0, so 0 is a valid identity for the max
    const int reps = pool ? 2 : 1;
    for (int dy = 0; dy < reps; ++dy)
      for (int dx = 0; dx < reps; ++dx) {
        const int y = pool ? 2 * oh + dy : oh, x = pool ? 2 * ow + dx : ow;
        float acc = 0.0f;
        for (int kh = 0; kh < 3; ++kh) {
          const int iy = y + kh - 1;
          if (iy < 0 || iy >= hw) continue;
          for (int kw = 0; kw < 3; ++kw) {
            const int ix = x + kw - 1;
            if (ix < 0 || ix >= hw) continue;
            const float* wt = wr + (kh * 3 + kw) * cin;
            if (in_kind == 0) {
              const float* p = in + ((static_cast<size_t>(n) * hw + iy) * hw + ix) * cin;
              for (int c = 0; c < cin; c += 4) {
                const float4 a = *reinterpret_cast<const float4*>(p + c);
                const float4 b = *reinterpret_cast<const float4*>(wt + c);
                acc = fmaf(a.x, b.x, acc); acc = fmaf(a.y, b.y, acc); acc = fmaf(a.z, b.z, acc); acc = fmaf(a.w, b.w, acc);
              }
            } else {
              for (int c = 0; c < 3; ++c) {
                float v;
                if (in_kind == 2) {
                  const float u = static_cast<float>(reinterpret_cast<const uint8_t*>(raw)[((static_cast<size_t>(n) * hw + iy) * hw + ix) * 3 + c]);
                  const float mean = c == 0 ? 0.485f : (c == 1 ? 0.456f : 0.406f);
                  const float sd = c == 0 ? 0.229f : (c == 1 ? 0.224f : 0.225f);
                  v = __fdiv_rn(__fdiv_rn(u, 255.0f) - mean, sd);
                } else {
                  v = reinterpret_cast<const float*>(raw)[((static_cast<size_t>(n) * 3 + c) * hw + iy) * hw + ix];
                }
                acc = fmaf(v, wt[c], acc);
              }
            }
          }
        }
        best = fmaxf(best, fmaf(acc, scale[co], shift[co]));
      }
    out[idx] = best;
  }
}

// out[M][N] (=|+=) act(A[M][K] * W[N][K]^T + bias).  64x64 tile, 256 threads, 4x4 per thread, K step 16.
static __global__ void __launch_bounds__(256)
linear_fp32_kernel(const float* __restrict__ A, const float* __restrict__ W, const float* __restrict__ bias,
                   float* __restrict__ out, int M, int N, int K, int act, int resid) {
  __shared__ float sa[16][64 + 4];
  __shared__ float sw[16][64 + 4];
  const int n0 = blockIdx.x * 64, m0 = blockIdx.y * 64;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[4][4] = {};
  const int lr = threadIdx.x >> 2;          // 0..63: row inside the tile
  const int lk = (threadIdx.x & 3) * 4;     // 0,4,8,12
  for (int k0 = 0; k0 < K; k0 += 16) {
    float4 a = make_float4(0, 0, 0, 0), b = make_float4(0, 0, 0, 0);
    if (m0 + lr < M) a = *reinterpret_cast<const float4*>(A + static_cast<size_t>(m0 + lr) * K + k0 + lk);
    if (n0 + lr < N) b = *reinterpret_cast<const float4*>(W + static_cast<size_t>(n0 + lr) * K + k0 + lk);
    sa[lk][lr] = a.x; sa[lk + 1][lr] = a.y; sa[lk + 2][lr] = a.z; sa[lk + 3][lr] = a.w;
    sw[lk][lr] = b.x; sw[lk + 1][lr] = b.y; sw[lk + 2][lr] = b.z; sw[lk + 3][lr] = b.w;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      float av[4], bv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { av[i] = sa[k][ty * 4 + i]; bv[i] = sw[k][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = acc[i][j] + (bias ? bias[n] : 0.0f);
      if (act == 1) v = fmaxf(v, 0.0f);
      else if (act == 2) v = 0.5f * v * (1.0f + erff(v * 0.70710678118654752440f));
      float* o = out + static_cast<size_t>(m) * N + n;
      *o = resid ? (*o + v) : v;
    }
  }
}

static __global__ void __launch_bounds__(256)
layernorm_f32_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                     float* __restrict__ y, int rows, float eps) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* xr = x + static_cast<size_t>(row) * 1024;
  float v[32];
  float s = 0.0f;
#pragma unroll
  for (int i = 0; i < 32; ++i) { v[i] = xr[i * 32 + lane]; s += v[i]; }
  const float mean = warp_sum(s) * (1.0f / 1024.0f);
  float q = 0.0f;
#pragma unroll
  for (int i = 0; i < 32; ++i) { const float d = v[i] - mean; q += d * d; }
  const float rstd = 1.0f / sqrtf(warp_sum(q) * (1.0f / 1024.0f) + eps);
  float* yr = y + static_cast<size_t>(row) * 1024;
#pragma unroll
  for (int i = 0; i < 32; ++i) yr[i * 32 + lane] = (v[i] - mean) * rstd * gamma[i * 32 + lane] + beta[i * 32 + lane];
}

static __global__ void __launch_bounds__(256)
attention2_f32_kernel(const float* __restrict__ qkv, float* __restrict__ out, int n_crops) {
  const int wid = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (wid >= n_crops * 8) return;
  const int b = wid >> 3, h = wid & 7;
  const float* r0 = qkv + static_cast<size_t>(2 * b) * 3072 + h * 128 + lane * 4;
  const float* r1 = r0 + 3072;
  const float4 q0 = *reinterpret_cast<const float4*>(r0), q1 = *reinterpret_cast<const float4*>(r1);
  const float4 k0 = *reinterpret_cast<const float4*>(r0 + 1024), k1 = *reinterpret_cast<const float4*>(r1 + 1024);
  const float4 v0 = *reinterpret_cast<const float4*>(r0 + 2048), v1 = *reinterpret_cast<const float4*>(r1 + 2048);
  auto dot = [](const float4& a, const float4& c) { return (a.x * c.x + a.y * c.y) + (a.z * c.z + a.w * c.w); };
  const float s = 0.03125f;
  const float d00 = warp_sum(dot(q0, k0)) * s, d01 = warp_sum(dot(q0, k1)) * s;
  const float d10 = warp_sum(dot(q1, k0)) * s, d11 = warp_sum(dot(q1, k1)) * s;
  const float m0 = fmaxf(d00, d01), m1 = fmaxf(d10, d11);
  const float e00 = expf(d00 - m0), e01 = expf(d01 - m0), e10 = expf(d10 - m1), e11 = expf(d11 - m1);
  const float i0 = 1.0f / (e00 + e01), i1 = 1.0f / (e10 + e11);
  const float a00 = e00 * i0, a01 = e01 * i0, a10 = e10 * i1, a11 = e11 * i1;
  float* o0 = out + static_cast<size_t>(2 * b) * 1024 + h * 128 + lane * 4;
  *reinterpret_cast<float4*>(o0) = make_float4(a00 * v0.x + a01 * v1.x, a00 * v0.y + a01 * v1.y, a00 * v0.z + a01 * v1.z, a00 * v0.w + a01 * v1.w);
  *reinterpret_cast<float4*>(o0 + 1024) = make_float4(a10 * v0.x + a11 * v1.x, a10 * v0.y + a11 * v1.y, a10 * v0.z + a11 * v1.z, a10 * v0.w + a11 * v1.w);
}

// ---- fp32 path of the ResVitKan trunk (ResVitKan/ResVitKan.py:150-240)
// raw input -> normalised NHWC fp32 with 4 channels (4th = 0).  in_kind 1: fp32 NCHW (already normalised);
// 2: uint8 NHWC crops, (u/255 - mean)/std in the reference's operation order (cvit_prediction.py:41-45,214-215).
static __global__ void __launch_bounds__(256)
nhwc4_f32_kernel(const void* __restrict__ raw, int in_kind, float4* __restrict__ out, int n_img) {
  const size_t total = static_cast<size_t>(n_img) * 224 * 224;
  const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (i >= total) return;
  float v[3];
  if (in_kind == 2) {
    const uint8_t* p = reinterpret_cast<const uint8_t*>(raw) + i * 3;
    const float mean[3] = {0.485f, 0.456f, 0.406f}, sd[3] = {0.229f, 0.224f, 0.225f};
#pragma unroll
    for (int c = 0; c < 3; ++c) v[c] = __fdiv_rn(__fdiv_rn(static_cast<float>(p[c]), 255.0f) - mean[c], sd[c]);
  } else {
    const size_t n = i / (224 * 224), px = i % (224 * 224);
#pragma unroll
    for (int c = 0; c < 3; ++c) v[c] = reinterpret_cast<const float*>(raw)[(n * 3 + c) * 224 * 224 + px];
  }
  out[i] = make_float4(v[0], v[1], v[2], 0.0f);
}

// Any k x k convolution (pad = k/2, stride 1 or 2) + (scale, shift) [+ ReLU] [+ residual, ReLU], NHWC fp32, cin % 4 == 0.
// One thread per output element; w = [cout][k*k][cin].  The bottleneck's two ReLUs (before and after the residual add,
// ResVitKan.py:169-176) are `relu` and the one implied by `resid`.
static __global__ void __launch_bounds__(256)
conv_fp32_kernel(const float* __restrict__ in, const float* __restrict__ w, const float* __restrict__ scale,
                 const float* __restrict__ shift, const float* __restrict__ resid, float* __restrict__ out, int n_img,
                 int in_hw, int out_hw, int cin, int cout, int k, int stride, int relu) {
  const size_t total = static_cast<size_t>(n_img) * out_hw * out_hw * cout;
  const int pad = k / 2;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int co = static_cast<int>(idx % cout);
    size_t t = idx / cout;
    const int ow = static_cast<int>(t % out_hw); t /= out_hw;
    const int oh = static_cast<int>(t % out_hw);
    const size_t n = t / out_hw;
    const float* wr = w + static_cast<size_t>(co) * k * k * cin;
    float acc = 0.0f;
    for (int kh = 0; kh < k; ++kh) {
      const int iy = oh * stride + kh - pad;
      if (iy < 0 || iy >= in_hw) continue;
      for (int kw = 0; kw < k; ++kw) {
        const int ix = ow * stride + kw - pad;
        if (ix < 0 || ix >= in_hw) continue;
        const float* p = in + ((n * in_hw + iy) * in_hw + ix) * cin;
        const float* wt = wr + (kh * k + kw) * cin;
        for (int c = 0; c < cin; c += 4) {
          const float4 a = *reinterpret_cast<const float4*>(p + c);
          const float4 b = *reinterpret_cast<const float4*>(wt + c);
          acc = fmaf(a.x, b.x, acc); acc = fmaf(a.y, b.y, acc); acc = fmaf(a.z, b.z, acc); acc = fmaf(a.w, b.w, acc);
        }
      }
    }
    float v = fmaf(acc, scale[co], shift[co]);
    if (relu) v = fmaxf(v, 0.0f);
    if (resid) v = fmaxf(v + resid[idx], 0.0f);
    out[idx] = v;
  }
}

// MaxPool2d(3, stride 2, pad 1), NHWC fp32 (ResVitKan.py:194,232)
static __global__ void __launch_bounds__(256)
maxpool3s2_f32_kernel(const float* __restrict__ in, float* __restrict__ out, int n_img, int in_hw, int c) {
  const int out_hw = in_hw / 2;
  const size_t total = static_cast<size_t>(n_img) * out_hw * out_hw * c;
  const size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (idx >= total) return;
  const int ch = static_cast<int>(idx % c);
  size_t t = idx / c;
  const int ow = static_cast<int>(t % out_hw); t /= out_hw;
  const int oh = static_cast<int>(t % out_hw);
  const size_t n = t / out_hw;
  float m = -INFINITY;
  for (int dy = -1; dy <= 1; ++dy)
    for (int dx = -1; dx <= 1; ++dx) {
      const int iy = 2 * oh + dy, ix = 2 * ow + dx;
      if (iy < 0 || iy >= in_hw || ix < 0 || ix >= in_hw) continue;
      m = fmaxf(m, in[((n * in_hw + iy) * in_hw + ix) * c + ch]);
    }
  out[idx] = m;
}

static __global__ void __launch_bounds__(256)
cls_gather_f32_kernel(const float* __restrict__ x, float* __restrict__ out, int n) {
  const int b = blockIdx.x;
  if (b >= n) return;
  reinterpret_cast<float4*>(out + static_cast<size_t>(b) * 1024)[threadIdx.x] =
      reinterpret_cast<const float4*>(x + static_cast<size_t>(2 * b) * 1024)[threadIdx.x];
}

}  // namespace ff
