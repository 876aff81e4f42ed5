// ff_c1.cuh — feature layer 1 as a stand-alone launch: Conv2d(3,32,3,p=1)+BN+ReLU (cvit.py:88-90).
// The default uint8 path fuses this layer into c12_kernel (ff_c12.cuh); the two kernels here serve
//   * conv1_f32_kernel: `model(x)` compatibility — fp32 NCHW input that the caller already normalised
//     (cvit_prediction.py:209-229).  K = 27 is too small for a TMA-fed implicit GEMM, so the CTA builds the A operand:
//     the (16+2) x (8+2) patch of an 8x16-pixel tile is staged in shared memory as bf16 with 4 channels per pixel
//     (c = 3 is zero, 0 outside the image = the padding), each thread copies per filter row the 3 neighbouring pixels
//     (24 contiguous bytes) into its 128-byte K-major row in the SWIZZLE_128B pattern (K index = kh*16 + kw*4 + c),
//     three tcgen05.mma (M=128, N=32, K=16) produce the tile, registers -> global epilogue;
//   * conv1_pair_kernel: uint8 crops, no im2col at all (debug tap 1 and the reference point of c12_kernel's first half).
// HBM-bound by construction: 150 KB (uint8) or 602 KB (fp32) read + 3.2 MB written per crop.
#pragma once
#include "ff_tc.cuh"

namespace ff {

struct C1Args {
  const float* x;                // fp32 NCHW [n,3,224,224], already normalised (what `model(x)` receives)
  __nv_bfloat16* out;            // bf16 NHWC [n,224,224,32]
  const __nv_bfloat16* w;        // [32][64] bf16: [cout][k], k = kh*16 + kw*4 + cin (other columns zero)
  int n_img;
  float scale[32];               // folded BN scale / shift, read as constant-bank operands by the epilogue FMAs
  float shift[32];
};

template <bool F16 = false>
__global__ void __launch_bounds__(128, 6)
conv1_f32_kernel(const __grid_constant__ C1Args a) {
  constexpr int HW = 224, TW = 8, TH = 16, PW = TW + 2, PH = TH + 2;
  constexpr int TILES_W = HW / TW, TILES_H = HW / TH, TILES = TILES_W * TILES_H;
  constexpr int NELEM = PH * PW * 3;
  __shared__ __align__(1024) uint8_t sA[128 * 128];
  __shared__ __align__(1024) uint8_t sB[32 * 128];
  __shared__ __align__(16) unsigned short s_in[PH * PW * 4];     // bf16 or fp16 bit patterns
  __shared__ __align__(8) uint64_t s_bar;
  __shared__ uint32_t s_tmem;

  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t bar = smem_u32(&s_bar);
  // ---- one-time setup: weights (swizzled), LUT, scale/shift, zeroed patch (channel 3 stays zero), barrier, TMEM
  for (int i = tid; i < 32 * 8; i += 128) {
    const int row = i >> 3, ch = i & 7;
    *reinterpret_cast<uint4*>(sB + row * 128 + ((ch ^ (row & 7)) << 4)) = reinterpret_cast<const uint4*>(a.w)[i];
  }
  for (int i = tid; i < PH * PW * 4 / 8; i += 128) reinterpret_cast<uint4*>(s_in)[i] = make_uint4(0, 0, 0, 0);
  for (int i = tid; i < 128 * 8; i += 128) reinterpret_cast<uint4*>(sA)[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc<32>(smem_u32(&s_tmem));
  fence_proxy_async_smem();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = s_tmem;
  if (tid == 0) pdl_trigger();
  pdl_wait();                      // inputs may come from, and bufA may still be read by, the previous kernel
  const uint32_t sA_addr = smem_u32(sA), sB_addr = smem_u32(sB);
  constexpr uint32_t idesc = make_idesc_16<F16>(128, 32);

  const int num_tiles = TILES * a.n_img;
  const int hl = tid >> 3, wl = tid & 7;
  // Software pipeline: the patch of the NEXT tile is fetched into registers while the current tile is processed.
  constexpr int NPF = (NELEM + 127) / 128;
  float pf[NPF];
  // element i of the patch -> (py, px, c), walked plane by plane (NCHW input)
  auto decode = [&](int i, int* py, int* px, int* c) {
    *c = i / (PH * PW); const int r = i - *c * (PH * PW); *py = r / PW; *px = r - *py * PW;
  };
  auto prefetch = [&](int t) {
    const int n = t / TILES;
    const int rem = t - n * TILES;
    const int th = rem / TILES_W, tw = rem - th * TILES_W;
#pragma unroll
    for (int j = 0; j < NPF; ++j) {
      const int i = tid + j * 128;
      float v = 0.0f;
      if (i < NELEM) {
        int py, px, c;
        decode(i, &py, &px, &c);
        const int gy = th * TH - 1 + py, gx = tw * TW - 1 + px;
        if (gy >= 0 && gy < HW && gx >= 0 && gx < HW) {
          v = a.x[((static_cast<size_t>(n) * 3 + c) * HW + gy) * HW + gx];
        }
      }
      pf[j] = v;
    }
  };
  if (static_cast<int>(blockIdx.x) < num_tiles) prefetch(blockIdx.x);
  int it = 0;
  for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
    const int n = t / TILES;
    const int rem = t - n * TILES;
    const int th = rem / TILES_W, tw = rem - th * TILES_W;
    const int h0 = th * TH, w0 = tw * TW;
    // ---- 1. publish the prefetched patch as normalised bf16 (0 outside the image), then fetch the next one
#pragma unroll
    for (int j = 0; j < NPF; ++j) {
      const int i = tid + j * 128;
      if (i < NELEM) {
        int py, px, c;
        decode(i, &py, &px, &c);
        const int gy = h0 - 1 + py, gx = w0 - 1 + px;
        const bool ok = gy >= 0 && gy < HW && gx >= 0 && gx < HW;
        s_in[(py * PW + px) * 4 + c] = static_cast<unsigned short>(pack16x2<F16>(ok ? pf[j] : 0.0f, 0.0f) & 0xffffu);
      }
    }
    __syncthreads();
    if (t + static_cast<int>(gridDim.x) < num_tiles) prefetch(t + gridDim.x);
    // ---- 2. this thread's K-major row: per filter row 3 pixels x 4 channels = 24 contiguous bytes of the patch
    {
      const int sw = tid & 7;
#pragma unroll
      for (int kh = 0; kh < 3; ++kh) {
        const uint2* src = reinterpret_cast<const uint2*>(s_in + ((hl + kh) * PW + wl) * 4);
        const uint2 p0 = src[0], p1 = src[1], p2 = src[2];
        *reinterpret_cast<uint4*>(sA + tid * 128 + (((2 * kh) ^ sw) << 4)) = make_uint4(p0.x, p0.y, p1.x, p1.y);
        *reinterpret_cast<uint2*>(sA + tid * 128 + (((2 * kh + 1) ^ sw) << 4)) = p2;      // upper 8 bytes stay zero
      }
    }
    fence_proxy_async_smem();       // generic-proxy smem writes -> visible to the tensor core (async proxy)
    tcgen05_fence_before();
    __syncthreads();
    // ---- 3. three MMAs (K = 48), one issuing thread
    if (tid == 0) {
      tcgen05_fence_after();
      const uint64_t ad = make_kmajor_desc<128>(sA_addr), bd = make_kmajor_desc<128>(sB_addr);
      umma_bf16_ss(tmem, ad, bd, idesc, 0u);
      umma_bf16_ss(tmem, ad + 2, bd + 2, idesc, 1u);
      umma_bf16_ss(tmem, ad + 4, bd + 4, idesc, 1u);
      umma_commit(bar);
    }
    mbar_wait(bar, it & 1);
    tcgen05_fence_after();
    // ---- 4. epilogue: TMEM lane = pixel row; scale/shift/ReLU; 64 contiguous bytes per pixel (2 full sectors)
    uint32_t v[32];
    tmem_ld_32x32(tmem + (static_cast<uint32_t>(warp * 32) << 16), v);
    tmem_ld_wait();
    uint32_t p[16];
#pragma unroll
    for (int c = 0; c < 32; c += 2) {
      const float x0 = fmaf(__uint_as_float(v[c]), a.scale[c], a.shift[c]);
      const float x1 = fmaf(__uint_as_float(v[c + 1]), a.scale[c + 1], a.shift[c + 1]);
      p[c >> 1] = pack16x2_relu<F16>(x0, x1);
    }
    __nv_bfloat16* o = a.out + ((static_cast<size_t>(n) * HW + (h0 + hl)) * HW + (w0 + wl)) * 32;
    st_global_v8(o, p);
    st_global_v8(o + 16, p + 8);
    tcgen05_fence_before();        // TMEM reads done before the next tile's MMAs (ordered by the __syncthreads above)
  }
  __syncthreads();
  if (warp == 0) { tcgen05_fence_after(); tmem_dealloc<32>(tmem); }
}

}  // namespace ff
namespace ff {

// -----------------------------------------------------------------------------------------------------------------
// Feature layer 1 without any im2col: pixel-pair GEMM straight out of the normalised patch.
//
// The patch is kept in shared memory as bf16 [18 rows][18 pixels][4 ch] (8 bytes per pixel, ch 3 = 0, slot i =
// image pixel w0-1+i).  A PAIR of output pixels (w0+2j, w0+2j+1) needs, per filter row kh, the 4 input pixels at
// slots 2j..2j+3 = 32 contiguous bytes = exactly one K=16 MMA step.  With a NON-swizzled K-major descriptor
// (8-row core matrices of 16-byte rows, LBO = 16 B between the two K chunks, SBO = patch row pitch between tile
// rows) the tensor core reads those overlapping windows directly: rows = 8 pairs x 16 tile rows = 128, N = 2 x 32
// (pair-expanded filter B[(p,co)][(q,c)] = W[co][kh][q-p][c]), and the whole 16x16-pixel tile is THREE tcgen05.mma.
// No per-thread operand build; the epilogue writes 128 contiguous bytes per thread.
struct C1PairArgs {
  __nv_bfloat16* out;
  const __nv_bfloat16* w;        // [3 kh][64 (p,co)][16 (q,c)] bf16
  int n_img;
  float na[3], nb[3];
  float scale[32];
  float shift[32];
};

__device__ __forceinline__ uint64_t make_kmajor_desc_noswz(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(lbo_bytes >> 4) << 16;
  d |= static_cast<uint64_t>(sbo_bytes >> 4) << 32;
  d |= 1ull << 46;               // descriptor version; layout_type 0 = SWIZZLE_NONE
  return d;
}

template <bool F16 = false>
__global__ void __launch_bounds__(128, 8)
conv1_pair_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ C1PairArgs a) {
  constexpr int HW = 224, TW = 16, TH = 16, PW = TW + 2, PH = TH + 2;
  constexpr int TILES_W = HW / TW, TILES_H = HW / TH, TILES = TILES_W * TILES_H;
  constexpr int RING = 3, RAW_ROW = 80, RAW_BYTES = PH * RAW_ROW, RAW_SLOT = 1536;
  constexpr int SPITCH = 160;    // bytes per patch row (18 pixels x 8 B = 144, padded to a multiple of 16)
  __shared__ __align__(128) uint8_t s_raw[RING][RAW_SLOT];
  __shared__ __align__(128) uint8_t s_in[PH * SPITCH];
  __shared__ __align__(128) uint8_t sB[3 * 2048];
  __shared__ __align__(8) uint64_t s_bar[1 + RING];
  __shared__ uint32_t s_tmem;

  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t bar_mma = smem_u32(&s_bar[0]);
  const uint32_t bar_raw = smem_u32(&s_bar[1]);
  // filter -> core-matrix layout: (n, chunk c) at ((n/8)*2 + c)*128 + (n%8)*16
  for (int i = tid; i < 3 * 64 * 2; i += 128) {
    const int kh = i / 128, rem = i % 128, n = rem >> 1, c = rem & 1;
    *reinterpret_cast<uint4*>(sB + kh * 2048 + ((n >> 3) * 2 + c) * 128 + (n & 7) * 16) = reinterpret_cast<const uint4*>(a.w)[i];
  }
  for (int i = tid; i < PH * SPITCH / 16; i += 128) reinterpret_cast<uint4*>(s_in)[i] = make_uint4(0, 0, 0, 0);
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
  const uint32_t sin_addr = smem_u32(s_in), sB_addr = smem_u32(sB);
  constexpr uint32_t idesc = make_idesc_16<F16>(128, 64);
  const int num_tiles = TILES * a.n_img;
  const int hl = tid >> 3, jl = tid & 7;

  int job_py[3], job_px[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const int pi = tid + j * 128;
    job_py[j] = pi / PW;
    job_px[j] = pi - job_py[j] * PW;
  }
  auto issue = [&](int t, int slot) {
    const int n = t / TILES;
    const int rem = t - n * TILES;
    const int th = rem / TILES_W, tw = rem - th * TILES_W;
    mbar_arrive_expect_tx(bar_raw + 8 * slot, RAW_BYTES);
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
            smem_u32(&s_raw[slot][0])),
        "l"(reinterpret_cast<uint64_t>(&tmX)), "r"(bar_raw + 8 * slot), "r"(48 * tw - 16), "r"(th * TH - 1), "r"(n)
        : "memory");
  };
  if (tid == 0) {
    for (int s = 0; s < RING - 1; ++s) {
      const int t = blockIdx.x + s * gridDim.x;
      if (t < num_tiles) issue(t, s);
    }
  }
  int it = 0;
  for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
    const int n = t / TILES;
    const int rem = t - n * TILES;
    const int th = rem / TILES_W, tw = rem - th * TILES_W;
    const int h0 = th * TH, w0 = tw * TW;
    const int slot = it % RING;
    // ---- 1. raw uint8 window -> normalised bf16 patch (0 outside the image = padding after normalisation)
    mbar_wait(bar_raw + 8 * slot, (it / RING) & 1);
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const int pi = tid + j * 128;
      if (pi < PH * PW) {
        const uint8_t* rp = &s_raw[slot][job_py[j] * RAW_ROW + 13 + 3 * job_px[j]];   // segment starts 13 B into the window
        const int gy = h0 - 1 + job_py[j], gx = w0 - 1 + job_px[j];
        const bool ok = gy >= 0 && gy < HW && gx >= 0 && gx < HW;
        const float v0 = ok ? fmaf(static_cast<float>(rp[0]), a.na[0], a.nb[0]) : 0.0f;
        const float v1 = ok ? fmaf(static_cast<float>(rp[1]), a.na[1], a.nb[1]) : 0.0f;
        const float v2 = ok ? fmaf(static_cast<float>(rp[2]), a.na[2], a.nb[2]) : 0.0f;
        *reinterpret_cast<uint2*>(s_in + job_py[j] * SPITCH + job_px[j] * 8) = make_uint2(pack16x2<F16>(v0, v1), pack16x2<F16>(v2, 0.0f));
      }
    }
    fence_proxy_async_smem();
    tcgen05_fence_before();
    __syncthreads();
    if (tid == 0) {
      const int tn = t + (RING - 1) * gridDim.x;
      if (tn < num_tiles) issue(tn, (it + RING - 1) % RING);
      // ---- 2. three MMAs: filter row kh reads patch rows (h_l + kh); K = the 4-pixel window of each pair
      tcgen05_fence_after();
#pragma unroll
      for (int kh = 0; kh < 3; ++kh) {
        const uint64_t ad = make_kmajor_desc_noswz(sin_addr + kh * SPITCH, 16, SPITCH);
        const uint64_t bd = make_kmajor_desc_noswz(sB_addr + kh * 2048, 128, 256);
        umma_bf16_ss(tmem, ad, bd, idesc, kh > 0 ? 1u : 0u);
      }
      umma_commit(bar_mma);
    }
    mbar_wait(bar_mma, it & 1);
    tcgen05_fence_after();
    // ---- 3. epilogue: thread = pixel pair (h_l, j): 2 x 32 channels = 128 contiguous bytes
    __nv_bfloat16* o = a.out + ((static_cast<size_t>(n) * HW + (h0 + hl)) * HW + (w0 + 2 * jl)) * 32;
#pragma unroll
    for (int p = 0; p < 2; ++p) {
      uint32_t v[32];
      tmem_ld_32x32(tmem + (static_cast<uint32_t>(warp * 32) << 16) + p * 32, v);
      tmem_ld_wait();
      uint32_t pk[16];
#pragma unroll
      for (int c = 0; c < 32; c += 2) {
        const float x0 = fmaf(__uint_as_float(v[c]), a.scale[c], a.shift[c]);
        const float x1 = fmaf(__uint_as_float(v[c + 1]), a.scale[c + 1], a.shift[c + 1]);
        pk[c >> 1] = pack16x2_relu<F16>(x0, x1);
      }
      st_global_v8(o + p * 32, pk);
      st_global_v8(o + p * 32 + 16, pk + 8);
    }
    tcgen05_fence_before();
  }
  __syncthreads();
  if (warp == 0) { tcgen05_fence_after(); tmem_dealloc<64>(tmem); }
}

}  // namespace ff
