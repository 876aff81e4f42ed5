// ff_c12.cuh — feature layers 1 AND 2 in one kernel (uint8 crops in, conv2 output out).
// DEFAULT on the uint8 input path (FF_C12=0 selects the separate conv1 / conv2 kernels).  History of this kernel on
// B200, layers 1-6 per 512 crops (separate kernels: 3.05-3.13 ms):
//   128 threads, five phases in sequence                      3.42 ms
//   256 threads (two per accumulator row)                     3.39 ms
//   + software pipeline over tiles, one role per CTA          3.37 ms   <- unchanged: clock64 instrumentation showed the
//                                                                          thread issuing the 24 conv2 MMAs BLOCKED for
//                                                                          ~2200 cycles per tile (busy tensor pipe, two
//                                                                          CTAs per SM) and with it every thread at the
//                                                                          next barrier
//   + dedicated MMA-issuing warp, mbarrier hand-offs          3.11 ms
//   + 5-deep ring of uint8 windows                            3.01-3.05 ms  (separate kernels in the same run: 3.08-3.09)
// i.e. the same time as the separate kernels with 3.2 GB less HBM traffic per step and 8 fewer launches; the sustained
// (power-capped) bench is unchanged at 71.6 k crops/s.  What is left per tile: the conv1 accumulators of strip 1 are
// converted by all rows although only pairs 8, 9 are kept (tcgen05.ld is warp-collective), and 30 MMAs x 53 cycles.
// tools/umma_rate_test.cu: the row-shifted SW128 windows and the overlapping non-swizzled windows cost the same as plain
// operands (53 / 64 cycles per MMA for N = 64 / 128).
//
// Layers 1-3 of the CViT stack are bound by HBM traffic, not by the tensor pipe (profiles/r01_ncu_ws2_kernels.txt,
// DESIGN.md §8): the 224x224x32 bf16 map between conv1 and conv2 is 3.2 MB per crop, written once and read once.
// Here it never leaves the SM.  Per 14-row x 16-pixel conv2 output tile:
//   1. TMA brings the uint8 window (18 rows x 80 B) into shared memory; 128 threads normalise it into the bf16
//      NHWC4 patch conv1 reads (18 rows x 22 pixels), zero outside the image (= conv1's padding);
//   2. conv1 as in conv1_pair_kernel (no im2col: non-swizzled K-major descriptors over overlapping 4-pixel windows),
//      on the 16 x 20-pixel region conv2 needs (its tile + halo): two strips of 8 pixel pairs x 16 rows, 6 MMAs
//      (M=128, N=64, K=16) into two TMEM accumulators;
//   3. epilogue 1: BN + ReLU -> bf16 -> written straight into the 128-byte-swizzled halo patch conv2's A descriptors
//      read ([16 rows][10 pairs][2 px x 32 ch]), with ZEROS where the position lies outside the image (conv2's padding
//      is of conv1's OUTPUT, so it cannot be produced by running conv1 on padded input);
//   4. conv2 exactly as ws2conv_kernel<64>: 24 MMAs over the patch with row-shifted SW128 descriptors;
//   5. epilogue 2: BN + ReLU -> 128 contiguous bytes per pixel pair to global (rows 14, 15 of the M tile are unused).
// All phases run on the CTA's 256 threads one after the other (two threads per accumulator row, one per pixel of the pair); two CTAs per SM overlap each other's tensor and
// CUDA-core phases.  conv1 is recomputed on the halo (16x20 / 14x16 = 1.43x of a cheap layer).
#pragma once
#include "ff_ws.cuh"

namespace ff {

struct C12Args {
  __nv_bfloat16* out;            // conv2 output [n][224][224][32] bf16
  const __nv_bfloat16* w1;       // conv1 pair-expanded filter [3 kh][64 (p,co)][16 (q,c)] (as conv1_pair_kernel)
  int n_img;
  float na[3], nb[3];
  float scale1[32], shift1[32];
  float scale2[32], shift2[32];
};

struct C12Smem {
  static constexpr int W2_BYTES = 6 * 64 * 128;                       // conv2 pair-expanded filter, 6 k-blocks of [64][64]
  static constexpr int PATCH_ROWS = 18;                               // 16 produced + 2 only read by the unused M rows
  static constexpr int PATCH_BYTES = PATCH_ROWS * 10 * 128;
  static constexpr int W2_OFF = 0;
  static constexpr int PATCH_OFF = W2_BYTES;                          // 1024-aligned (49152)
  static constexpr int SIN_PITCH = 192;                               // 22 px x 8 B = 176, padded to a multiple of 16
  static constexpr int SIN_ROWS = 18;
  static constexpr int PATCH_STRIDE = ((PATCH_BYTES + 1023) / 1024) * 1024;
  static constexpr int SIN_OFF = PATCH_OFF + 2 * PATCH_STRIDE;        // two patches: tile k+1 is produced while conv2 reads tile k's
  static constexpr int SIN_BYTES = (SIN_ROWS + 1) * SIN_PITCH;        // +1 row: the last window of the last row reads 16 B past it
  static constexpr int B1_OFF = SIN_OFF + ((SIN_BYTES + 127) / 128) * 128;
  static constexpr int B1_BYTES = 3 * 2048;
  static constexpr int RAW_SLOT = 1536, RING = 5;
  static constexpr int RAW_OFF = B1_OFF + B1_BYTES;
  static constexpr int BAR_OFF = RAW_OFF + RING * RAW_SLOT;           // w2, mma1, mma2[2], raw[RING], sin_ready, patch_ready
  static constexpr int SLOT_OFF = BAR_OFF + (6 + RING) * 8;
  static constexpr int TOTAL = SLOT_OFF + 16 + 1024;
};

constexpr int C12_THREADS = 288;      // 8 worker warps + 1 MMA-issuing warp
// Developer aid (-DFF_C12_TRACE, never shipped): block 0 stamps clock64 at the pipeline hand-offs of tiles 6..9 and prints them.
#ifdef FF_C12_TRACE
__device__ long long c12_trace_buf[10][16];
#define C12_TRACE(ev, it) do { if (blockIdx.x == 0 && (it) >= 6 && (it) < 10) c12_trace_buf[ev][(it)] = clock64(); } while (0)
#else
#define C12_TRACE(ev, it) do { } while (0)
#endif
template <bool F16 = false>
__global__ void __launch_bounds__(C12_THREADS, 2)
c12_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ C12Args a) {
  using L = C12Smem;
  constexpr int HW = 224, TW = 16, TH = 14;
  constexpr int TILES_W = HW / TW, TILES_H = HW / TH, TILES = TILES_W * TILES_H;
  constexpr int RAW_ROW = 80, RAW_BYTES = 18 * RAW_ROW;
  constexpr int PW = 22, PH = 18;                                     // conv1 input patch, pixels x rows

  extern __shared__ uint8_t smem_raw[];
#ifdef FF_C12_TRACE
  if (blockIdx.x == 0 && threadIdx.x == 0) c12_trace_buf[9][0] = clock64();
#endif
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
  uint8_t* s_in = base_ptr + L::SIN_OFF;
  uint8_t* s_b1 = base_ptr + L::B1_OFF;
  uint8_t* s_rawp = base_ptr + L::RAW_OFF;
  uint8_t* s_patch = base_ptr + L::PATCH_OFF;
  const uint32_t bar_w2 = base + L::BAR_OFF;
  const uint32_t bar_mma1 = bar_w2 + 8;
  const uint32_t bar_mma2 = bar_w2 + 16;      // two: one per conv2 accumulator
  const uint32_t bar_raw = bar_w2 + 32;
  const uint32_t bar_sin = bar_raw + 8 * L::RING;       // 8 worker-warp arrivals: conv1's input patch of the next tile is written
  const uint32_t bar_patch = bar_sin + 8;               // 8 worker-warp arrivals: conv2's patch of this tile is written
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(base_ptr + L::SLOT_OFF);

  // 256 threads: thread pair (r, r + 128) shares accumulator row r; `half` picks the pixel of the pair it converts
  const int tid = threadIdx.x, warp = tid >> 5, half = tid >> 7, lane_grp = warp & 3;
  // conv1 filter -> core-matrix layout: (n, chunk c) at ((n/8)*2 + c)*128 + (n%8)*16
  for (int i = tid; i < 3 * 64 * 2; i += C12_THREADS) {
    const int kh = i / 128, rem = i % 128, n = rem >> 1, c = rem & 1;
    *reinterpret_cast<uint4*>(s_b1 + kh * 2048 + ((n >> 3) * 2 + c) * 128 + (n & 7) * 16) = reinterpret_cast<const uint4*>(a.w1)[i];
  }
  for (int i = tid; i < L::SIN_BYTES / 16; i += C12_THREADS) reinterpret_cast<uint4*>(s_in)[i] = make_uint4(0, 0, 0, 0);
  for (int i = tid; i < 2 * L::PATCH_STRIDE / 16; i += C12_THREADS) reinterpret_cast<uint4*>(s_patch)[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmW2);
    mbar_init(bar_w2, 1);
    mbar_init(bar_mma1, 1);
    mbar_init(bar_mma2, 1);
    mbar_init(bar_mma2 + 8, 1);
    for (int s = 0; s < L::RING; ++s) mbar_init(bar_raw + 8 * s, 1);
    mbar_init(bar_sin, 8);            // one arrival per worker warp
    mbar_init(bar_patch, 8);
    fence_mbar_init();
  }
  if (warp == 8) tmem_alloc<256>(smem_u32(const_cast<uint32_t*>(tmem_slot)));
  fence_proxy_async_smem();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tm_c1 = tmem;            // conv1: columns 0..63 = strip 0 (pairs 0..7), 64..127 = strip 1 (pairs 2..9)
  const uint32_t tm_c2 = tmem + 128;      // conv2: two accumulators of 64 columns
  if (tid == 256) {
    mbar_arrive_expect_tx(bar_w2, L::W2_BYTES);
    for (int kb = 0; kb < 6; ++kb) tma_load_2d(base + L::W2_OFF + kb * 8192, &tmW2, bar_w2, kb * 64, 0);
    pdl_trigger();
  }
  pdl_wait();
#ifdef FF_C12_TRACE
  if (blockIdx.x == 0 && threadIdx.x == 0) c12_trace_buf[9][1] = clock64();
#endif
  const uint32_t sin_addr = base + L::SIN_OFF, b1_addr = base + L::B1_OFF, patch_addr = base + L::PATCH_OFF;
  constexpr uint32_t idesc = make_idesc_16<F16>(128, 64);
  // base descriptors, built once: the issuing thread only adds compile-time offsets between MMAs (a single thread that
  // rebuilds descriptors between tcgen05.mma issues at ~104 cycles per MMA instead of 53, tools/umma_rate_test.cu)
  const uint64_t ad1 = make_kmajor_desc_noswz(sin_addr, 16, L::SIN_PITCH);
  const uint64_t bd1 = make_kmajor_desc_noswz(b1_addr, 128, 256);
  const uint64_t ad2 = make_kmajor_desc_sbo<128>(patch_addr + 64, 10 * 128);
  const uint64_t bd2 = make_kmajor_desc<128>(base + L::W2_OFF);
  const int num_tiles = TILES * a.n_img;
  const int hl = (tid & 127) >> 3, jl = tid & 7;          // workers only (tid < 256)

  auto issue = [&](int t, int slot) {
    const int n = t / TILES;
    const int rem = t - n * TILES;
    const int th = rem / TILES_W, tw = rem - th * TILES_W;
    mbar_arrive_expect_tx(bar_raw + 8 * slot, RAW_BYTES);
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
            smem_u32(s_rawp + slot * L::RAW_SLOT)),
        "l"(reinterpret_cast<uint64_t>(&tmX)), "r"(bar_raw + 8 * slot), "r"(48 * tw - 16), "r"(th * TH - 2), "r"(n)
        : "memory");
  };
  if (tid == 256)
    for (int s = 0; s < L::RING - 1; ++s) {
      const int t = blockIdx.x + s * gridDim.x;
      if (t < num_tiles) issue(t, s);
    }
  // Pipeline over this CTA's tiles k = 0, 1, ...   (a thread blocks while it issues tcgen05.mma into a busy tensor pipe —
  // ~2200 cycles for the 24 conv2 MMAs when both CTAs of the SM are issuing — so the issuer is a warp of its own):
  //   issuer (warp 8, one lane):  wait sin_ready(k) -> conv1 MMAs(k);   wait patch_ready(k) -> conv2 MMAs(k)
  //   workers (warps 0-7):        a. wait conv1(k), epilogue 1 -> patch[k & 1], arrive patch_ready
  //                               c. convert tile k+1 -> s_in, arrive sin_ready
  //                               d. wait conv2(k-1), epilogue 2 of tile k-1 -> global
  auto tile_coords = [&](int t, int* n, int* h0, int* w0) {
    *n = t / TILES;
    const int rem = t - *n * TILES;
    const int th = rem / TILES_W;
    *h0 = th * TH;
    *w0 = (rem - th * TILES_W) * TW;
  };

  if (warp == 8) {
    // The issuing warp runs in uniform control flow and elects one lane around the TMA / MMA instructions only: inside
    // `if (tid == 256)` every UTCHMMA was preceded by 2-4 R2UR moves of its descriptors plus an ELECT loop.
    {
      // conv1 of tile k on 16 rows x 10 pairs: strip 0 = pairs 0..7, strip 1 = pairs 2..9 (32 bytes further)
      auto issue_conv1 = [&](int t, int it) {
        mbar_wait(bar_sin, it & 1);          // all 256 workers: s_in(k) written — and, by their program order, epilogue 1
        tcgen05_fence_after();               // of tile k-1 finished (the conv1 accumulators are free again)
        if (tid == 256) C12_TRACE(0, it);
        const int tn = t + (L::RING - 1) * gridDim.x;      // every worker has consumed the raw window of tile k
        if (elect_one()) {
          if (tn < num_tiles) issue(tn, (it + L::RING - 1) % L::RING);
#pragma unroll
          for (int strip = 0; strip < 2; ++strip)
#pragma unroll
            for (int kh = 0; kh < 3; ++kh)     // descriptors = per-kernel constants + compile-time offsets (16-byte units)
              umma_bf16_ss(tm_c1 + strip * 64, ad1 + static_cast<uint64_t>((kh * L::SIN_PITCH + strip * 32) >> 4),
                           bd1 + static_cast<uint64_t>((kh * 2048) >> 4), idesc, kh > 0 ? 1u : 0u);
          umma_commit(bar_mma1);
        }
        __syncwarp();
        if (tid == 256) C12_TRACE(1, it);
      };
      // Order in the tensor pipe: conv1(0), [conv1(k+1), conv2(k)] ...  conv1(k+1) goes in FRONT of conv2(k) so that the
      // workers' epilogue 1 of tile k+1 runs while the pipe executes the 24 conv2 MMAs of tile k (before, conv1(k+1) was
      // queued behind conv2(k) and the pipe idled through every epilogue 1).
      bool w2_ready = false;
      int it = 0;
      if (static_cast<int>(blockIdx.x) < num_tiles) issue_conv1(blockIdx.x, 0);
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
        // patch(k) is waited for first (it completes before sin(k+1) in every worker's program order): no barrier
        // phase can then run ahead of this warp
        mbar_wait(bar_patch, it & 1);
        if (tid == 256) C12_TRACE(2, it);
        if (t + static_cast<int>(gridDim.x) < num_tiles) issue_conv1(t + gridDim.x, it + 1);
        // conv2 of tile k: as ws2conv_kernel<64> over patch[k & 1] into accumulator c2[k & 1]
        if (!w2_ready) { mbar_wait(bar_w2, 0); w2_ready = true; }
        tcgen05_fence_after();
        const uint64_t adesc0 = ad2 + static_cast<uint64_t>(((it & 1) * L::PATCH_STRIDE) >> 4);
        const uint32_t d2 = tm_c2 + (it & 1) * 64;
        if (elect_one()) {
#pragma unroll
          for (int kh = 0; kh < 3; ++kh) {
#pragma unroll
            for (int c = 0; c < 8; ++c)
              umma_bf16_ss(d2, adesc0 + static_cast<uint64_t>(((kh * 10 * 128 + 32 * c) >> 4)),
                           bd2 + static_cast<uint64_t>(((kh * 2 + (c >> 2)) * 8192 + 32 * (c & 3)) >> 4), idesc, (kh > 0 || c > 0) ? 1u : 0u);
          }
          umma_commit(bar_mma2 + 8 * (it & 1));
        }
        __syncwarp();
        if (tid == 256) C12_TRACE(3, it);
        __syncwarp();
      }
    }
  } else {
    auto convert = [&](int t, int it) {        // raw uint8 window -> normalised bf16 NHWC4 patch; pixel (py, px) = image (h0-2+py, w0-3+px)
      int n, h0, w0;
      tile_coords(t, &n, &h0, &w0);
      const int slot = it % L::RING;
      mbar_wait(bar_raw + 8 * slot, (it / L::RING) & 1);
      for (int pi = tid; pi < PH * PW; pi += 256) {
        const int py = pi / PW, px = pi - py * PW;
        const uint8_t* rp = s_rawp + slot * L::RAW_SLOT + py * RAW_ROW + 7 + 3 * px;     // the window starts 7 bytes before pixel w0-3
        const int gy = h0 - 2 + py, gx = w0 - 3 + px;
        const bool ok = gy >= 0 && gy < HW && gx >= 0 && gx < HW;
        const float v0 = ok ? fmaf(static_cast<float>(rp[0]), a.na[0], a.nb[0]) : 0.0f;
        const float v1 = ok ? fmaf(static_cast<float>(rp[1]), a.na[1], a.nb[1]) : 0.0f;
        const float v2 = ok ? fmaf(static_cast<float>(rp[2]), a.na[2], a.nb[2]) : 0.0f;
        *reinterpret_cast<uint2*>(s_in + py * L::SIN_PITCH + px * 8) = make_uint2(pack16x2<F16>(v0, v1), pack16x2<F16>(v2, 0.0f));
      }
      fence_proxy_async_smem();
      tcgen05_fence_before();
      __syncwarp();
      if ((tid & 31) == 0) mbar_arrive(bar_sin);
    };
    // Output epilogue of tile `it`: done by the FOUR warps whose half equals the tile's parity (= its accumulator), each
    // thread converting BOTH pixels of pair (hl, jl) = 128 contiguous bytes.  With one pixel per thread (the two halves
    // of a pair in different warps) every store instruction scattered 32 quarter-lines: a timing experiment with one
    // contiguous KB per instruction made the kernel 13 % faster.  Now the four 32-byte pieces are transposed across the
    // four lanes that hold neighbouring pairs, and instruction q writes piece b of pair 4a+q from lane (a, b): four
    // consecutive lanes fill one 128-byte line.  The other four warps do the next tile; every warp still finishes its
    // share of tile k-2 before it arrives on patch_ready(k), which is what frees accumulator k & 1 for the issuer.
    auto epilogue2 = [&](int t, int it) {
      if (half != (it & 1)) return;
      int n, h0, w0;
      tile_coords(t, &n, &h0, &w0);
      mbar_wait(bar_mma2 + 8 * (it & 1), (it >> 1) & 1);
      tcgen05_fence_after();
      if ((tid & 127) == 0) C12_TRACE(7, it);
      uint32_t pk[2][16];
#pragma unroll
      for (int p = 0; p < 2; ++p) {
        uint32_t v[32];
        tmem_ld_32x32(tm_c2 + (it & 1) * 64 + (static_cast<uint32_t>(lane_grp * 32) << 16) + p * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 32; c += 2) {
          const float x0 = fmaf(__uint_as_float(v[c]), a.scale2[c], a.shift2[c]);
          const float x1 = fmaf(__uint_as_float(v[c + 1]), a.scale2[c + 1], a.shift2[c + 1]);
          pk[p][c >> 1] = pack16x2_relu<F16>(x0, x1);
        }
      }
      uint32_t (*pc)[8] = reinterpret_cast<uint32_t (*)[8]>(&pk[0][0]);     // pc[k] = 32-byte piece k of the pair's 128 bytes
      const int b4 = tid & 3;
#pragma unroll
      for (int m = 2; m >= 1; m >>= 1) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (q & m) continue;
          const bool up = (b4 & m) != 0;
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const uint32_t send = up ? pc[q][e] : pc[q | m][e];
            const uint32_t recv = __shfl_xor_sync(0xffffffffu, send, m);
            if (up) pc[q][e] = recv; else pc[q | m][e] = recv;
          }
        }
      }
      if (hl < TH) {      // pc[q] = piece b4 of pair (jl & ~3) + q of row hl
        __nv_bfloat16* o = a.out + ((static_cast<size_t>(n) * HW + (h0 + hl)) * HW + (w0 + 2 * (jl & ~3))) * 32 + b4 * 16;
#pragma unroll
        for (int q = 0; q < 4; ++q) st_global_v8(o + q * 64, pc[q]);
      }
      tcgen05_fence_before();
    };

    if (blockIdx.x < num_tiles) convert(blockIdx.x, 0);
    int it = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
      int n, h0, w0;
      tile_coords(t, &n, &h0, &w0);
      uint8_t* patch = s_patch + (it & 1) * L::PATCH_STRIDE;
      // ---- a. epilogue 1: thread (hl, jl, half) converts pixel `half` of conv1 pair (row hl, pair jl) from strip 0 and of
      //         (row hl, pair jl+2) from strip 1 (kept only for jl >= 6).  Patch row hl = image row h0-1+hl, pair j = image
      //         pixels w0-2+2j, +1; positions outside the image are conv2's zero padding.
      mbar_wait(bar_mma1, it & 1);
      tcgen05_fence_after();
      if (tid == 0) C12_TRACE(4, it);
      {
        const int gy = h0 - 1 + hl;
#pragma unroll
        for (int strip = 0; strip < 2; ++strip) {
          const bool keep = strip == 0 || jl >= 6;     // tcgen05.ld is warp-collective: every thread loads, few store
          const int pj = jl + 2 * strip;
          const int gx = w0 - 2 + 2 * pj;
          const bool inside = gy >= 0 && gy < HW && gx >= 0 && gx < HW;     // pairs never straddle the image border
          const int rowidx = hl * 10 + pj;
          uint8_t* prow = patch + rowidx * 128;
          const int sw = rowidx & 7;
          uint32_t v[32];
          tmem_ld_32x32(tm_c1 + (static_cast<uint32_t>(lane_grp * 32) << 16) + strip * 64 + half * 32, v);
          tmem_ld_wait();
          uint32_t pk[16];
#pragma unroll
          for (int c = 0; c < 32; c += 2) {
            const float x0 = fmaf(__uint_as_float(v[c]), a.scale1[c], a.shift1[c]);
            const float x1 = fmaf(__uint_as_float(v[c + 1]), a.scale1[c + 1], a.shift1[c + 1]);
            pk[c >> 1] = inside ? pack16x2_relu<F16>(x0, x1) : 0u;
          }
          if (keep) {
#pragma unroll
            for (int q = 0; q < 4; ++q)      // pixel `half` of the pair = 16-byte chunks 4*half .. 4*half+3 of the 128-byte row
              *reinterpret_cast<uint4*>(prow + (((4 * half + q) ^ sw) << 4)) = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
          }
        }
      }
      fence_proxy_async_smem();
      tcgen05_fence_before();
      __syncwarp();
      if ((tid & 31) == 0) mbar_arrive(bar_patch);
      if (tid == 0) C12_TRACE(5, it);
      // ---- c. next tile: conversion (s_in was released by the conv1 MMAs of tile k, waited for in step a)
      const int tn = t + gridDim.x;
      if (tn < num_tiles) convert(tn, it + 1);
      if (tid == 0) C12_TRACE(6, it);
      // ---- d. previous tile: output epilogue while the issuer feeds the conv2 MMAs of tile k
      if (it >= 1) epilogue2(t - gridDim.x, it - 1);
      if (tid == 0) C12_TRACE(8, it);
    }
    if (it >= 1) epilogue2(blockIdx.x + (it - 1) * gridDim.x, it - 1);
  }
  __syncthreads();
#ifdef FF_C12_TRACE
  if (blockIdx.x == 0 && tid == 0 && num_tiles > 10 * static_cast<int>(gridDim.x)) {
    printf("c12 block 0: %d tiles; start -> after pdl_wait %lld cycles; -> patch(6) ready %lld; whole kernel %lld cycles\n",
           (num_tiles - 1 - static_cast<int>(blockIdx.x)) / static_cast<int>(gridDim.x) + 1, c12_trace_buf[9][1] - c12_trace_buf[9][0],
           c12_trace_buf[2][6] - c12_trace_buf[9][0], clock64() - c12_trace_buf[9][0]);
    const long long t0 = c12_trace_buf[2][6];
    const char* nm[9] = {"issuer: s_in(k) ready", "issuer: conv1(k) issued", "issuer: patch(k) ready", "issuer: conv2(k) issued",
                         "worker: conv1(k) done", "worker: epilogue1(k) done", "worker: convert(k+1) done", "worker: conv2(k) done (seen in it k+1)",
                         "worker: epilogue2(k-1) done"};
    printf("c12 trace, block 0, cycles since patch(6) was ready at the issuer:\n");
    for (int e = 0; e < 9; ++e) printf("  %-40s k=6: %7lld  k=7: %7lld  k=8: %7lld  k=9: %7lld\n", nm[e], c12_trace_buf[e][6] - t0,
                                       c12_trace_buf[e][7] - t0, c12_trace_buf[e][8] - t0, c12_trace_buf[e][9] - t0);
  }
#endif
  if (warp == 8) { tcgen05_fence_after(); tmem_dealloc<256>(tmem); }
}

template <bool F16 = false>
inline cudaError_t launch_c12(int grid, cudaStream_t st, const CUtensorMap& x, const CUtensorMap& w2, const C12Args& args) {
  return ffh::launch_smem(c12_kernel<F16>, dim3(grid), dim3(C12_THREADS), C12Smem::TOTAL, st, true, x, w2, args);
}

}  // namespace ff
