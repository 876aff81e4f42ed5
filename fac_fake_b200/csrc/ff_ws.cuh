// ff_ws.cuh — persistent, weight-stationary, halo-patch 3x3 convolution for the wide-and-shallow layers
// (feature layers 2..6: Cin, Cout in {32, 64} at 224x224 / 112x112; /root/reference/CViT-main/model/cvit.py:91-108).
//
// Why a second conv kernel: with K = 9*Cin <= 576 the per-tap implicit GEMM of ff_tc.cuh re-reads every activation
// 9 times from L2 and pays CTA setup per 128 pixels; measured, those layers were L2->SMEM-bandwidth bound
// (~6.4 TB/s) at 210-360 TFLOP/s.  Here
//   * each CTA is persistent (grid = #SMs) and keeps ALL 9 x [Cout][Cin] filter taps resident in shared memory;
//   * one TMA box {Cin, 10, 18, 1} brings the (8+2) x (16+2) halo patch of a tile of 8 x 16 output pixels ONCE
//     (1.4x read amplification instead of 9x; out-of-bounds rows/cols are zero-filled = the conv padding);
//   * the 9 taps are 9 shared-memory descriptors into that one patch: tap (kh,kw) starts (kh*10 + kw) rows into
//     the patch and uses a stride-byte-offset of 10 rows between the 8-row core groups.  (Measured on B200:
//     tcgen05 applies the 128B/64B swizzle on absolute smem address bits, so descriptors may start at any row
//     and use any 16-byte-multiple SBO with base_offset = 0 — tools/umma_shift_test.cu.)
//   * the accumulator is double-buffered in TMEM so the epilogue of tile i overlaps the MMAs of tile i+1;
//   * the epilogue goes straight from registers to global memory (one output pixel = one thread = one contiguous
//     Cout*2-byte run); the 2x2 max-pool is two warp shuffles (w-neighbour = lane^1, h-neighbour = lane^8).
#pragma once
#include "ff_tc.cuh"

namespace ff {

// K-major descriptor with an explicit stride-byte-offset (bytes between consecutive 8-row groups).
template <int ROWB>
__device__ __forceinline__ uint64_t make_kmajor_desc_sbo(uint32_t smem_addr, uint32_t sbo_bytes) {
  constexpr uint64_t layout = (ROWB == 128) ? 2ull : 4ull;
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= 1ull << 16;
  d |= static_cast<uint64_t>(sbo_bytes >> 4) << 32;
  d |= 1ull << 46;
  d |= layout << 61;
  return d;
}

// Folded BN scale / shift passed by value (kernel-parameter constant bank): the epilogue FMAs take them as
// constant operands instead of shared-memory loads.
struct WsEpi {
  float scale[64];
  float shift[64];
};

// =================================================================================================================
// Pixel-pair formulation for the Cin = 32 layers (feature layers 2, 3, 4).
//
// Measured on B200 (tools/umma_rate_test.cu): one tcgen05.mma M=128,K=16 costs >= 64 cycles whatever N <= 128 is
// (the 128x32-byte A fetch), so with N = Cout = 32 the tensor pipe runs at 25 %.  Two horizontally adjacent pixels
// of a 32-channel NHWC tensor are one contiguous 128-byte row, so the GEMM is re-shaped as
//     rows    = pixel PAIRS (2j, 2j+1)                       M = 128 pairs = a 16 x 16 pixel tile
//     columns = (pixel-in-pair p, cout)                      N = 2*Cout (64 or 128)
//     K       = (kh, window pixel q in 0..3, cin)            K = 3*4*32 = 384, window = pixels 2j-1 .. 2j+2
// with the pair-expanded filter  B[(p,co)][(kh,q,ci)] = W[co][kh][q-p][ci]  (zero unless 0 <= q-p <= 2).
// 24 MMAs per 256 pixels instead of 36: 1.5x fewer tensor cycles at N = 64, 3x fewer at N = 128.
// A is still the single halo patch: TMA box {64 elems = 1 pair, 10 pairs, 18 rows}; the K window of a row starts
// 64 bytes into pair j and runs 256 bytes, i.e. descriptor start = patch + kh*10 rows + 64 + 32*kstep bytes.
// The output row (pair, N values) is 2*Cout contiguous bf16 in NHWC.
template <int BN, int STAGES>
struct Ws2Smem {
  static constexpr int W_TILE = BN * 128;                             // one 64-element k-block of the filter
  static constexpr int W_BYTES = 6 * W_TILE;                          // 3 kh x 2 k-blocks
  static constexpr int PATCH_BYTES = 180 * 128;
  static constexpr int PATCH_STRIDE = (PATCH_BYTES + 1023) / 1024 * 1024;
  static constexpr int W_OFF = 0;
  static constexpr int P_OFF = W_BYTES;
  static constexpr int BAR_OFF = P_OFF + STAGES * PATCH_STRIDE;       // w, full[S], empty[S], tfull[2], tempty[2]
  static constexpr int SLOT_OFF = BAR_OFF + (2 * STAGES + 5) * 8;
  static constexpr int TOTAL = SLOT_OFF + 16 + 1024;
};

// Epilogue for one accumulator row = one pixel pair.  COUT = BN/2.  Processes the pair one pixel (COUT columns) at
// a time to bound registers.  POOL: horizontal max = the two pixels of the pair, vertical max = lane ^ 8.
template <int BN, bool POOL, bool ARRIVE_ON_LEADER = false, bool F16 = false>
__device__ __forceinline__ void ws2_epilogue(uint32_t taddr, const WsEpi& ss, const TcArgs& a, int pw0, int h0, int n, int r,
                                             int lane, uint32_t arrive_bar) {
  constexpr int COUT = BN / 2;
  __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(a.out);
  const int hl = r >> 3, jl = r & 7;
  const int Wp = a.W >> 1;                                            // pairs per image row
  uint32_t pk[2][COUT / 2];
#pragma unroll
  for (int p = 0; p < 2; ++p) {
    uint32_t v[COUT];
#pragma unroll
    for (int c0 = 0; c0 < COUT; c0 += 32) tmem_ld_32x32(taddr + p * COUT + c0, *reinterpret_cast<uint32_t(*)[32]>(&v[c0]));
    tmem_ld_wait();
    if (p == 1) {
      tcgen05_fence_before();
      __syncwarp();                                    // one arrival per warp (leader CTA of a pair, else this CTA)
      if (lane == 0) { if (ARRIVE_ON_LEADER) mbar_arrive_cluster(arrive_bar, 0); else mbar_arrive(arrive_bar); }
    }
#pragma unroll
    for (int c = 0; c < COUT; c += 2) {
      const float x0 = fmaf(__uint_as_float(v[c]), ss.scale[c], ss.shift[c]);
      const float x1 = fmaf(__uint_as_float(v[c + 1]), ss.scale[c + 1], ss.shift[c + 1]);
      pk[p][c >> 1] = pack16x2_relu<F16>(x0, x1);
    }
    if (!POOL && n < a.n_img) {
      if (COUT == 64 && a.out_blocked) {
        // channel-blocked [n][2][H][W][32]: each 32-channel half of this pixel is one 64-byte run
#pragma unroll
        for (int b = 0; b < COUT / 32; ++b) {
          const size_t pix = (static_cast<size_t>((a.img_off_out + n) * 2 + b) * a.H + (h0 + hl)) * a.W + (2 * (pw0 + jl) + p);
          __nv_bfloat16* o = out + pix * 32;
          st_global_v8(o, &pk[p][16 * b]);
          st_global_v8(o + 16, &pk[p][16 * b + 8]);
        }
      } else {
        const size_t pair = (static_cast<size_t>(a.img_off_out + n) * a.H + (h0 + hl)) * Wp + (pw0 + jl);
        __nv_bfloat16* o = out + pair * BN + p * COUT;
#pragma unroll
        for (int i = 0; i < COUT / 16; ++i) st_global_v8(o + i * 16, &pk[p][8 * i]);
      }
    }
  }
  if (POOL) {
#pragma unroll
    for (int i = 0; i < COUT / 2; ++i) {
      const uint32_t mu = max16x2<F16>(pk[0][i], pk[1][i]);
      pk[0][i] = max16x2<F16>(mu, __shfl_xor_sync(0xffffffffu, mu, 8));
    }
    if ((lane & 8) == 0 && n < a.n_img) {
      const size_t pix = (static_cast<size_t>(a.img_off_out + n) * (a.H >> 1) + ((h0 + hl) >> 1)) * Wp + (pw0 + jl);
      __nv_bfloat16* o = out + pix * COUT;
#pragma unroll
      for (int i = 0; i < COUT / 16; ++i) st_global_v8(o + i * 16, &pk[0][8 * i]);
    }
  }
}

// Cout = 64 variant for EIGHT epilogue warps (two per TMEM lane group): this thread owns channels [32*HALF, 32*HALF+32)
// of BOTH pixels of its pair, so the 2x2 max-pool stays inside the thread (+ one shuffle for the row neighbour) and every
// store is one 64-byte run.  ncu on the 4-warp epilogue (profiles/r02_ncu_stage12_before.txt): with 128 columns per
// thread the N = 128 kernels kept the tensor pipe's shared-memory port only 48-50 % busy — the MMA thread waited for
// TMEM to drain.  HALF is a template parameter so that the folded BN constants stay immediate constant-bank operands.
template <int HALF, bool POOL, bool ARRIVE_ON_LEADER, bool F16>
__device__ __forceinline__ void ws2_epilogue_c64_half(uint32_t taddr, const WsEpi& ss, const TcArgs& a, int pw0, int h0, int n, int r,
                                                      int lane, uint32_t arrive_bar) {
  constexpr int COUT = 64, C0 = 32 * HALF;
  __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(a.out);
  const int hl = r >> 3, jl = r & 7;
  const int Wp = a.W >> 1;
  uint32_t pk[2][16];
#pragma unroll
  for (int p = 0; p < 2; ++p) {
    uint32_t v[32];
    tmem_ld_32x32(taddr + p * COUT + C0, v);
    tmem_ld_wait();
    if (p == 1) {
      tcgen05_fence_before();
      // one remote arrival per warp: a cluster-scope release costs a MEMBAR, and 512 of them per tile pair showed up as
      // the top stall of this kernel (ncu: stall_membar 2.0 per issued instruction)
      __syncwarp();
      if (lane == 0) { if (ARRIVE_ON_LEADER) mbar_arrive_cluster(arrive_bar, 0); else mbar_arrive(arrive_bar); }
    }
#pragma unroll
    for (int c = 0; c < 32; c += 2) {
      const float x0 = fmaf(__uint_as_float(v[c]), ss.scale[C0 + c], ss.shift[C0 + c]);
      const float x1 = fmaf(__uint_as_float(v[c + 1]), ss.scale[C0 + c + 1], ss.shift[C0 + c + 1]);
      pk[p][c >> 1] = pack16x2_relu<F16>(x0, x1);
    }
    if (!POOL && !a.out_blocked && n < a.n_img) {
      const size_t pair = (static_cast<size_t>(a.img_off_out + n) * a.H + (h0 + hl)) * Wp + (pw0 + jl);
      __nv_bfloat16* o = out + pair * (2 * COUT) + p * COUT + C0;
      st_global_v8(o, &pk[p][0]);
      st_global_v8(o + 16, &pk[p][8]);
    }
  }
  if (!POOL && a.out_blocked) {
    // Channel-blocked [n][2][H][W][32]: this thread owns 128 contiguous bytes (pixels 2j, 2j+1 of block HALF) = four
    // 32-byte pieces.  Written directly, every store instruction scatters 32 quarter-lines (32 sectors per request:
    // measured 13 % of conv4, 3 % of conv5).  A 4x4 transpose of the pieces across the four lanes that hold neighbouring
    // pairs lets instruction q write piece b of pair 4a+q from lane (a, b): four consecutive lanes fill one 128-byte line.
    uint32_t (*pc)[8] = reinterpret_cast<uint32_t (*)[8]>(&pk[0][0]);     // pc[k] = piece k = pk[k >> 1][8 * (k & 1) ..]
    const int b = lane & 3;
#pragma unroll
    for (int m = 2; m >= 1; m >>= 1) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (q & m) continue;
        const bool up = (b & m) != 0;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const uint32_t send = up ? pc[q][e] : pc[q | m][e];
          const uint32_t recv = __shfl_xor_sync(0xffffffffu, send, m);
          if (up) pc[q][e] = recv; else pc[q | m][e] = recv;
        }
      }
    }
    if (n < a.n_img) {
      // pc[q] now holds piece b of pair (jl & ~3) + q of this row
      const size_t row = (static_cast<size_t>((a.img_off_out + n) * 2 + HALF) * a.H + (h0 + hl)) * a.W;
      __nv_bfloat16* o = out + (row + 2 * (pw0 + (jl & ~3))) * 32 + b * 16;
#pragma unroll
      for (int q = 0; q < 4; ++q) st_global_v8(o + q * 64, pc[q]);
    }
  }
  if (POOL) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const uint32_t mu = max16x2<F16>(pk[0][i], pk[1][i]);
      pk[0][i] = max16x2<F16>(mu, __shfl_xor_sync(0xffffffffu, mu, 8));
    }
    if ((lane & 8) == 0 && n < a.n_img) {
      const size_t pix = (static_cast<size_t>(a.img_off_out + n) * (a.H >> 1) + ((h0 + hl) >> 1)) * Wp + (pw0 + jl);
      __nv_bfloat16* o = out + pix * COUT + C0;
      st_global_v8(o, &pk[0][0]);
      st_global_v8(o + 16, &pk[0][8]);
    }
  }
}

// TcArgs: H, W, tiles_w (= W/16), tiles_h (= H/16), n_img, img_off_out, out.  epi.scale/shift indexed by cout.
template <int BN, bool POOL, int STAGES, bool F16 = false>
__global__ void __launch_bounds__(BN == 128 ? 320 : 192, 1)
ws2conv_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, const TcArgs a,
               const __grid_constant__ WsEpi epi) {
  using L = Ws2Smem<BN, STAGES>;
  constexpr int TMEM_COLS = 2 * BN;       // double-buffered accumulator (128 or 256 columns)
  constexpr int EPI_WARPS = BN == 128 ? 8 : 4;         // N = 128: two epilogue warps per TMEM lane group
  static_assert(BN == 64 || BN == 128, "BN");

  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bar_w = base + L::BAR_OFF;
  const uint32_t bar_full = bar_w + 8;
  const uint32_t bar_empty = bar_full + STAGES * 8;
  const uint32_t bar_tfull = bar_empty + STAGES * 8;
  const uint32_t bar_tempty = bar_tfull + 16;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(base_ptr + L::SLOT_OFF);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int tiles_per_img = a.tiles_w * a.tiles_h;
  const int num_tiles = tiles_per_img * a.n_img;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmW);
    mbar_init(bar_w, 1);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    mbar_init(bar_tfull, 1);
    mbar_init(bar_tfull + 8, 1);
    mbar_init(bar_tempty, EPI_WARPS);           // one arrival per epilogue warp
    mbar_init(bar_tempty + 8, EPI_WARPS);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<TMEM_COLS>(smem_u32(const_cast<uint32_t*>(tmem_slot)));
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(bar_w, L::W_BYTES);
      for (int kb = 0; kb < 6; ++kb) tma_load_2d(base + L::W_OFF + kb * L::W_TILE, &tmW, bar_w, kb * 64, 0);
      pdl_trigger();
      pdl_wait();                  // activations are produced by the previous kernel (filters are static)
      int it = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
        const int s = it % STAGES;
        if (it >= STAGES) mbar_wait(bar_empty + 8 * s, ((it / STAGES) - 1) & 1);
        const int n = t / tiles_per_img;
        const int rem = t - n * tiles_per_img;
        const int th = rem / a.tiles_w, tw = rem - th * a.tiles_w;
        mbar_arrive_expect_tx(bar_full + 8 * s, L::PATCH_BYTES);
        tma_load_4d(base + L::P_OFF + s * L::PATCH_STRIDE, &tmA, bar_full + 8 * s, 0, tw * 8 - 1, th * 16 - 1, n);
      }
    }
  } else if (warp == 1) {
    // whole warp in uniform control flow, one elected lane around the MMAs (keeps the descriptors in uniform registers:
    // under `if (lane == 0)` every UTCHMMA was preceded by R2UR moves and an ELECT loop)
    {
      constexpr uint32_t idesc = make_idesc_16<F16>(128, BN);
      const uint64_t bdesc0 = make_kmajor_desc<128>(base + L::W_OFF);
      mbar_wait(bar_w, 0);
      int it = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
        const int s = it % STAGES;
        const int acc = it & 1;
        if (it >= 2) mbar_wait(bar_tempty + 8 * acc, ((it >> 1) - 1) & 1);
        mbar_wait(bar_full + 8 * s, (it / STAGES) & 1);
        tcgen05_fence_after();
        const uint32_t patch = base + L::P_OFF + s * L::PATCH_STRIDE;
        const uint32_t d_tmem = tmem_base + acc * BN;
        // One descriptor per tile, then constant 64-bit offsets (address field is in 16-byte units): the single issuing
        // thread must not spend more scalar work per MMA than the tensor pipe needs to execute it (tools/umma_rate_test.cu:
        // 53-64 cycles per MMA with ready descriptors, 104 when they are rebuilt between the MMAs).
        const uint64_t adesc0 = make_kmajor_desc_sbo<128>(patch + 64, 10 * 128);
        if (elect_one()) {
#pragma unroll
          for (int kh = 0; kh < 3; ++kh) {
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              // A: window starts 64 bytes (one pixel) into pair j of patch row (h_l + kh); 32 bytes per k-step
              const uint64_t adesc = adesc0 + static_cast<uint64_t>((kh * 10 * 128 + 32 * c) >> 4);
              const uint64_t bdesc = bdesc0 + static_cast<uint64_t>(((kh * 2 + (c >> 2)) * L::W_TILE + 32 * (c & 3)) >> 4);
              umma_bf16_ss(d_tmem, adesc, bdesc, idesc, (kh > 0 || c > 0) ? 1u : 0u);
            }
          }
          umma_commit(bar_empty + 8 * s);
          umma_commit(bar_tfull + 8 * acc);
        }
        __syncwarp();
      }
    }
  } else {
    const int g = warp & 3;
    const int r = g * 32 + lane;
    int it = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
      const int acc = it & 1;
      const int n = t / tiles_per_img;
      const int rem = t - n * tiles_per_img;
      const int th = rem / a.tiles_w, tw = rem - th * a.tiles_w;
      mbar_wait(bar_tfull + 8 * acc, (it >> 1) & 1);
      tcgen05_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(g * 32) << 16) + acc * BN;
      if (BN == 128) {
        if (warp < 6) ws2_epilogue_c64_half<0, POOL, false, F16>(taddr, epi, a, tw * 8, th * 16, n, r, lane, bar_tempty + 8 * acc);
        else ws2_epilogue_c64_half<1, POOL, false, F16>(taddr, epi, a, tw * 8, th * 16, n, r, lane, bar_tempty + 8 * acc);
      } else {
        ws2_epilogue<BN, POOL, false, F16>(taddr, epi, a, tw * 8, th * 16, n, r, lane, bar_tempty + 8 * acc);
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}


// =================================================================================================================
// Pixel-pair formulation for Cin = 64 (feature layers 5, 6) on a CTA PAIR (cta_group::2).
//
// With 64 input channels the pair-expanded filter is [N = 2*64 = 128][K = 2 blocks * 3 * 4 * 32 = 768] = 196 KB and
// no longer fits one SM next to the activation patches.  Two CTAs of a cluster each keep HALF of it (64 of the 128
// (pixel, cout) rows, 96 KB) and issue tcgen05.mma.cta_group::2 with M = 256: each SM multiplies its own 128 pixel
// pairs against the full N = 128 (reading the peer's filter half through the pair datapath), so the MMAs run at
// N = 128 = full tensor rate instead of N = 64.  The input is channel-blocked [n][2][H][W][32] (written that way
// by the producer layer) so that a pixel pair of one 32-channel block is one contiguous 128-byte row.
//   * both CTAs issue their own TMA patch loads (2 blocks x 23 KB) whose bytes are credited to the leader's barrier;
//   * only the leader's thread issues the 48 MMAs per tile pair; tcgen05.commit multicasts stage-free / accumulator-
//     ready to both CTAs; both epilogues report "TMEM drained" to the leader's barrier (256 arrivals).
struct Ws2xSmem {
  static constexpr int W_TILE = 64 * 128;                             // this CTA's 64 filter rows of one k-block
  static constexpr int W_BYTES = 12 * W_TILE;                         // (block, kh, half) k-blocks
  static constexpr int PATCH_ONE = 23552;                             // 180 rows x 128 B, 1024-aligned
  static constexpr int PATCH_BYTES = 2 * 23040;                       // bytes actually transferred per stage
  static constexpr int PATCH_STRIDE = 2 * PATCH_ONE;
  static constexpr int STAGES = 2;
  static constexpr int W_OFF = 0;
  static constexpr int P_OFF = W_BYTES;
  static constexpr int BAR_OFF = P_OFF + STAGES * PATCH_STRIDE;       // w, full[2], empty[2], tfull[2], tempty[2]
  static constexpr int SLOT_OFF = BAR_OFF + 9 * 8;
  static constexpr int TOTAL = SLOT_OFF + 16 + 1024;
};

template <bool POOL, bool F16 = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(320, 1)
ws2x_conv_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, const TcArgs a,
                 const __grid_constant__ WsEpi epi) {
  using L = Ws2xSmem;
  constexpr int BN = 128, STAGES = L::STAGES;
  constexpr int TMEM_COLS = 2 * BN;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bar_w = base + L::BAR_OFF;
  const uint32_t bar_full = bar_w + 8;
  const uint32_t bar_empty = bar_full + STAGES * 8;
  const uint32_t bar_tfull = bar_empty + STAGES * 8;
  const uint32_t bar_tempty = bar_tfull + 16;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(base_ptr + L::SLOT_OFF);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
  const int tiles_per_img = a.tiles_w * a.tiles_h;
  const int num_tiles = tiles_per_img * a.n_img;
  const int num_pairs = (num_tiles + 1) >> 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmW);
    mbar_init(bar_w, 1);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    mbar_init(bar_tfull, 1);
    mbar_init(bar_tfull + 8, 1);
    mbar_init(bar_tempty, 16);         // one arrival per epilogue warp of both CTAs (only the leader's copy is used)
    mbar_init(bar_tempty + 8, 16);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc_2cta<TMEM_COLS>(smem_u32(const_cast<uint32_t*>(tmem_slot)));
  if (warp == 0 && lane == 0) {
    // this CTA's half of the pair-expanded filter: rows [64*rank, 64*rank + 64) of every k-block
    mbar_arrive_expect_tx(bar_w, L::W_BYTES);
    for (int kb = 0; kb < 12; ++kb) tma_load_2d(base + L::W_OFF + kb * L::W_TILE, &tmW, bar_w, kb * 64, 64 * rank);
  }
  tcgen05_fence_before();
  cluster_sync_all();                  // barrier inits + TMEM allocation visible to the peer
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  mbar_wait(bar_w, 0);
  cluster_sync_all();                  // both filter halves have landed before the leader issues any MMA

  auto tile_of = [&](int pair_idx, int* n, int* th, int* tw) {
    const int t = 2 * pair_idx + static_cast<int>(rank);
    *n = t / tiles_per_img;            // t >= num_tiles -> n >= n_img: TMA reads past the valid images, stores are masked
    const int rem = t - *n * tiles_per_img;
    *th = rem / a.tiles_w;
    *tw = rem - *th * a.tiles_w;
  };

  if (warp == 0) {
    if (lane == 0) {
      pdl_trigger();
      pdl_wait();
      int it = 0;
      for (int pi = cluster_id; pi < num_pairs; pi += num_clusters, ++it) {
        const int s = it & 1;
        if (it >= STAGES) mbar_wait(bar_empty + 8 * s, ((it >> 1) - 1) & 1);
        int n, th, tw;
        tile_of(pi, &n, &th, &tw);
        if (leader) mbar_arrive_expect_tx(bar_full + 8 * s, 2 * L::PATCH_BYTES);    // own + peer's bytes
        const uint32_t dst = base + L::P_OFF + s * L::PATCH_STRIDE;
        tma_load_4d_2cta(dst, &tmA, bar_full + 8 * s, 0, tw * 8 - 1, th * 16 - 1, 2 * n);
        tma_load_4d_2cta(dst + L::PATCH_ONE, &tmA, bar_full + 8 * s, 0, tw * 8 - 1, th * 16 - 1, 2 * n + 1);
      }
    }
  } else if (warp == 1) {
    // The whole warp walks the tiles in uniform control flow and elects one lane around the MMAs only: under
    // `if (lane == 0)` ptxas keeps the descriptors in vector registers and wraps every UTCHMMA in an ELECT + R2UR.BROADCAST
    // loop (3 per instruction here; ~150 cycles per MMA in the encoder kernel, ff_xf.cuh).
    if (leader) {
      constexpr uint32_t idesc = make_idesc_16<F16>(256, BN);
      int it = 0;
      for (int pi = cluster_id; pi < num_pairs; pi += num_clusters, ++it) {
        const int s = it & 1;
        const int acc = it & 1;
        if (it >= 2) mbar_wait(bar_tempty + 8 * acc, ((it >> 1) - 1) & 1);
        mbar_wait(bar_full + 8 * s, (it >> 1) & 1);
        tcgen05_fence_after();
        const uint32_t patch = base + L::P_OFF + s * L::PATCH_STRIDE;
        const uint32_t d_tmem = tmem_base + acc * BN;
        if (elect_one()) {
#pragma unroll
          for (int cb = 0; cb < 2; ++cb) {
#pragma unroll
            for (int kh = 0; kh < 3; ++kh) {
#pragma unroll
              for (int c = 0; c < 8; ++c) {
                const uint64_t adesc = make_kmajor_desc_sbo<128>(patch + cb * L::PATCH_ONE + kh * 10 * 128 + 64 + 32 * c, 10 * 128);
                const uint64_t bdesc = make_kmajor_desc<128>(base + L::W_OFF + (cb * 6 + kh * 2 + (c >> 2)) * L::W_TILE) + 2 * (c & 3);
                umma_bf16_ss_2cta(d_tmem, adesc, bdesc, idesc, (cb > 0 || kh > 0 || c > 0) ? 1u : 0u);
              }
            }
          }
          umma_commit_2cta(bar_empty + 8 * s, 0x3);
          umma_commit_2cta(bar_tfull + 8 * acc, 0x3);
        }
        __syncwarp();
      }
    }
  } else {
    const int g = warp & 3;
    const int r = g * 32 + lane;
    int it = 0;
    for (int pi = cluster_id; pi < num_pairs; pi += num_clusters, ++it) {
      const int acc = it & 1;
      int n, th, tw;
      tile_of(pi, &n, &th, &tw);
      mbar_wait(bar_tfull + 8 * acc, (it >> 1) & 1);
      tcgen05_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(g * 32) << 16) + acc * BN;
      if (warp < 6) ws2_epilogue_c64_half<0, POOL, true, F16>(taddr, epi, a, tw * 8, th * 16, n, r, lane, bar_tempty + 8 * acc);
      else ws2_epilogue_c64_half<1, POOL, true, F16>(taddr, epi, a, tw * 8, th * 16, n, r, lane, bar_tempty + 8 * acc);
    }
  }

  tcgen05_fence_before();
  cluster_sync_all();                  // the peer's smem / TMEM stay alive until the leader's last MMA has retired
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc_2cta<TMEM_COLS>(tmem_base);
  }
}


template <int BN, bool POOL, int STAGES, bool F16 = false>
inline cudaError_t launch_ws2(int grid, cudaStream_t st, const CUtensorMap& a, const CUtensorMap& w, const TcArgs& args, const WsEpi& epi) {
  return ffh::launch_smem(ws2conv_kernel<BN, POOL, STAGES, F16>, dim3(grid), dim3(BN == 128 ? 320 : 192), Ws2Smem<BN, STAGES>::TOTAL, st, true, a, w, args, epi);
}
template <bool POOL, bool F16 = false>
inline cudaError_t launch_ws2x(int grid, cudaStream_t st, const CUtensorMap& a, const CUtensorMap& w, const TcArgs& args, const WsEpi& epi) {
  return ffh::launch_smem(ws2x_conv_kernel<POOL, F16>, dim3(grid), dim3(320), Ws2xSmem::TOTAL, st, true, a, w, args, epi);
}

}  // namespace ff
