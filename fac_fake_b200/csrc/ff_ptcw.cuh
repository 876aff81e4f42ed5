// ff_ptcw.cuh — persistent implicit-GEMM 3x3 / pad 1 convolution of feature layers 7..9 (Cout = 128 at 56 x 56; reference
// op: nn.Conv2d + BatchNorm2d(eval) + ReLU [+ MaxPool2d(2)], /root/reference/CViT-main/model/cvit.py:110-116), the filter
// as the M operand and 256 pixels as the N operand:  D[cout][pixel] = sum_k W[cout][k] * X[pixel][k].
//
// Why (profiles/r02_ptc_trace.txt).  the round-1/2 kernel for these layers (ptc_conv_kernel<128,2>, since removed) multiplied 128 pixels (M) by 128 channels (N): eight N = 128 MMAs
// per 48 KB k-block.  Its clock64 trace shows ~700 cycles per k-block where the MMAs need 512, and with the TMA loads AND the
// epilogue compiled out the layer still took 0.39 of its 0.41 ms: the N = 128 instruction stream itself is the bound (every
// MMA re-reads 4 KB of A and 4 KB of B = 128 B/clk, the whole shared-memory port), not the operand fill.  The layers that run
// N = 256 reach 87-90 % of the tensor peak.  Cout is only 128 here, but the PIXEL dimension is as long as one likes, so
//   * M = 128 output channels (one filter tile = the A operand), N = 256 pixels (two 8 x 8 x 2-image sub-tiles = the B operand),
//     K = 16: four N = 256 MMAs per k-block = the same 512 tensor cycles from 12 KB instead of 16 KB of operand reads per
//     128 cycles, and half as many instructions;
//   * the accumulator is [128 TMEM lanes = channels][256 columns = pixels], double-buffered in the 512 columns.  The epilogue
//     thread owns ONE channel (its BN scale / shift live in two registers, no shared-memory reads), exchanges values with its
//     lane neighbour so that each thread holds two adjacent channels of one pixel, and stores 4 bytes: the warp's 32 channels
//     of a pixel are one 64-byte run of the NHWC row.  The 2 x 2 max-pool is in-thread (pixels are register indices).
// Operand fill: the three taps (kh, 0..2) of one filter row read the same pixels shifted by one, so ONE TMA box {64 ch,
// BW + 2 pixels, BH rows, BI images} (20 KB) per (kh, 64-channel block) and sub-tile carries the left/right halo once, and
// tap kw is a descriptor that starts kw rows (128 B) into it with a stride-byte-offset of BW + 2 rows between the 8-pixel core
// groups (tcgen05 applies the 128-byte swizzle on absolute address bits: tools/umma_shift_test.cu; same trick as
// ws2conv_kernel).  The two sub-tile boxes are adjacent, 16 groups x 1280 B each, so one descriptor walks all 32 groups.
// Pixels and filter tiles travel in two rings: 40 KB of pixels per three 16 KB filter tiles = 88 KB per 12 MMAs (57 B/clk at
// the full tensor rate instead of 94).
// Roles: warp 0 = TMA producer, warp 1 = TMEM allocation + MMA issue (uniform control flow, one elected lane),
// warps 2..5 = epilogue.
#pragma once
#include "ff_ws.cuh"

namespace ff {

// Developer aid (-DFF_PTC_TRACE, `make trace`, never shipped): block 0 stamps clock64 at the ring hand-offs of tiles 3 and 4.
#ifdef FF_PTC_TRACE
__device__ long long ptc_trace_buf[4][64];
#define PTC_TRACE(ev, i) do { const int i_ = (i); if (blockIdx.x == 0 && i_ >= 0 && i_ < 64) ptc_trace_buf[ev][i_] = clock64(); } while (0)
#else
#define PTC_TRACE(ev, i) do { } while (0)
#endif

template <int NA, int NB>
struct PtcwSmem {
  static constexpr int BOX_BYTES = 10 * 8 * 2 * 128;                  // {64 ch, 10 px, 8 rows, 2 images} bf16
  static constexpr int A_BYTES = 2 * BOX_BYTES;                       // both pixel sub-tiles of one (kh, channel block)
  static constexpr int B_BYTES = 128 * 128;                           // one tap's filter tile: 128 couts x 64 cin
  static constexpr int B_OFF = NA * A_BYTES;
  static constexpr int BAR_OFF = B_OFF + NB * B_BYTES;                // fullA[NA], emptyA[NA], fullB[NB], emptyB[NB], tfull[2], tempty[2]
  static constexpr int SLOT_OFF = BAR_OFF + (2 * NA + 2 * NB + 4) * 8;
  static constexpr int TOTAL = SLOT_OFF + 16 + 1024;
};

template <bool POOL, int NA, int NB, bool F16 = false>
__global__ void __launch_bounds__(192, 1)
ptcw_conv_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TcArgs a) {
  using L = PtcwSmem<NA, NB>;
  constexpr int COUT = 128, NPIX = 256, MSUB = 2, BKE = 64;
  constexpr int BW = 8, LG_BW = 3, LG_BH = 3, LG_BI = 1;
  constexpr int TMEM_COLS = 2 * NPIX;
  constexpr uint32_t SBO = (BW + 2) * 128;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bar_fullA = base + L::BAR_OFF;
  const uint32_t bar_emptyA = bar_fullA + NA * 8;
  const uint32_t bar_fullB = bar_emptyA + NA * 8;
  const uint32_t bar_emptyB = bar_fullB + NB * 8;
  const uint32_t bar_tfull = bar_emptyB + NB * 8;
  const uint32_t bar_tempty = bar_tfull + 16;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(base_ptr + L::SLOT_OFF);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m_tiles = a.tiles_w * a.tiles_h * ((a.n_img + 1) >> LG_BI);
  const int num_tiles = (m_tiles + MSUB - 1) / MSUB;
  const int row_groups = 3 * a.kb_per_tap;                            // (kh, 64-channel block)

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < 2 * NA + 2 * NB; ++s) mbar_init(bar_fullA + 8 * s, 1);
    mbar_init(bar_tfull, 1);
    mbar_init(bar_tfull + 8, 1);
    mbar_init(bar_tempty, 128);
    mbar_init(bar_tempty + 8, 128);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<TMEM_COLS>(smem_u32(const_cast<uint32_t*>(tmem_slot)));
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // sub-tile j of tile t -> first pixel / image; a sub-tile past the end lands on an image index beyond the tensor
  // (TMA zero-fills, the epilogue masks it by n < n_img)
  auto tile_coords = [&](int t, int j, int* w0, int* h0, int* n0) {
    const int mt = t * MSUB + j;
    const int tw = mt % a.tiles_w;
    const int th = (mt / a.tiles_w) % a.tiles_h;
    const int nb = mt / (a.tiles_w * a.tiles_h);
    *w0 = tw << LG_BW;
    *h0 = th << LG_BH;
    *n0 = nb << LG_BI;
  };

  if (warp == 0) {
    if (lane == 0) {
      pdl_trigger();
      pdl_wait();
      int sa = 0, pa = 0, sb = 0, pb = 0;     // ring positions / phases carried across tiles
      int itp = 0;
      (void)itp;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++itp) {
        int w0[MSUB], h0[MSUB], n0[MSUB];
#pragma unroll
        for (int j = 0; j < MSUB; ++j) tile_coords(t, j, &w0[j], &h0[j], &n0[j]);
        for (int rg = 0; rg < row_groups; ++rg) {
          const int kh = rg / a.kb_per_tap;
          const int cc = rg - kh * a.kb_per_tap;
          mbar_wait(bar_emptyA + 8 * sa, pa ^ 1);
          mbar_arrive_expect_tx(bar_fullA + 8 * sa, L::A_BYTES);
#pragma unroll
          for (int j = 0; j < MSUB; ++j)
            tma_load_4d(base + sa * L::A_BYTES + j * L::BOX_BYTES, &tmA, bar_fullA + 8 * sa, cc * BKE, w0[j] - 1, h0[j] + kh - 1, n0[j]);
          if (++sa == NA) { sa = 0; pa ^= 1; }
#pragma unroll 1
          for (int kw = 0; kw < 3; ++kw) {
            mbar_wait(bar_emptyB + 8 * sb, pb ^ 1);
            if (itp == 3 || itp == 4) PTC_TRACE(1, (itp - 3) * 32 + rg * 3 + kw);
            mbar_arrive_expect_tx(bar_fullB + 8 * sb, L::B_BYTES);
            tma_load_2d(base + L::B_OFF + sb * L::B_BYTES, &tmB, bar_fullB + 8 * sb, ((kh * 3 + kw) * a.kb_per_tap + cc) * BKE, 0);
            if (++sb == NB) { sb = 0; pb ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = make_idesc_16<F16>(COUT, NPIX);
    int sa = 0, pa = 0, sb = 0, pb = 0, it = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
      const int acc = it & 1;
      if (it >= 2) mbar_wait(bar_tempty + 8 * acc, ((it >> 1) - 1) & 1);
      tcgen05_fence_after();
      const uint32_t d_tmem = tmem_base + acc * NPIX;
      for (int rg = 0; rg < row_groups; ++rg) {
        mbar_wait(bar_fullA + 8 * sa, pa);
        if (lane == 0 && it == 3) PTC_TRACE(3, rg);
        const uint32_t x_smem = base + sa * L::A_BYTES;
#pragma unroll 1
        for (int kw = 0; kw < 3; ++kw) {
          mbar_wait(bar_fullB + 8 * sb, pb);
          tcgen05_fence_after();
          if (lane == 0 && (it == 3 || it == 4)) PTC_TRACE(0, (it - 3) * 32 + rg * 3 + kw);
          const uint64_t wdesc = make_kmajor_desc<128>(base + L::B_OFF + sb * L::B_BYTES);          // M side: 128 couts
          const uint64_t xdesc = make_kmajor_desc_sbo<128>(x_smem + kw * 128, SBO);                 // N side: 256 pixels, tap kw
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_bf16_ss(d_tmem, wdesc + 2 * k, xdesc + 2 * k, idesc, (rg > 0 || kw > 0 || k > 0) ? 1u : 0u);
            umma_commit(bar_emptyB + 8 * sb);
            if (kw == 2) umma_commit(bar_emptyA + 8 * sa);
          }
          __syncwarp();
          if (++sb == NB) { sb = 0; pb ^= 1; }
        }
        if (++sa == NA) { sa = 0; pa ^= 1; }
      }
      if (elect_one()) umma_commit(bar_tfull + 8 * acc);
      __syncwarp();
    }
  } else {
    const int g = warp & 3;                       // TMEM lane group = output channels 32g .. 32g+31
    const int ch = g * 32 + lane;
    const int odd = lane & 1;
    const float sc = a.scale[ch], sh = a.shift[ch];
    const int OH = POOL ? a.H >> 1 : a.H, OW = POOL ? a.W >> 1 : a.W;
    uint32_t* out = reinterpret_cast<uint32_t*>(reinterpret_cast<__nv_bfloat16*>(a.out) + (ch & ~1));
    int it = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
      const int acc = it & 1;
      mbar_wait(bar_tfull + 8 * acc, (it >> 1) & 1);
      tcgen05_fence_after();
      if (threadIdx.x == 64) PTC_TRACE(2, it);
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(g * 32) << 16) + acc * NPIX;
#pragma unroll 1
      for (int c0 = 0; c0 < NPIX; c0 += 32) {     // 32 columns = 4 rows x 8 pixels of one image of sub-tile c0 / 128
        uint32_t v[32];
        tmem_ld_32x32(taddr + c0, v);
        tmem_ld_wait();
        if (c0 + 32 == NPIX) {                    // accumulator fully read: hand the TMEM buffer back to the MMA warp
          tcgen05_fence_before();
          mbar_arrive(bar_tempty + 8 * acc);
        }
        int w0, h0, n0;
        tile_coords(t, c0 >> 7, &w0, &h0, &n0);
        const int n = n0 + ((c0 >> 6) & 1);
        const int h = h0 + ((c0 >> 3) & 7);
        if (n >= a.n_img) continue;
        float x[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) x[i] = fmaf(__uint_as_float(v[i]), sc, sh);
        if (!POOL) {
          uint32_t* orow = out + ((static_cast<size_t>(a.img_off_out + n) * OH + h) * OW + w0 + odd) * (COUT / 2);
#pragma unroll
          for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int wp = 0; wp < 4; ++wp) {
              // this lane holds channel ch of pixels 2wp, 2wp+1; after the exchange an even lane holds channels (ch, ch+1) of
              // pixel 2wp and an odd lane channels (ch-1, ch) of pixel 2wp+1
              const float lo = x[r * 8 + 2 * wp], hi = x[r * 8 + 2 * wp + 1];
              const float got = __shfl_xor_sync(0xffffffffu, odd ? lo : hi, 1);
              orow[(static_cast<size_t>(r) * OW + 2 * wp) * (COUT / 2)] = odd ? pack16x2_relu<F16>(got, hi) : pack16x2_relu<F16>(lo, got);
            }
        } else {
          uint32_t* orow = out + ((static_cast<size_t>(a.img_off_out + n) * OH + (h >> 1)) * OW + (w0 >> 1) + odd) * (COUT / 2);
#pragma unroll
          for (int rp = 0; rp < 2; ++rp)
#pragma unroll
            for (int q = 0; q < 2; ++q) {
              float m[2];
#pragma unroll
              for (int e = 0; e < 2; ++e) {
                const int i = rp * 16 + (2 * q + e) * 2;
                m[e] = fmaxf(fmaxf(x[i], x[i + 1]), fmaxf(x[i + 8], x[i + 9]));
              }
              const float got = __shfl_xor_sync(0xffffffffu, odd ? m[0] : m[1], 1);
              orow[(static_cast<size_t>(rp) * OW + 2 * q) * (COUT / 2)] = odd ? pack16x2_relu<F16>(got, m[1]) : pack16x2_relu<F16>(m[0], got);
            }
        }
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
#ifdef FF_PTC_TRACE
  if (blockIdx.x == 0 && threadIdx.x == 0 && num_tiles > 6 * static_cast<int>(gridDim.x) && row_groups == 6) {
    const long long t0 = ptc_trace_buf[0][0];
    printf("ptcw trace (Cin 128), block 0, tiles 3 and 4, cycles since the first filter tile of tile 3 was full:\n");
    for (int k = 0; k < 18; ++k)
      printf("  tile 3 (rg %d, kw %d): B slot acquired %7lld  seen full %7lld   | tile 4: acquired %7lld  seen full %7lld\n", k / 3, k % 3,
             ptc_trace_buf[1][k] - t0, ptc_trace_buf[0][k] - t0, ptc_trace_buf[1][32 + k] - t0, ptc_trace_buf[0][32 + k] - t0);
    for (int k = 0; k < 6; ++k) printf("  tile 3 rg %d: pixel boxes seen full at %lld\n", k, ptc_trace_buf[3][k] - t0);
    for (int i = 2; i < 7; ++i) printf("  accumulator of tile %d complete at %lld\n", i, ptc_trace_buf[2][i] - t0);
  }
#endif
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

template <bool POOL, bool F16 = false>
inline cudaError_t launch_ptcw(int grid, cudaStream_t st, const CUtensorMap& a, const CUtensorMap& b, const TcArgs& args) {
  constexpr int NA = 2, NB = 8;   // (3, 6 measured the same)
  return ffh::launch_smem(ptcw_conv_kernel<POOL, NA, NB, F16>, dim3(grid), dim3(192), PtcwSmem<NA, NB>::TOTAL, st, true, a, b, args);
}

}  // namespace ff
