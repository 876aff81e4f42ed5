// ff_tc.cuh — the two generic tensor-core kernels of the engine:
//   * tc_gemm_kernel: y = x W^T (+bias, activation, residual) for the patch embedding, the per-op encoder linears
//     (fallback when the one-launch encoder cannot be co-resident) and the MLP head (cvit.py:26-28,40-41,155,161-165):
//     one 128 x BN tile per CTA, fp32 accumulator in TMEM, 64-element (SW128) k-blocks.
//       warp 0   TMA producer (A: 2-D box of the activation rows, B: 2-D box {64, BN} of the [out][in] weights)
//       warp 1   TMEM alloc + single-thread tcgen05.mma issue (cta_group::1, kind::f16, M=128, N=BN, K=16),
//                tcgen05.commit releases smem stages / signals the epilogue
//       warps 2-5 epilogue: tcgen05.ld 32 lanes x 32 columns, bias / activation / residual, 32-byte global stores
//   * ptc_conv_kernel: persistent implicit-GEMM 3x3 / pad 1 convolution over NHWC bf16 activations (reference op:
//     nn.Conv2d + BatchNorm2d(eval) + ReLU [+ MaxPool2d(2)], /root/reference/CViT-main/model/cvit.py:110-119).
#pragma once
#include "ff_host.h"
#include "ff_ptx.cuh"

namespace ff {

enum { EPI_STORE_F32 = 0, EPI_STORE_BF16 = 1, EPI_RESID_F32 = 2 };
enum { ACT_NONE = 0, ACT_RELU = 1, ACT_GELU = 2 };

struct TcArgs {
  // ---- conv geometry
  int H, W;               // conv input == output spatial size (before pooling)
  int tiles_w, tiles_h;   // M-tiles per image along w / h
  int lg_bw, lg_bh;       // log2 of the tile box (BW x BH pixels x BI images = 128 rows)
  int n_img;              // valid images in this launch
  int img_off_out;        // image offset added on output (placement inside a larger buffer)
  int cout;               // total output channels (row pitch of the NHWC output)
  int out_blocked;        // pair kernels: write the output channel-blocked [n][C/32][H][W][32] (feeds ws2x)
  int taps;               // conv: 0/9 = 3x3 (pad 1), 1 = 1x1 (pad 0)
  int stride;             // conv: 0/1 = stride 1, 2 = stride 2 (the tensor map carries elementStrides = 2)
  int conv_act;           // conv epilogue: 0 = ReLU (CViT layers), 1 = none (ResNet downsample / channel conv)
  const void* resid;      // rvk_conv2_kernel: bf16 residual tensor (non-null selects the residual epilogue)
  int kb_per_tap;         // Cin / (channels per k-block)
  int cin;                // input channels
  // ---- gemm geometry
  int M, N;               // valid rows / columns
  int ldo;                // row pitch of the outputs (elements)
  // ---- k range
  int kb_total;           // number of k-blocks of the whole reduction
  int kb_per_split;       // k-blocks handled by one blockIdx.z
  // ---- epilogue
  const float* scale;     // conv: per-channel scale (gamma / sqrt(var+eps))
  const float* shift;     // conv: per-channel shift;  gemm: bias or nullptr
  void* out;              // conv: bf16 NHWC;  gemm: see epi
  long long split_stride; // gemm split-K: element offset of split z's private output slab
  int epi;                // gemm: EPI_*
  int act;                // gemm: ACT_*
};

template <int ROWB, int BN, int STAGES>
struct TcSmem {
  static constexpr int A_BYTES = 128 * ROWB;
  static constexpr int B_BYTES = BN * ROWB;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int PIPE_BYTES = STAGES * STAGE_BYTES;
  static constexpr int STAGING_BYTES = 128 * BN * 2;
  static constexpr int MAIN_BYTES = PIPE_BYTES > STAGING_BYTES ? PIPE_BYTES : STAGING_BYTES;
  static constexpr int SS_OFF = MAIN_BYTES;                 // bias floats [BN]
  static constexpr int BAR_OFF = SS_OFF + 2 * BN * 4;       // full[S], empty[S], tmem_full
  static constexpr int SLOT_OFF = BAR_OFF + (2 * STAGES + 1) * 8;
  static constexpr int TOTAL = SLOT_OFF + 16 + 1024;        // + manual 1024-B alignment slack
};

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

// KPS = k-blocks per stage (one full/empty barrier pair per stage): a wait + commit per FOUR MMAs left the tensor pipe dry
// between k-blocks when one CTA owns the SM (ff_xf.cuh, profiles/r02_xf_trace.txt); with KPS = 2 it is eight.
template <int BN, int STAGES, int KPS = 1>
__global__ void __launch_bounds__(192, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TcArgs a) {
  constexpr int ROWB = 128;
  using L = TcSmem<ROWB, BN, STAGES * KPS>;      // STAGES * KPS k-block slots; barrier pair s guards slots s*KPS .. s*KPS+KPS-1
  constexpr int BKE = ROWB / 2;        // bf16 elements per k-block row
  constexpr int KSTEPS = ROWB / 32;    // UMMA K=16 steps per k-block
  constexpr int TMEM_COLS = BN;
  static_assert(BN == 64 || BN == 128, "BN");

  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
  float* ss = reinterpret_cast<float*>(base_ptr + L::SS_OFF);
  const uint32_t bar_full = base + L::BAR_OFF;
  const uint32_t bar_empty = bar_full + STAGES * 8;
  const uint32_t bar_tmem = bar_empty + STAGES * 8;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(base_ptr + L::SLOT_OFF);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int m0 = blockIdx.x * 128;
  const int col0 = blockIdx.y * BN;   // first output channel / column of this CTA
  const int kb_begin = blockIdx.z * a.kb_per_split;
  const int kb_end = min(a.kb_total, kb_begin + a.kb_per_split);

  // ---- one-time setup
  if (warp == 0 && lane == 0) {
    pdl_trigger();   // <= 2 waves of CTAs: let the next (small) kernel pre-stage
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    mbar_init(bar_tmem, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc<TMEM_COLS>(smem_u32(const_cast<uint32_t*>(tmem_slot)));
  }
  if (warp >= 2) {
    const int t = threadIdx.x - 64;
    for (int i = t; i < BN; i += 128) {
      const int c = col0 + i;
      ss[i] = (a.shift != nullptr && c < a.N) ? a.shift[c] : 0.0f;
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // =========================== TMA producer ===========================
    if (lane == 0) {
      pdl_wait();                  // A (activations) is produced by the previous kernel
      int it = 0, s = 0;
      for (int kb = kb_begin; kb < kb_end; kb += KPS) {
        if (it > 0) mbar_wait(bar_empty + 8 * s, (it - 1) & 1);
        const int nk = min(KPS, kb_end - kb);
        const uint32_t bar = bar_full + 8 * s;
        mbar_arrive_expect_tx(bar, nk * L::STAGE_BYTES);
        for (int sub = 0; sub < nk; ++sub) {
          const uint32_t sa = base + (s * KPS + sub) * L::STAGE_BYTES;
          tma_load_2d(sa, &tmA, bar, (kb + sub) * BKE, m0);
          tma_load_2d(sa + L::A_BYTES, &tmB, bar, (kb + sub) * BKE, col0);
        }
        if (++s == STAGES) { s = 0; ++it; }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    // whole warp in uniform control flow, one elected lane around the MMAs; the next stage's barrier is polled early
    {
      constexpr uint32_t idesc = make_idesc_bf16(128, BN);
      int it = 0, s = 0;
      bool ready = kb_begin < kb_end && mbar_test_wait(bar_full, 0);
      for (int kb = kb_begin; kb < kb_end; kb += KPS) {
        if (!ready) mbar_wait(bar_full + 8 * s, it & 1);
        const int sn = s + 1 == STAGES ? 0 : s + 1;
        const int itn = s + 1 == STAGES ? it + 1 : it;
        ready = kb + KPS < kb_end && mbar_test_wait(bar_full + 8 * sn, itn & 1);
        tcgen05_fence_after();
        const int nk = min(KPS, kb_end - kb);
        if (elect_one()) {
          for (int sub = 0; sub < nk; ++sub) {
            const uint32_t sa = base + (s * KPS + sub) * L::STAGE_BYTES;
            const uint64_t adesc = make_kmajor_desc<ROWB>(sa);
            const uint64_t bdesc = make_kmajor_desc<ROWB>(sa + L::A_BYTES);
#pragma unroll
            for (int k = 0; k < KSTEPS; ++k) {
              // advance 32 bytes (16 bf16) along K inside the swizzle atom: +2 in the 16-byte address field
              umma_bf16_ss(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, (kb > kb_begin || sub > 0 || k > 0) ? 1u : 0u);
            }
          }
          umma_commit(bar_empty + 8 * s);   // frees this smem stage once the MMAs above have read it
        }
        __syncwarp();
        s = sn; it = itn;
      }
      if (elect_one()) umma_commit(bar_tmem);              // accumulator complete
      __syncwarp();
    }
  } else {
    // =========================== epilogue (warps 2..5) ===========================
    const int g = warp & 3;               // TMEM lane group this warp may access
    const int r = g * 32 + lane;          // accumulator row == TMEM lane
    mbar_wait(bar_tmem, 0);
    tcgen05_fence_after();
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(g * 32) << 16);

    {
      // ---- GEMM epilogue: direct stores, one accumulator row per thread
      const int m = m0 + r;
      const bool row_ok = m < a.M;
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t v[32];
        tmem_ld_32x32(taddr + c0, v);
        tmem_ld_wait();
        const int nb = col0 + c0;
        if (row_ok && nb < a.N) {   // N is a multiple of 32 for every linear on the path
          float f[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            float x = __uint_as_float(v[i]) + ss[c0 + i];
            if (a.act == ACT_RELU) x = fmaxf(x, 0.0f);
            else if (a.act == ACT_GELU) x = gelu_erf(x);
            f[i] = x;
          }
          const size_t off = static_cast<size_t>(blockIdx.z) * a.split_stride + static_cast<size_t>(m) * a.ldo + nb;
          if (a.epi == EPI_STORE_F32) {
            float* o = reinterpret_cast<float*>(a.out) + off;
#pragma unroll
            for (int i = 0; i < 4; ++i) st_global_v8(o + 8 * i, reinterpret_cast<const uint32_t*>(&f[8 * i]));
          } else if (a.epi == EPI_STORE_BF16) {
            __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(a.out) + off;
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              uint32_t pk[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) pk[e] = pack_bf16x2(f[16 * i + 2 * e], f[16 * i + 2 * e + 1]);
              st_global_v8(o + 16 * i, pk);
            }
          } else if (a.epi == EPI_RESID_F32) {
            float* o = reinterpret_cast<float*>(a.out) + off;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              uint32_t t[8];
              ld_global_v8(o + 8 * i, t);
#pragma unroll
              for (int e = 0; e < 8; ++e) t[e] = __float_as_uint(__uint_as_float(t[e]) + f[8 * i + e]);
              st_global_v8(o + 8 * i, t);
            }
          }
        }
      }
    }
  }

  // ---- teardown
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

}  // namespace ff

namespace ff {

// =================================================================================================================
// Persistent variant of the implicit-GEMM conv (feature layers 7..17): grid = #SMs, each CTA walks a static
// round-robin list of (pixel tile, channel tile) pairs.  The smem ring keeps streaming across tile boundaries,
// the fp32 accumulator is double-buffered in TMEM (2 x BN columns) so the epilogue of tile i overlaps the MMAs
// of tile i+1, and the epilogue goes straight from registers to global memory: one thread = one output pixel =
// one contiguous run of BN bf16; the 2x2 max-pool is two warp shuffles (w-neighbour lane^1, h-neighbour lane^BW).
template <int BN, int MSUB, int STAGES>
struct PtcSmem {
  static constexpr int A_BYTES = MSUB * 128 * 128;                    // MSUB pixel sub-tiles share one B tile
  static constexpr int B_BYTES = BN * 128;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int SS_OFF = STAGES * STAGE_BYTES;                 // scale/shift floats [2][512]
  static constexpr int BAR_OFF = SS_OFF + 2 * 512 * 4;                // full[S], empty[S], tfull[2], tempty[2]
  static constexpr int SLOT_OFF = BAR_OFF + (2 * STAGES + 4) * 8;
  static constexpr int TOTAL = SLOT_OFF + 16 + 1024;
};

// MSUB = 2 (used for Cout = 128): one CTA tile is two 128-pixel sub-tiles against the same 128 output channels,
// i.e. 48 KB of operands per 8 MMAs — the same bytes-per-MMA-cycle ratio as the 128 x 256 tile (with a single
// 128 x 128 tile the smem ring could not be refilled at the rate the N=128 MMAs drain it).
template <int BN, int MSUB, bool POOL, int STAGES, bool F16 = false>
__global__ void __launch_bounds__(192, 1)
ptc_conv_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TcArgs a) {
  using L = PtcSmem<BN, MSUB, STAGES>;
  constexpr int BKE = 64;
  constexpr int TMEM_COLS = 2 * MSUB * BN;
  static_assert(TMEM_COLS == 256 || TMEM_COLS == 512, "TMEM columns");

  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
  float* ss = reinterpret_cast<float*>(base_ptr + L::SS_OFF);
  const uint32_t bar_full = base + L::BAR_OFF;
  const uint32_t bar_empty = bar_full + STAGES * 8;
  const uint32_t bar_tfull = bar_empty + STAGES * 8;
  const uint32_t bar_tempty = bar_tfull + 16;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(base_ptr + L::SLOT_OFF);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_tiles = a.cout / BN;
  const int lg_bi = 7 - a.lg_bw - a.lg_bh;
  const int m_tiles = a.tiles_w * a.tiles_h * ((a.n_img + (1 << lg_bi) - 1) >> lg_bi);
  const int num_tiles = ((m_tiles + MSUB - 1) / MSUB) * n_tiles;
  const int kb_total = a.kb_total;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    mbar_init(bar_tfull, 1);
    mbar_init(bar_tfull + 8, 1);
    mbar_init(bar_tempty, 128);
    mbar_init(bar_tempty + 8, 128);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<TMEM_COLS>(smem_u32(const_cast<uint32_t*>(tmem_slot)));
  if (warp >= 2)
    for (int i = threadIdx.x - 64; i < a.cout; i += 128) {
      ss[2 * i] = a.scale[i];          // interleaved (scale, shift) pairs: the epilogue reads two channels per 16-byte load
      ss[2 * i + 1] = a.shift[i];
    }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // tile t -> channel tile and MSUB consecutive pixel tiles; a pixel tile past the end lands on an image index
  // beyond the tensor (TMA zero-fills, the epilogue masks it by n < n_img)
  auto tile_coords = [&](int t, int j, int* w0, int* h0, int* n0, int* col0) {
    const int nt = t % n_tiles, mt = (t / n_tiles) * MSUB + j;
    const int tw = mt % a.tiles_w;
    const int th = (mt / a.tiles_w) % a.tiles_h;
    const int nb = mt / (a.tiles_w * a.tiles_h);
    *w0 = tw << a.lg_bw;
    *h0 = th << a.lg_bh;
    *n0 = nb << lg_bi;
    *col0 = nt * BN;
  };

  if (warp == 0) {
    if (lane == 0) {
      pdl_trigger();
      pdl_wait();
      int s = 0, ph = 0;     // ring position / phase carried across tiles
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        int w0[MSUB], h0[MSUB], n0[MSUB], col0;
#pragma unroll
        for (int j = 0; j < MSUB; ++j) tile_coords(t, j, &w0[j], &h0[j], &n0[j], &col0);
        for (int kb = 0; kb < kb_total; ++kb) {
          mbar_wait(bar_empty + 8 * s, ph ^ 1);
          const uint32_t sa = base + s * L::STAGE_BYTES;
          const uint32_t bar = bar_full + 8 * s;
          mbar_arrive_expect_tx(bar, L::STAGE_BYTES);
          const int tap = kb / a.kb_per_tap;
          const int cc = kb - tap * a.kb_per_tap;
          const int kh = tap / 3, kw = tap - kh * 3;
#pragma unroll
          for (int j = 0; j < MSUB; ++j) tma_load_4d(sa + j * 128 * 128, &tmA, bar, cc * BKE, w0[j] + kw - 1, h0[j] + kh - 1, n0[j]);
          tma_load_2d(sa + L::A_BYTES, &tmB, bar, kb * BKE, col0);
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // whole warp in uniform control flow, one elected lane around the MMAs (descriptors stay in uniform registers, no
    // ELECT loop per UTCHMMA); the next stage's barrier is polled before this stage's MMAs are issued
    {
      constexpr uint32_t idesc = make_idesc_16<F16>(128, BN);
      int s = 0, ph = 0, it = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
        const int acc = it & 1;
        if (it >= 2) mbar_wait(bar_tempty + 8 * acc, ((it >> 1) - 1) & 1);
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + acc * MSUB * BN;
        bool ready = mbar_test_wait(bar_full + 8 * s, ph);
        for (int kb = 0; kb < kb_total; ++kb) {
          if (!ready) mbar_wait(bar_full + 8 * s, ph);
          const int sn = s + 1 == STAGES ? 0 : s + 1;
          const int phn = s + 1 == STAGES ? ph ^ 1 : ph;
          ready = kb + 1 < kb_total && mbar_test_wait(bar_full + 8 * sn, phn);
          tcgen05_fence_after();
          const uint32_t sa = base + s * L::STAGE_BYTES;
          const uint64_t bdesc = make_kmajor_desc<128>(sa + L::A_BYTES);
          if (elect_one()) {
#pragma unroll
            for (int j = 0; j < MSUB; ++j) {
              const uint64_t adesc = make_kmajor_desc<128>(sa + j * 128 * 128);
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_bf16_ss(d_tmem + j * BN, adesc + 2 * k, bdesc + 2 * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
            }
            umma_commit(bar_empty + 8 * s);
          }
          __syncwarp();
          s = sn; ph = phn;
        }
        if (elect_one()) umma_commit(bar_tfull + 8 * acc);
        __syncwarp();
      }
    }
  } else {
    const int g = warp & 3;
    const int r = g * 32 + lane;
    const int BW = 1 << a.lg_bw, BH = 1 << a.lg_bh;
    const int wl = r & (BW - 1);
    const int hl = (r >> a.lg_bw) & (BH - 1);
    const int nl = r >> (a.lg_bw + a.lg_bh);
    __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(a.out);
    int it = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
      const int acc = it & 1;
      mbar_wait(bar_tfull + 8 * acc, (it >> 1) & 1);
      tcgen05_fence_after();
#pragma unroll 1
      for (int j = 0; j < MSUB; ++j) {
        int w0, h0, n0, col0;
        tile_coords(t, j, &w0, &h0, &n0, &col0);
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(g * 32) << 16) + (acc * MSUB + j) * BN;
        const int n = n0 + nl;
        const bool img_ok = n < a.n_img;
        __nv_bfloat16* orow;
        if (!POOL) orow = out + ((static_cast<size_t>(a.img_off_out + n) * a.H + (h0 + hl)) * a.W + (w0 + wl)) * a.cout + col0;
        else orow = out + ((static_cast<size_t>(a.img_off_out + n) * (a.H >> 1) + ((h0 + hl) >> 1)) * (a.W >> 1) + ((w0 + wl) >> 1)) * a.cout + col0;
        const bool writer = img_ok && (!POOL || (((wl | hl) & 1) == 0));
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
          uint32_t v[32];
          tmem_ld_32x32(taddr + c0, v);
          tmem_ld_wait();
          if (j == MSUB - 1 && c0 + 32 == BN) {   // accumulators fully read: hand the TMEM buffer back to the MMA warp
            tcgen05_fence_before();
            mbar_arrive(bar_tempty + 8 * acc);
          }
          uint32_t p[16];
#pragma unroll
          for (int c = 0; c < 32; c += 2) {
            // (scale, shift) of channels c, c+1 in ONE broadcast 16-byte load: with one 4-byte load per operand the
            // epilogue's shared-memory wavefronts were 30 % of the smem pipe next to the MMA operand reads (ncu, conv7)
            const float4 q = *reinterpret_cast<const float4*>(ss + 2 * (col0 + c0 + c));
            const float x0 = fmaf(__uint_as_float(v[c]), q.x, q.y);
            const float x1 = fmaf(__uint_as_float(v[c + 1]), q.z, q.w);
            p[c >> 1] = pack16x2_relu<F16>(x0, x1);
          }
          if (POOL) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const uint32_t mu = max16x2<F16>(p[i], __shfl_xor_sync(0xffffffffu, p[i], 1));
              p[i] = max16x2<F16>(mu, __shfl_xor_sync(0xffffffffu, mu, BW));
            }
          }
          if (writer) {
            st_global_v8(orow + c0, p);
            st_global_v8(orow + c0 + 16, p + 8);
          }
        }
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


// ---- host launchers (opt the kernel in to its dynamic shared memory on the current device, PDL attribute set)
template <int BN>
inline cudaError_t launch_tc_gemm(dim3 grid, cudaStream_t st, const CUtensorMap& a, const CUtensorMap& b, const TcArgs& args) {
  // (a deeper TMA ring - 8 stages - was measured slower: 192 KB of smem leaves one CTA per SM instead of two)
  // BN = 128: one CTA per SM anyway (4 x 32 KB) -> 3 stages of two k-blocks; BN = 64 keeps 4 x 24 KB and two CTAs per SM,
  // whose hand-offs hide behind each other's MMAs
  if (BN == 128) return ffh::launch_smem(tc_gemm_kernel<BN, 3, 2>, grid, dim3(192), TcSmem<128, BN, 6>::TOTAL, st, true, a, b, args);
  return ffh::launch_smem(tc_gemm_kernel<BN, 4, 1>, grid, dim3(192), TcSmem<128, BN, 4>::TOTAL, st, true, a, b, args);
}
template <int BN, int MSUB, bool POOL, int STAGES, bool F16 = false>
inline cudaError_t launch_ptc(int grid, cudaStream_t st, const CUtensorMap& a, const CUtensorMap& b, const TcArgs& args) {
  return ffh::launch_smem(ptc_conv_kernel<BN, MSUB, POOL, STAGES, F16>, dim3(grid), dim3(192), PtcSmem<BN, MSUB, STAGES>::TOTAL, st, true, a, b, args);
}

}  // namespace ff
