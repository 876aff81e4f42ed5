// ff_tc.cuh — the generic tensor-core GEMM kernel of the engine and the argument block the conv kernels share:
//   * tc_gemm_kernel: y = x W^T (+bias, activation, residual) for the patch embedding, the per-op encoder linears
//     (fallback when the one-launch encoder cannot be co-resident) and the MLP head (cvit.py:26-28,40-41,155,161-165):
//     one 128 x BN tile per CTA, fp32 accumulator in TMEM, 64-element (SW128) k-blocks.
//       warp 0   TMA producer (A: 2-D box of the activation rows, B: 2-D box {64, BN} of the [out][in] weights)
//       warp 1   TMEM alloc + single-thread tcgen05.mma issue (cta_group::1, kind::f16, M=128, N=BN, K=16),
//                tcgen05.commit releases smem stages / signals the epilogue
//       warps 2-5 epilogue: tcgen05.ld 32 lanes x 32 columns, bias / activation / residual, 32-byte global stores
//   * TcArgs: geometry / epilogue arguments of the persistent implicit-GEMM convolutions (ff_ptcw.cuh: layers 7-9,
//     ff_ptc2.cuh: layers 10-17, ff_ws.cuh: layers 3-6, ff_rvk.cuh).
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

// ---- host launchers (opt the kernel in to its dynamic shared memory on the current device, PDL attribute set)
template <int BN>
inline cudaError_t launch_tc_gemm(dim3 grid, cudaStream_t st, const CUtensorMap& a, const CUtensorMap& b, const TcArgs& args) {
  // (a deeper TMA ring - 8 stages - was measured slower: 192 KB of smem leaves one CTA per SM instead of two)
  // BN = 128: one CTA per SM anyway (4 x 32 KB) -> 3 stages of two k-blocks; BN = 64 keeps 4 x 24 KB and two CTAs per SM,
  // whose hand-offs hide behind each other's MMAs
  if (BN == 128) return ffh::launch_smem(tc_gemm_kernel<BN, 3, 2>, grid, dim3(192), TcSmem<128, BN, 6>::TOTAL, st, true, a, b, args);
  return ffh::launch_smem(tc_gemm_kernel<BN, 4, 1>, grid, dim3(192), TcSmem<128, BN, 4>::TOTAL, st, true, a, b, args);
}

}  // namespace ff
