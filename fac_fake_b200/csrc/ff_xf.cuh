// ff_xf.cuh — the whole ViT encoder of the CViT path as ONE kernel launch.
//
// Reference op: Transformer / Residual / PreNorm / Attention / FeedForward,
//   /root/reference/CViT-main/model/cvit.py:5-78 (called at :176):
//   6 x { x += Wo . Attn(LN(x)) + bo ;  x += W2 . GELU(W1 . LN(x) + b1) + b2 }.
//
// Why one kernel: at the benchmark batch the encoder sees only M = 2 tokens x 512 crops = 1024 rows, so each of its
// 24 GEMMs is 2-6 GFLOP — 2-4 us of tensor work that cost 15-20 us as a separate launch (prologue, cold smem ring,
// weights fetched from HBM behind a 4-deep ring, drain, launch gap; 7 launches per layer).  Every dependency of the
// encoder stays inside a 128-row token tile (LayerNorm is per row, attention mixes the two tokens of one crop), so a
// tile never has to wait for another:
//
//   * a GROUP of 16 CTAs owns one 128-row tile for all layers; CTA r of the group computes the r-th N-slice of every
//     linear (qkv 192, out 64, ff1 128, ff2 64 columns: one tcgen05 tile, M = 128, per phase);
//   * phases qkv | attention | out+residual | ff1+GELU | ff2+residual are separated by a group barrier
//     (one L2 atomic + an acquire spin per CTA) instead of kernel boundaries; activations are exchanged through L2.
//   * LayerNorm has no phase of its own (it needed two of the seven barriers per layer): it is folded into the linear
//     that follows it.  The residual epilogues (and tokens_kernel for layer 0) write xb = bf16(x - s) next to the fp32
//     stream, s = the row mean at the PREVIOUS LayerNorm point (keeps xb centred, so its bf16 rounding is relative to
//     the deviation from the mean as LN(x) rounded to bf16 would be), and per-segment (sum, centred sum of squares) of
//     the 32 columns each epilogue thread owns.  The consumer GEMM runs on xb with W' = W * gamma and its epilogue applies
//     rstd * (acc - (mean - s) * c1[n]) + c2[n]   (c1 = W' . 1, c2 = W . beta + bias; upload_folded_linear in ff_cvit.cu).
//     Two statistics buffers alternate between the two LayerNorm points of a layer, so a row's 32 segment slots are never
//     read while they are rewritten.  Plain stores, fixed summation order: deterministic.
//     The launch is cooperative, so all CTAs are co-resident and the spin cannot deadlock.  (16-CTA thread-block
//     clusters were tried first: only 7 such clusters are co-resident on a B200, the 8th tile ran as a second wave.)
//   * the 200 KB TMA ring and the TMEM allocation live for the whole kernel.  Every GEMM re-cuts the ring: 5 stages of one
//     k-block (40 KB, qkv), 3 stages of two k-blocks (64 KB, ff1), 4 stages of two k-blocks (48 KB, out, ff2).  Weights do not depend
//     on the previous phase: once the MMAs of a GEMM are complete the producer re-cuts the ring, loads the first B
//     tiles of the NEXT GEMM under the epilogue and the group barrier and L2-prefetches the rest of that GEMM's weight
//     slice — after the barrier only L2 hits are on the critical path.
//   * what paces the mainloops (clock64 traces, tools/xf_trace.py, profiles/r02_xf_trace.txt): with one k-block per
//     stage a k-block completed every ~550-600 cycles (4 MMAs need 192-384) whatever the stage size, the ring depth, the
//     number of CTAs, the k-block order across the 16 CTAs, the source of the A tile (fresh activations or static
//     memory) and the number of TMA operations per tile — a fixed cost per full/empty hand-off (barrier poll + commit in
//     the issuing warp, ~300 cycles), first hidden behind a slow MMA issue (below).  Two k-blocks per stage: ~390.
//
//   warp 0    TMA producer (one elected lane)        warp 1   tcgen05.mma issuer (one elected lane), TMEM owner
//   warps 2-9 epilogues (two warps per TMEM lane group, half of the columns each) + row statistics + 2-token attention
#pragma once
#include "ff_ptx.cuh"
#include "ff_small.cuh"
#include "ff_tc.cuh"

namespace ff {

constexpr int XF_CS = 16;        // CTAs per group = N-slices per 128-row token tile
constexpr int XF_STAGES = 8;     // mbarrier pairs; a GEMM uses 3 to 5 of them
constexpr int XF_THREADS = 320;
constexpr int XF_MAX_DEPTH = 6;
constexpr int XF_MAX_GROUPS = 64;
constexpr int XF_A_BYTES = 128 * 128;          // 128 token rows x 64 bf16
constexpr int XF_BBOX_BYTES = 64 * 128;        // one weight box: 64 output features x 64 bf16
constexpr int XF_RING_BYTES = 5 * (XF_A_BYTES + 3 * XF_BBOX_BYTES);   // 200 KB = 5 x 40 KB (qkv) >= 3 x 64 KB (ff1), 4 x 48 KB (out, ff2)
constexpr int XF_BAR_OFF = XF_RING_BYTES;                // full[S], empty[S], acc
constexpr int XF_SLOT_OFF = XF_BAR_OFF + (2 * XF_STAGES + 1) * 8;
constexpr int XF_VEC_OFF = ((XF_SLOT_OFF + 16 + 15) / 16) * 16;      // c1 | c2 slice of the current folded GEMM (2 x 192 floats)
constexpr int XF_SMEM_TOTAL = XF_VEC_OFF + 2 * 192 * 4 + 1024;
constexpr int XF_TMEM_COLS = 256;

struct XfLayerP {
  const float *c1q, *c2q, *c1f, *c2f, *b_out, *b_ff2;     // folded-LayerNorm vectors of to_qkv / net.0, biases of to_out / net.2
};
struct XfArgs {
  float* x;                  // [rows][1024] fp32 residual stream (in/out)
  __nv_bfloat16* xn;         // [rows_cap][1024] xb = bf16(x - shift): A operand of qkv / ff1
  __nv_bfloat16* qkv;        // [rows_cap][3072]
  __nv_bfloat16* att;        // [rows_cap][1024]
  __nv_bfloat16* ffh;        // [rows_cap][2048]
  // tensor maps, in the kernel parameter (constant) space like every other kernel's: [0] xn, [1] att, [2] ffh (box {64, 128});
  // [3 + 4*l + {0,1,2,3}] = to_qkv, to_out, net.0, net.2 weights of layer l (box {64, 64})
  CUtensorMap maps[3 + 4 * XF_MAX_DEPTH];
  unsigned int* sync;        // [2][XF_MAX_GROUPS] group-barrier arrival counters + exit counters; zero at rest
  float2* stats;             // [2][rows_cap][32] per-segment (sum, centred sum of squares): [0] read by qkv, [1] by ff1
  long long stats_stride;    // rows_cap * 32
  int rows, n_crops, depth;
  float eps1, eps2;
  XfLayerP L[XF_MAX_DEPTH];
};

// CTA-wide barrier that tolerates diverged warps (the elected producer / MMA lanes arrive on their own)
__device__ __forceinline__ void xf_cta_sync() { asm volatile("barrier.sync 0;" ::: "memory"); }

// Barrier of the 16 CTAs of a group.  Writers have executed fence.proxy.async; the CTA barrier orders their stores
// before thread 0's gpu-scope fence (release side: fence + relaxed arrival); the spin is a RELAXED load — an acquire
// load costs a MEMBAR.ALL.GPU per iteration — followed by one fence (acquire side), and the second CTA barrier orders
// the other CTAs' stores before every reader of this CTA.  Two gpu-scope fences per barrier, none inside the spin.
__device__ __forceinline__ void xf_group_sync(unsigned int* counter, unsigned int target) {
  xf_cta_sync();
  if (threadIdx.x == 0) {
    asm volatile("fence.acq_rel.gpu;" ::: "memory");
    asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
    unsigned int v;
    const long long t0 = clock64();
    while (true) {
      asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
      if (static_cast<int>(v - target) >= 0) break;
      if (clock64() - t0 > 4000000000LL) {   // a protocol bug becomes a launch error instead of a hung GPU
        printf("ff: encoder group barrier timeout block %d target %u seen %u\n", blockIdx.x, target, v);
        __trap();
      }
    }
    asm volatile("fence.acq_rel.gpu;" ::: "memory");
  }
  xf_cta_sync();
}
// generic-proxy global writes of this thread -> visible to later async-proxy (TMA) reads
__device__ __forceinline__ void xf_fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void xf_prefetch_l2_2d(const CUtensorMap* m, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(reinterpret_cast<uint64_t>(m)),
               "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void xf_prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

__device__ __forceinline__ float4 xf_ldcg4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }

// Row statistics of the 32 rows of a TMEM lane group, by one warp: for each row the 32 lanes read its 32 segment entries
// with ONE coalesced 256-byte request (a thread reading its own row would fetch sixteen 32-byte sectors for 16 bytes each:
// 256 KB of L2->SM traffic per CTA and phase, as much as the A operand, and the A loads queued behind it — measured),
// reduce them in butterfly order, and lane i keeps the result of row i.
//   mean, rstd: LayerNorm statistics (cvit.py:16-20), Chan's combination of the per-segment centred sums.
__device__ __forceinline__ void xf_warp_row_stats(const float2* __restrict__ st, int lane, float eps, float* mean, float* rstd) {
  float my_mean = 0.0f, my_m2 = 0.0f;
  float2 t[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) t[i] = __ldcg(st + i * 32 + lane);     // all 32 requests in flight before the first use
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    const float mu = warp_sum_bfly(t[i].x) * (1.0f / 1024.0f);
    const float d = t[i].x * (1.0f / 32.0f) - mu;
    const float m2 = warp_sum_bfly(fmaf(32.0f * d, d, t[i].y));
    if (i == lane) { my_mean = mu; my_m2 = m2; }
  }
  *mean = my_mean;
  *rstd = rsqrtf(my_m2 * (1.0f / 1024.0f) + eps);
}
// Row means only (the shift xb was / will be written with): same reduction order as above and as tokens_kernel.
__device__ __forceinline__ float xf_warp_row_mean(const float2* __restrict__ st, int lane) {
  float my_mean = 0.0f;
  float s[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) s[i] = __ldcg(st + i * 32 + lane).x;
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    const float mu = warp_sum_bfly(s[i]) * (1.0f / 1024.0f);
    if (i == lane) my_mean = mu;
  }
  return my_mean;
}

// 2-token attention of one crop, heads h0..h0+3, by one warp (cvit.py:43-60; scale = dim**-0.5 = 1/32, cvit.py:38).
__device__ __forceinline__ void xf_attention_crop(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out,
                                                  int crop, int h0, int lane) {
  auto ld4 = [](const __nv_bfloat16* p) {
    const uint2 u = __ldcg(reinterpret_cast<const uint2*>(p));
    const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
    const float2 c = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
    return make_float4(a.x, a.y, c.x, c.y);
  };
  auto dot = [](const float4& a, const float4& c) { return (a.x * c.x + a.y * c.y) + (a.z * c.z + a.w * c.w); };
  float4 q0[4], q1[4], k0[4], k1[4], v0[4], v1[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __nv_bfloat16* r0 = qkv + static_cast<size_t>(2 * crop) * 3072 + (h0 + i) * 128 + lane * 4;
    const __nv_bfloat16* r1 = r0 + 3072;
    q0[i] = ld4(r0); q1[i] = ld4(r1);
    k0[i] = ld4(r0 + 1024); k1[i] = ld4(r1 + 1024);
    v0[i] = ld4(r0 + 2048); v1[i] = ld4(r1 + 2048);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float s = 0.03125f;
    const float d00 = warp_sum(dot(q0[i], k0[i])) * s, d01 = warp_sum(dot(q0[i], k1[i])) * s;
    const float d10 = warp_sum(dot(q1[i], k0[i])) * s, d11 = warp_sum(dot(q1[i], k1[i])) * s;
    const float m0 = fmaxf(d00, d01), m1 = fmaxf(d10, d11);
    const float e00 = expf(d00 - m0), e01 = expf(d01 - m0), e10 = expf(d10 - m1), e11 = expf(d11 - m1);
    const float i0 = 1.0f / (e00 + e01), i1 = 1.0f / (e10 + e11);
    const float a00 = e00 * i0, a01 = e01 * i0, a10 = e10 * i1, a11 = e11 * i1;
    __nv_bfloat16* o0 = out + static_cast<size_t>(2 * crop) * 1024 + (h0 + i) * 128 + lane * 4;
    *reinterpret_cast<uint2*>(o0) =
        make_uint2(pack_bf16x2(a00 * v0[i].x + a01 * v1[i].x, a00 * v0[i].y + a01 * v1[i].y),
                   pack_bf16x2(a00 * v0[i].z + a01 * v1[i].z, a00 * v0[i].w + a01 * v1[i].w));
    *reinterpret_cast<uint2*>(o0 + 1024) =
        make_uint2(pack_bf16x2(a10 * v0[i].x + a11 * v1[i].x, a10 * v0[i].y + a11 * v1[i].y),
                   pack_bf16x2(a10 * v0[i].z + a11 * v1[i].z, a10 * v0[i].w + a11 * v1[i].w));
  }
}

enum { XF_G_QKV = 0, XF_G_OUT = 1, XF_G_FF1 = 2, XF_G_FF2 = 3 };

// Exact (erf) GELU of nn.GELU (cvit.py:28) with erfc from Abramowitz-Stegun 7.1.26 (|error| <= 1.5e-7, far below the bf16
// rounding of the result): 0.5 x (1 + erf(x / sqrt 2)) = x (1 - q) for x >= 0, x q for x < 0, q = 0.5 erfc(|x| / sqrt 2).
// About half the instructions of erff(); the ff1 epilogue is issue-bound on this function (64 values per thread).
__device__ __forceinline__ float xf_gelu(float x) {
  const float z = fabsf(x) * 0.70710678118654752440f;
  float t, ex;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.0f)));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex) : "f"(-1.4426950408889634f * z * z));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  const float q = 0.5f * p * t * ex;
  return x * (x >= 0.0f ? 1.0f - q : q);
}

// Developer aid, compiled only with -DFF_XF_TRACE (never in the shipped library): block 0 records clock64 after every
// group barrier (thread 0) and when every accumulator is complete / published (first epilogue thread) and prints them.
#ifdef FF_XF_TRACE
__device__ long long xf_trace_buf[9][64];
#define XF_TRACE(who, idx) do { const int i_ = (idx); if (blockIdx.x == 0 && i_ < 64) xf_trace_buf[who][i_] = clock64(); } while (0)
#else
#define XF_TRACE(who, idx) do { } while (0)
#endif

__global__ void __launch_bounds__(XF_THREADS, 1) xf_kernel(const __grid_constant__ XfArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bar_full = base + XF_BAR_OFF;
  const uint32_t bar_empty = bar_full + XF_STAGES * 8;
  const uint32_t bar_acc = bar_empty + XF_STAGES * 8;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(base_ptr + XF_SLOT_OFF);
  float* const s_vec = reinterpret_cast<float*>(base_ptr + XF_VEC_OFF);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int rank = blockIdx.x % XF_CS;
  const int group = blockIdx.x / XF_CS;
  const int ngroups = gridDim.x / XF_CS;
  const int ntiles = (a.rows + 127) >> 7;
  unsigned int* const sync_ctr = a.sync + group;
  float2* const stats_a = a.stats;                      // written by tokens_kernel / the ff2 epilogue, LayerNorm 1
  float2* const stats_b = a.stats + a.stats_stride;     // written by the to_out epilogue, LayerNorm 2
  unsigned int sync_target = 0;

  if (threadIdx.x == 0) {
    pdl_trigger();     // the next kernel in the stream (launched with PDL) may be scheduled as SMs free up
    for (int s = 0; s < XF_STAGES; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    mbar_init(bar_acc, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<XF_TMEM_COLS>(smem_u32(const_cast<uint32_t*>(tmem_slot)));
  // folded-LayerNorm vectors and biases of every layer -> L2 (cold after the conv stack swept the cache); 128-byte lines
  for (int i = threadIdx.x; i < a.depth * 384; i += XF_THREADS) {
    const XfLayerP& P = a.L[i / 384];
    const int j = i % 384;       // 96 + 96 lines of c1q / c2q, 64 + 64 of c1f / c2f, 32 + 32 of the biases
    const float* p = j < 96 ? P.c1q + j * 32 : j < 192 ? P.c2q + (j - 96) * 32 : j < 256 ? P.c1f + (j - 192) * 32
                   : j < 320 ? P.c2f + (j - 256) * 32 : j < 352 ? P.b_out + (j - 320) * 32 : P.b_ff2 + (j - 352) * 32;
    xf_prefetch_l2(p);
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // static description of GEMM g of a layer for this CTA
  auto gemm_kb = [](int g) { return g == XF_G_FF2 ? 32 : 16; };
  auto gemm_nb = [](int g) { return g == XF_G_QKV ? 3 : (g == XF_G_FF1 ? 2 : 1); };   // 64-row weight boxes
  auto gemm_amap = [](int g) { return g == XF_G_OUT ? 1 : (g == XF_G_FF2 ? 2 : 0); };

  // Both single-thread roles run as WHOLE warps in uniform control flow and elect one lane only around the TMA / MMA
  // instructions themselves: inside `if (lane == 0)` ptxas keeps the descriptors in vector registers and wraps every
  // UTCHMMA / UTMALDG in an ELECT + 4x R2UR.BROADCAST loop — measured ~150 cycles per tcgen05.mma whatever its N, i.e.
  // the issue rate, not the operand fill, set the pace of these mainloops (ff_xf history).
  const bool is_producer = (warp == 0);
  const bool is_mma = (warp == 1);
  const bool is_epi = warp >= 2;
  const int ew = warp - 2;                 // epilogue warp 0..7
  const int hcol = ew >> 2;                // which half of the accumulator columns
  const int g4 = warp & 3;                 // TMEM lane group this warp may access
  const int r = g4 * 32 + lane;            // accumulator row of an epilogue thread
  const uint32_t taddr = tmem_base + (static_cast<uint32_t>(g4 * 32) << 16);

  int trace_mma = 0, trace_mma2 = 0, trace_prod = 0;
  (void)trace_mma; (void)trace_mma2; (void)trace_prod;
  // A stage holds gemm_kps(g) k-blocks (each = A tile + its weight boxes, back to back) behind ONE full/empty barrier pair:
  // the N = 64 / 128 GEMMs take two k-blocks per stage, i.e. eight MMAs per barrier wait and per tcgen05.commit
  // (one k-block per stage left the tensor pipe dry between k-blocks: wait + commit cost as much as four N = 64 MMAs).
  auto gemm_kps = [](int g) { return g == XF_G_QKV ? 1 : 2; };
  auto gemm_units = [&](int g) { return gemm_kb(g) / gemm_kps(g); };
  auto gemm_kblock_bytes = [&](int g) { return XF_A_BYTES + gemm_nb(g) * XF_BBOX_BYTES; };
  auto gemm_stage_bytes = [&](int g) { return gemm_kps(g) * gemm_kblock_bytes(g); };
  auto gemm_stages = [](int g) { return g == XF_G_QKV ? 5 : (g == XF_G_FF1 ? 3 : 4); };   // 5 x 40, 3 x 64, 4 x 48 KB
  // ---- producer state: B pointer runs ahead (acquires stages), A pointer follows.  Stage geometry belongs to the GEMM
  //      being loaded; bit s of pe = parity of the next acquisition of stage s (stages are used unevenly)
  int sB = 0, sA = 0, npre = 0, acc_phase_p = 0;
  uint32_t pe = 0;
  auto issue_b = [&](int layer, int g, int u) {     // acquire the next stage and load the weight boxes of stage unit u
    mbar_wait(bar_empty + 8 * sB, ((pe >> sB) & 1u) ^ 1u);
    pe ^= 1u << sB;
    if (lane == 0 && trace_prod == 5 && g == XF_G_OUT) XF_TRACE(7, u);
    const int nb = gemm_nb(g), kps = gemm_kps(g);
    const uint32_t bar = bar_full + 8 * sB;
    if (elect_one()) {
      mbar_arrive_expect_tx(bar, gemm_stage_bytes(g));
      const CUtensorMap* tmB = a.maps + 3 + 4 * layer + g;
      for (int sub = 0; sub < kps; ++sub) {
        const uint32_t sb = base + sB * gemm_stage_bytes(g) + sub * gemm_kblock_bytes(g) + XF_A_BYTES;
        for (int j = 0; j < nb; ++j) tma_load_2d(sb + j * XF_BBOX_BYTES, tmB, bar, (u * kps + sub) * 64, (rank * nb + j) * 64);
      }
    }
    __syncwarp();
    if (++sB == gemm_stages(g)) sB = 0;
  };
  auto issue_a = [&](int g, int u, int m0) {
    if (elect_one()) {
      const int kps = gemm_kps(g);
      for (int sub = 0; sub < kps; ++sub)
        tma_load_2d(base + sA * gemm_stage_bytes(g) + sub * gemm_kblock_bytes(g), a.maps + gemm_amap(g), bar_full + 8 * sA,
                    (u * kps + sub) * 64, m0);
    }
    __syncwarp();
    if (++sA == gemm_stages(g)) sA = 0;
  };
  // weights of GEMM (layer, g) — the ring is empty: re-cut it, first stage units into the ring, the rest of this CTA's slice into L2
  auto preissue = [&](int layer, int g) {
    const int ut = gemm_units(g), nb = gemm_nb(g), kps = gemm_kps(g);
    const CUtensorMap* tmB = a.maps + 3 + 4 * layer + g;
    sB = 0; sA = 0;
    npre = min(gemm_stages(g), ut);
    for (int u = 0; u < npre; ++u) issue_b(layer, g, u);
    if (elect_one())
      for (int kb = npre * kps; kb < ut * kps; ++kb)
        for (int j = 0; j < nb; ++j) xf_prefetch_l2_2d(tmB, kb * 64, (rank * nb + j) * 64);
    __syncwarp();
  };
  // after the barrier that publishes the A operand: finish the loads of GEMM (layer, g); when its MMAs are complete
  // (the ring is empty again) run ahead into the next GEMM
  auto produce = [&](int layer, int g, int m0, bool more_tiles) {
    xf_fence_proxy_async();
    const int ut = gemm_units(g);
    for (int u = 0; u < ut; ++u) {
      if (u >= npre) issue_b(layer, g, u);
      issue_a(g, u, m0);
    }
    if (lane == 0) XF_TRACE(5, trace_prod++);
    mbar_wait(bar_acc, acc_phase_p);
    acc_phase_p ^= 1;
    if (g < 3) preissue(layer, g + 1);
    else if (layer + 1 < a.depth) preissue(layer + 1, 0);
    else if (more_tiles) preissue(0, 0);
  };

  // ---- MMA state: bit s of pf = parity of the next completion of full[s]
  uint32_t pf = 0;
  auto mma = [&](int g) {
    const int ut = gemm_units(g), ns = gemm_stages(g), sbytes = gemm_stage_bytes(g), kbytes = gemm_kblock_bytes(g), kps = gemm_kps(g);
    const uint32_t idesc = g == XF_G_QKV ? make_idesc_bf16(128, 192) : (g == XF_G_FF1 ? make_idesc_bf16(128, 128) : make_idesc_bf16(128, 64));
    tcgen05_fence_after();
    int sM = 0;
    // The state of the NEXT stage's barrier is polled (non-blocking test_wait) before this stage's MMAs are issued, so the
    // latency of the poll overlaps the issue.
    bool ready = mbar_test_wait(bar_full, pf & 1u);
    for (int u = 0; u < ut; ++u) {
      if (!ready) mbar_wait(bar_full + 8 * sM, (pf >> sM) & 1u);
      pf ^= 1u << sM;
      const int sN = sM + 1 == ns ? 0 : sM + 1;
      ready = u + 1 < ut && mbar_test_wait(bar_full + 8 * sN, (pf >> sN) & 1u);
      tcgen05_fence_after();
      if (lane == 0 && u == 0) XF_TRACE(3, trace_mma++);
      if (lane == 0 && trace_mma == 6) XF_TRACE(6, u);
      if (lane == 0 && u == ut - 1) XF_TRACE(4, trace_mma2++);
      const uint32_t sa = base + sM * sbytes;
      if (elect_one()) {
        for (int sub = 0; sub < kps; ++sub) {
          const uint64_t adesc = make_kmajor_desc<128>(sa + sub * kbytes);
          const uint64_t bdesc = make_kmajor_desc<128>(sa + sub * kbytes + XF_A_BYTES);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16_ss(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, (u > 0 || sub > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(bar_empty + 8 * sM);
      }
      __syncwarp();
      sM = sN;
    }
    if (elect_one()) umma_commit(bar_acc);
    __syncwarp();
  };

  int trace_bar = 0, trace_acc = 0, trace_pub = 0;
  (void)trace_bar; (void)trace_acc; (void)trace_pub;
  int acc_phase = 0;
  auto acc_wait = [&]() {
    mbar_wait(bar_acc, acc_phase);
    acc_phase ^= 1;
    tcgen05_fence_after();
    if (threadIdx.x == 64) XF_TRACE(1, trace_acc++);
  };
  // every global store of a phase is followed by this before the group barrier
  auto publish = [&]() {
    xf_fence_proxy_async();
    tcgen05_fence_before();
    if (threadIdx.x == 64) XF_TRACE(2, trace_pub++);
  };
  auto group_sync = [&]() {
    sync_target += XF_CS;
    xf_group_sync(sync_ctr, sync_target);
    if (threadIdx.x == 0) XF_TRACE(0, trace_bar++);
  };

  if (is_producer && group < ntiles) preissue(0, XF_G_QKV);

#pragma unroll 1
  for (int tile = group; tile < ntiles; tile += ngroups) {
    const int m0 = tile * 128;
    const int m = m0 + r;
    const bool row_ok = m < a.rows;
    const bool more_tiles = tile + ngroups < ntiles;
#pragma unroll 1
    for (int l = 0; l < a.depth; ++l) {
      const XfLayerP& P = a.L[l];
      // ---------------- qkv = LN1(x) . Wqkv^T   (no bias, cvit.py:40), LayerNorm folded: A = xb, statistics in stats[0]
      if (is_producer) produce(l, XF_G_QKV, m0, more_tiles);
      else if (is_mma) mma(XF_G_QKV);
      else if (is_epi) {
        // this CTA's slices of c1 / c2 -> shared memory (read back as broadcasts in the epilogue)
        for (int i = threadIdx.x - 64; i < 384; i += 256) s_vec[i] = __ldg((i < 192 ? P.c1q : P.c2q - 192) + rank * 192 + i);
        float mean, rstd;
        xf_warp_row_stats(stats_a + static_cast<size_t>(m0 + g4 * 32) * 32, lane, a.eps1, &mean, &rstd);
        const float dm = mean - xf_warp_row_mean(stats_b + static_cast<size_t>(m0 + g4 * 32) * 32, lane);
        asm volatile("bar.sync 1, 256;" ::: "memory");      // the 8 epilogue warps: s_vec is complete
        acc_wait();
        const int n0 = rank * 192 + hcol * 96;
        __nv_bfloat16* o = a.qkv + static_cast<size_t>(m) * 3072 + n0;
#pragma unroll 1
        for (int c0 = 0; c0 < 96; c0 += 32) {
          uint32_t v[32];
          tmem_ld_32x32(taddr + hcol * 96 + c0, v);
          tmem_ld_wait();
          if (row_ok) {
            uint32_t pk[16];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const float4 k1 = reinterpret_cast<const float4*>(s_vec + hcol * 96 + c0)[e];
              const float4 k2 = reinterpret_cast<const float4*>(s_vec + 192 + hcol * 96 + c0)[e];
              pk[2 * e] = pack_bf16x2(fmaf(rstd, fmaf(-dm, k1.x, __uint_as_float(v[4 * e])), k2.x),
                                      fmaf(rstd, fmaf(-dm, k1.y, __uint_as_float(v[4 * e + 1])), k2.y));
              pk[2 * e + 1] = pack_bf16x2(fmaf(rstd, fmaf(-dm, k1.z, __uint_as_float(v[4 * e + 2])), k2.z),
                                          fmaf(rstd, fmaf(-dm, k1.w, __uint_as_float(v[4 * e + 3])), k2.w));
            }
            st_global_v8(o + c0, pk);
            st_global_v8(o + c0 + 16, pk + 8);
          }
        }
        publish();
      }
      group_sync();
      // ---------------- attention -> att
      if (is_epi) {
        const int crop = tile * 64 + rank * 4 + (ew & 3);
        if (crop < a.n_crops) xf_attention_crop(a.qkv, a.att, crop, hcol * 4, lane);
        publish();
      }
      group_sync();
      // ---------------- pass 0: x += att . Wo^T + bo
      // ---------------- pass 1: ffh = GELU(LN2(x) . W1^T + b1) (folded like qkv, statistics in stats[1]),
      //                          x += ffh . W2^T + b2  (same residual epilogue)
#pragma unroll 1
      for (int pass = 0; pass < 2; ++pass) {
        const int g = pass == 0 ? XF_G_OUT : XF_G_FF2;
        if (pass == 1) {
          if (is_producer) produce(l, XF_G_FF1, m0, more_tiles);
          else if (is_mma) mma(XF_G_FF1);
          else if (is_epi) {
            for (int i = threadIdx.x - 64; i < 256; i += 256) s_vec[i] = __ldg((i < 128 ? P.c1f : P.c2f - 128) + rank * 128 + i);
            float mean, rstd;
            xf_warp_row_stats(stats_b + static_cast<size_t>(m0 + g4 * 32) * 32, lane, a.eps2, &mean, &rstd);
            const float dm = mean - xf_warp_row_mean(stats_a + static_cast<size_t>(m0 + g4 * 32) * 32, lane);
            asm volatile("bar.sync 1, 256;" ::: "memory");
            acc_wait();
            const int n0 = rank * 128 + hcol * 64;
            __nv_bfloat16* o = a.ffh + static_cast<size_t>(m) * 2048 + n0;
#pragma unroll 1
            for (int c0 = 0; c0 < 64; c0 += 32) {
              uint32_t v[32];
              tmem_ld_32x32(taddr + hcol * 64 + c0, v);
              tmem_ld_wait();
              if (row_ok) {
                uint32_t pk[16];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                  const float4 k1 = reinterpret_cast<const float4*>(s_vec + hcol * 64 + c0)[e];
                  const float4 k2 = reinterpret_cast<const float4*>(s_vec + 128 + hcol * 64 + c0)[e];
                  pk[2 * e] = pack_bf16x2(xf_gelu(fmaf(rstd, fmaf(-dm, k1.x, __uint_as_float(v[4 * e])), k2.x)),
                                          xf_gelu(fmaf(rstd, fmaf(-dm, k1.y, __uint_as_float(v[4 * e + 1])), k2.y)));
                  pk[2 * e + 1] = pack_bf16x2(xf_gelu(fmaf(rstd, fmaf(-dm, k1.z, __uint_as_float(v[4 * e + 2])), k2.z)),
                                              xf_gelu(fmaf(rstd, fmaf(-dm, k1.w, __uint_as_float(v[4 * e + 3])), k2.w)));
                }
                st_global_v8(o + c0, pk);
                st_global_v8(o + c0 + 16, pk + 8);
              }
            }
            publish();
          }
          group_sync();
        }
        if (is_producer) produce(l, g, m0, more_tiles);
        else if (is_mma) mma(g);
        else if (is_epi) {
          const float* bias = (pass == 0 ? P.b_out : P.b_ff2) + rank * 64 + hcol * 32;
          float* xr = a.x + static_cast<size_t>(m) * 1024 + rank * 64 + hcol * 32;
          // residual + bias prefetched under the MMAs (this thread is the only writer of these 32 floats)
          float4 res[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            res[i] = row_ok ? xf_ldcg4(xr + 4 * i) : make_float4(0.f, 0.f, 0.f, 0.f);
            const float4 t = __ldg(reinterpret_cast<const float4*>(bias) + i);
            res[i].x += t.x; res[i].y += t.y; res[i].z += t.z; res[i].w += t.w;
          }
          // xb of the new x is centred on the row mean at the previous LayerNorm point (pass 0: stats[0], pass 1: stats[1]);
          // the new statistics go to the other buffer
          const float2* st_src = pass == 0 ? stats_a : stats_b;
          float2* st_dst = pass == 0 ? stats_b : stats_a;
          const float shift = xf_warp_row_mean(st_src + static_cast<size_t>(m0 + g4 * 32) * 32, lane);
          acc_wait();
          uint32_t v[32];
          tmem_ld_32x32(taddr + hcol * 32, v);
          tmem_ld_wait();
          if (row_ok) {
            float t[32];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              t[4 * i] = res[i].x + __uint_as_float(v[4 * i]);
              t[4 * i + 1] = res[i].y + __uint_as_float(v[4 * i + 1]);
              t[4 * i + 2] = res[i].z + __uint_as_float(v[4 * i + 2]);
              t[4 * i + 3] = res[i].w + __uint_as_float(v[4 * i + 3]);
            }
            float sum = 0.0f;
#pragma unroll
            for (int i = 0; i < 32; ++i) sum += t[i];
            const float mu = sum * (1.0f / 32.0f);
            float m2 = 0.0f;
#pragma unroll
            for (int i = 0; i < 32; ++i) { const float d = t[i] - mu; m2 = fmaf(d, d, m2); }
            st_dst[static_cast<size_t>(m) * 32 + rank * 2 + hcol] = make_float2(sum, m2);
            uint32_t pk[16];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              uint32_t u[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) u[e] = __float_as_uint(t[8 * i + e]);
              st_global_v8(xr + 8 * i, u);
            }
#pragma unroll
            for (int i = 0; i < 16; ++i) pk[i] = pack_bf16x2(t[2 * i] - shift, t[2 * i + 1] - shift);
            __nv_bfloat16* xbr = a.xn + static_cast<size_t>(m) * 1024 + rank * 64 + hcol * 32;
            st_global_v8(xbr, pk);
            st_global_v8(xbr + 16, pk + 8);
          }
          publish();
        }
        group_sync();
      }
    }
  }
  // Every CTA of the group has left the last barrier once it arrives here; the 16th arrival re-arms both counters for
  // the next launch (nobody reads them any more), so no memset is needed between launches.
  if (threadIdx.x == 0) {
    unsigned int* done = a.sync + XF_MAX_GROUPS + group;
    if (atomicAdd(done, 1u) == XF_CS - 1) {
      *sync_ctr = 0u;
      *done = 0u;
      __threadfence();
    }
  }

  tcgen05_fence_before();
  __syncthreads();
#ifdef FF_XF_TRACE
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    // layer 1 (second layer, steady state): barriers 5..9, accumulators 4..7, publishes 5..9
    const long long t0 = xf_trace_buf[0][4];
    printf("xf trace (cycles since the barrier that ends layer 0; tile 0 group 0):\n");
    for (int i = 5; i < 10; ++i) printf("  barrier %d passed at %lld\n", i - 5, xf_trace_buf[0][i] - t0);
    for (int i = 4; i < 8; ++i) printf("  accumulator %d complete at %lld\n", i - 4, xf_trace_buf[1][i] - t0);
    for (int i = 5; i < 10; ++i) printf("  publish %d done at %lld\n", i - 5, xf_trace_buf[2][i] - t0);
    for (int i = 0; i < 16; ++i) printf("  out-proj k-block %2d: producer acquired the stage at %lld, MMA saw it full at %lld\n", i,
                                        xf_trace_buf[7][i] - t0, xf_trace_buf[6][i] - t0);
    for (int i = 4; i < 8; ++i) printf("  gemm %d: first stage full at %lld, last stage full at %lld, producer done at %lld\n", i - 4,
                                       xf_trace_buf[3][i] - t0, xf_trace_buf[4][i] - t0, xf_trace_buf[5][i] - t0);
    printf("  whole kernel: %lld cycles from first to last barrier (%d barriers)\n", xf_trace_buf[0][trace_bar - 1] - xf_trace_buf[0][0], trace_bar);
  }
#endif
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc<XF_TMEM_COLS>(tmem_base);
  }
}

}  // namespace ff
