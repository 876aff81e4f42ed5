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
//   * phases LN1 | qkv | attention | out+residual | LN2 | ff1+GELU | ff2+residual are separated by a group barrier
//     (one L2 atomic + an acquire spin per CTA) instead of kernel boundaries; activations are exchanged through L2.
//     The launch is cooperative, so all CTAs are co-resident and the spin cannot deadlock.  (16-CTA thread-block
//     clusters were tried first: only 7 such clusters are co-resident on a B200, the 8th tile ran as a second wave.)
//   * the TMA ring (5 x 40 KB) and the TMEM allocation live for the whole kernel; weights do not depend on the
//     previous phase, so the producer pre-issues the B tiles of the NEXT GEMM while the current phase drains and
//     L2-prefetches the rest of that GEMM's weight slice — after the barrier only L2 hits are on the critical path.
//
//   warp 0    TMA producer (one elected lane)        warp 1   tcgen05.mma issuer (one elected lane), TMEM owner
//   warps 2-9 epilogues (two warps per TMEM lane group, half of the columns each) + LayerNorm + 2-token attention
#pragma once
#include "ff_ptx.cuh"
#include "ff_small.cuh"
#include "ff_tc.cuh"

namespace ff {

constexpr int XF_CS = 16;        // CTAs per group = N-slices per 128-row token tile
constexpr int XF_STAGES = 5;
constexpr int XF_THREADS = 320;
constexpr int XF_MAX_DEPTH = 6;
constexpr int XF_MAX_GROUPS = 64;
constexpr int XF_A_BYTES = 128 * 128;          // 128 token rows x 64 bf16
constexpr int XF_BBOX_BYTES = 64 * 128;        // one weight box: 64 output features x 64 bf16
constexpr int XF_STAGE_BYTES = XF_A_BYTES + 3 * XF_BBOX_BYTES;
constexpr int XF_BAR_OFF = XF_STAGES * XF_STAGE_BYTES;   // full[S], empty[S], acc
constexpr int XF_SLOT_OFF = XF_BAR_OFF + (2 * XF_STAGES + 1) * 8;
constexpr int XF_SMEM_TOTAL = XF_SLOT_OFF + 16 + 1024;
constexpr int XF_TMEM_COLS = 256;

struct XfLayerP {
  const float *ln1_g, *ln1_b, *ln2_g, *ln2_b, *b_out, *b_ff1, *b_ff2;
};
struct XfArgs {
  float* x;                  // [rows][1024] fp32 residual stream (in/out)
  __nv_bfloat16* xn;         // [rows_cap][1024] LayerNorm output (A operand of qkv / ff1)
  __nv_bfloat16* qkv;        // [rows_cap][3072]
  __nv_bfloat16* att;        // [rows_cap][1024]
  __nv_bfloat16* ffh;        // [rows_cap][2048]
  // device array of tensor maps: [0] xn, [1] att, [2] ffh (box {64, 128});
  // [3 + 4*l + {0,1,2,3}] = to_qkv, to_out, net.0, net.2 weights of layer l (box {64, 64})
  const CUtensorMap* maps;
  unsigned int* sync;        // [2][XF_MAX_GROUPS] group-barrier arrival counters + exit counters; zero at rest
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

// LayerNorm of one row by one warp (cvit.py:16-20); x is read through L2 (written by other CTAs of the group).
__device__ __forceinline__ void xf_layernorm_row(const float* __restrict__ x, const float* __restrict__ gamma,
                                                 const float* __restrict__ beta, __nv_bfloat16* __restrict__ y, int row,
                                                 float eps, int lane) {
  float4 v[8], g[8], b[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = xf_ldcg4(x + static_cast<size_t>(row) * 1024 + (i * 32 + lane) * 4);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    g[i] = __ldg(reinterpret_cast<const float4*>(gamma) + i * 32 + lane);
    b[i] = __ldg(reinterpret_cast<const float4*>(beta) + i * 32 + lane);
  }
  float s = 0.0f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  const float mean = warp_sum(s) * (1.0f / 1024.0f);
  float q = 0.0f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float a = v[i].x - mean, bb = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
    q += (a * a + bb * bb) + (c * c + d * d);
  }
  const float rstd = rsqrtf(warp_sum(q) * (1.0f / 1024.0f) + eps);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float o0 = (v[i].x - mean) * rstd * g[i].x + b[i].x;
    const float o1 = (v[i].y - mean) * rstd * g[i].y + b[i].y;
    const float o2 = (v[i].z - mean) * rstd * g[i].z + b[i].z;
    const float o3 = (v[i].w - mean) * rstd * g[i].w + b[i].w;
    reinterpret_cast<uint2*>(y + static_cast<size_t>(row) * 1024)[i * 32 + lane] =
        make_uint2(pack_bf16x2(o0, o1), pack_bf16x2(o2, o3));
  }
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

__global__ void __launch_bounds__(XF_THREADS, 1) xf_kernel(const __grid_constant__ XfArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bar_full = base + XF_BAR_OFF;
  const uint32_t bar_empty = bar_full + XF_STAGES * 8;
  const uint32_t bar_acc = bar_empty + XF_STAGES * 8;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(base_ptr + XF_SLOT_OFF);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int rank = blockIdx.x % XF_CS;
  const int group = blockIdx.x / XF_CS;
  const int ngroups = gridDim.x / XF_CS;
  const int ntiles = (a.rows + 127) >> 7;
  unsigned int* const sync_ctr = a.sync + group;
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
  // LayerNorm affine vectors and biases of every layer -> L2 (cold after the conv stack swept the cache)
  for (int i = threadIdx.x; i < a.depth * 7 * 32; i += XF_THREADS) {
    const int l = i / (7 * 32), w = (i / 32) % 7, line = i % 32;
    const float* const* arr = reinterpret_cast<const float* const*>(&a.L[l]);
    xf_prefetch_l2(arr[w] + line * 32);
    if (w == 5) xf_prefetch_l2(arr[w] + (32 + line) * 32);   // b_ff1 has 2048 floats, the rest 1024
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // static description of GEMM g of a layer for this CTA
  auto gemm_kb = [](int g) { return g == XF_G_FF2 ? 32 : 16; };
  auto gemm_nb = [](int g) { return g == XF_G_QKV ? 3 : (g == XF_G_FF1 ? 2 : 1); };   // 64-row weight boxes
  auto gemm_amap = [](int g) { return g == XF_G_OUT ? 1 : (g == XF_G_FF2 ? 2 : 0); };

  const bool is_producer = (warp == 0 && lane == 0);
  const bool is_mma = (warp == 1 && lane == 0);
  const bool is_epi = warp >= 2;
  const int ew = warp - 2;                 // epilogue warp 0..7
  const int hcol = ew >> 2;                // which half of the accumulator columns
  const int g4 = warp & 3;                 // TMEM lane group this warp may access
  const int r = g4 * 32 + lane;            // accumulator row of an epilogue thread
  const uint32_t taddr = tmem_base + (static_cast<uint32_t>(g4 * 32) << 16);

  // ---- producer state: B pointer runs ahead (acquires stages), A pointer follows
  int sB = 0, phB = 0, sA = 0, npre = 0;
  auto issue_b = [&](int layer, int g, int kb) {     // acquire the next stage and load the weight boxes of k-block kb
    mbar_wait(bar_empty + 8 * sB, phB ^ 1);
    const int nb = gemm_nb(g);
    const uint32_t bar = bar_full + 8 * sB;
    mbar_arrive_expect_tx(bar, XF_A_BYTES + nb * XF_BBOX_BYTES);
    const CUtensorMap* tmB = a.maps + 3 + 4 * layer + g;
    const uint32_t sb = base + sB * XF_STAGE_BYTES + XF_A_BYTES;
    for (int j = 0; j < nb; ++j) tma_load_2d(sb + j * XF_BBOX_BYTES, tmB, bar, kb * 64, (rank * nb + j) * 64);
    if (++sB == XF_STAGES) { sB = 0; phB ^= 1; }
  };
  auto issue_a = [&](int g, int kb, int m0) {
    tma_load_2d(base + sA * XF_STAGE_BYTES, a.maps + gemm_amap(g), bar_full + 8 * sA, kb * 64, m0);
    if (++sA == XF_STAGES) sA = 0;
  };
  // weights of GEMM (layer, g): first k-blocks into the ring, the rest of this CTA's slice into L2
  auto preissue = [&](int layer, int g) {
    const int kbt = gemm_kb(g), nb = gemm_nb(g);
    const CUtensorMap* tmB = a.maps + 3 + 4 * layer + g;
    for (int kb = XF_STAGES; kb < kbt; ++kb)
      for (int j = 0; j < nb; ++j) xf_prefetch_l2_2d(tmB, kb * 64, (rank * nb + j) * 64);
    npre = min(XF_STAGES, kbt);
    for (int kb = 0; kb < npre; ++kb) issue_b(layer, g, kb);
  };
  // after the barrier that publishes the A operand: finish the loads of GEMM (layer, g), then run ahead into the next
  auto produce = [&](int layer, int g, int m0, bool more_tiles) {
    xf_fence_proxy_async();
    const int kbt = gemm_kb(g);
    for (int kb = 0; kb < kbt; ++kb) {
      if (kb >= npre) issue_b(layer, g, kb);
      issue_a(g, kb, m0);
    }
    if (g < 3) preissue(layer, g + 1);
    else if (layer + 1 < a.depth) preissue(layer + 1, 0);
    else if (more_tiles) preissue(0, 0);
  };

  // ---- MMA state
  int sM = 0, phM = 0;
  auto mma = [&](int g) {
    const int kbt = gemm_kb(g);
    const uint32_t idesc = g == XF_G_QKV ? make_idesc_bf16(128, 192) : (g == XF_G_FF1 ? make_idesc_bf16(128, 128) : make_idesc_bf16(128, 64));
    tcgen05_fence_after();
    for (int kb = 0; kb < kbt; ++kb) {
      mbar_wait(bar_full + 8 * sM, phM);
      tcgen05_fence_after();
      const uint32_t sa = base + sM * XF_STAGE_BYTES;
      const uint64_t adesc = make_kmajor_desc<128>(sa);
      const uint64_t bdesc = make_kmajor_desc<128>(sa + XF_A_BYTES);
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_bf16_ss(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
      umma_commit(bar_empty + 8 * sM);
      if (++sM == XF_STAGES) { sM = 0; phM ^= 1; }
    }
    umma_commit(bar_acc);
  };

  int acc_phase = 0;
  auto acc_wait = [&]() {
    mbar_wait(bar_acc, acc_phase);
    acc_phase ^= 1;
    tcgen05_fence_after();
  };
  // every global store of a phase is followed by this before the group barrier
  auto publish = [&]() {
    xf_fence_proxy_async();
    tcgen05_fence_before();
  };
  auto group_sync = [&]() {
    sync_target += XF_CS;
    xf_group_sync(sync_ctr, sync_target);
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
      // ---------------- LN1 -> xn
      if (is_epi) {
        const int row = m0 + rank * 8 + ew;
        if (row < a.rows) xf_layernorm_row(a.x, P.ln1_g, P.ln1_b, a.xn, row, a.eps1, lane);
        publish();
      }
      group_sync();
      // ---------------- qkv = xn . Wqkv^T   (no bias, cvit.py:40)
      if (is_producer) produce(l, XF_G_QKV, m0, more_tiles);
      else if (is_mma) mma(XF_G_QKV);
      else if (is_epi) {
        acc_wait();
        __nv_bfloat16* o = a.qkv + static_cast<size_t>(m) * 3072 + rank * 192 + hcol * 96;
#pragma unroll 1
        for (int c0 = 0; c0 < 96; c0 += 32) {
          uint32_t v[32];
          tmem_ld_32x32(taddr + hcol * 96 + c0, v);
          tmem_ld_wait();
          if (row_ok) {
            uint32_t pk[16];
#pragma unroll
            for (int e = 0; e < 16; ++e) pk[e] = pack_bf16x2(__uint_as_float(v[2 * e]), __uint_as_float(v[2 * e + 1]));
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
      // ---------------- pass 1: LN2, ffh = GELU(xn . W1^T + b1), x += ffh . W2^T + b2  (same residual epilogue)
#pragma unroll 1
      for (int pass = 0; pass < 2; ++pass) {
        const int g = pass == 0 ? XF_G_OUT : XF_G_FF2;
        if (pass == 1) {
          if (is_epi) {
            const int row = m0 + rank * 8 + ew;
            if (row < a.rows) xf_layernorm_row(a.x, P.ln2_g, P.ln2_b, a.xn, row, a.eps2, lane);
            publish();
          }
          group_sync();
          if (is_producer) produce(l, XF_G_FF1, m0, more_tiles);
          else if (is_mma) mma(XF_G_FF1);
          else if (is_epi) {
            const float* bias = P.b_ff1 + rank * 128 + hcol * 64;
            float bv[64];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const float4 t = __ldg(reinterpret_cast<const float4*>(bias) + i);
              bv[4 * i] = t.x; bv[4 * i + 1] = t.y; bv[4 * i + 2] = t.z; bv[4 * i + 3] = t.w;
            }
            acc_wait();
            __nv_bfloat16* o = a.ffh + static_cast<size_t>(m) * 2048 + rank * 128 + hcol * 64;
#pragma unroll
            for (int c0 = 0; c0 < 64; c0 += 32) {
              uint32_t v[32];
              tmem_ld_32x32(taddr + hcol * 64 + c0, v);
              tmem_ld_wait();
              if (row_ok) {
                uint32_t pk[16];
#pragma unroll
                for (int e = 0; e < 16; ++e)
                  pk[e] = pack_bf16x2(gelu_erf(__uint_as_float(v[2 * e]) + bv[c0 + 2 * e]),
                                      gelu_erf(__uint_as_float(v[2 * e + 1]) + bv[c0 + 2 * e + 1]));
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
          acc_wait();
          uint32_t v[32];
          tmem_ld_32x32(taddr + hcol * 32, v);
          tmem_ld_wait();
          if (row_ok) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float4 r0 = res[2 * i], r1 = res[2 * i + 1];
              const float rr[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
              uint32_t t[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) t[e] = __float_as_uint(rr[e] + __uint_as_float(v[8 * i + e]));
              st_global_v8(xr + 8 * i, t);
            }
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
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc<XF_TMEM_COLS>(tmem_base);
  }
}

}  // namespace ff
