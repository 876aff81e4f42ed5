// ff_ptc2.cuh — the persistent implicit-GEMM conv of
// feature layers 10..17 (Cout >= 256; reference op: nn.Conv2d + BatchNorm2d(eval) + ReLU [+ MaxPool2d(2)],
// /root/reference/CViT-main/model/cvit.py:117-147) on a CTA PAIR, tcgen05.mma.cta_group::2, tile 256 pixels x 256
// channels.
//
// Why: the single-CTA version of this kernel (128 pixels x 256 channels per CTA) filled 48 KB of shared memory per k-block per SM (16 KB of pixels + the whole 32 KB
// filter tile); over layers 7..17 that is 51 GB of L2->SM traffic per 512-crop step on a power-capped
// part.  With cta_group::2 each SM of the pair loads its own 128 pixel rows and only HALF of the filter tile
// (128 of the 256 output channels; the MMA reads the peer's half through the pair datapath): 32 KB per k-block per
// SM (-33 %), and the ring gets 6 stages instead of 4.
//
// Protocol (same as ws2x_conv_kernel, ff_ws.cuh): both CTAs issue their own TMA loads whose bytes are credited to the
// leader's full barrier; only the leader's elected thread issues the MMAs (M = 256, N = 256, K = 16); tcgen05.commit
// multicasts "stage free" and "accumulator ready" to both CTAs; both epilogues report "TMEM drained" to the
// leader's barrier (256 arrivals).  Accumulators are double-buffered in all 512 TMEM columns.
#pragma once
#include "ff_ptx.cuh"
#include "ff_tc.cuh"

namespace ff {

__device__ __forceinline__ void tma_load_2d_2cta(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          dst),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar & 0xFEFFFFFFu), "r"(c0), "r"(c1)
      : "memory");
}

struct Ptc2Smem {
  static constexpr int STAGES = 6;
  static constexpr int A_BYTES = 128 * 128;      // this CTA's 128 pixels x 64 channels
  static constexpr int B_BYTES = 128 * 128;      // this CTA's half of the 256-channel filter tile
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int SS_OFF = STAGES * STAGE_BYTES;                 // scale/shift floats [2][512]
  static constexpr int BAR_OFF = SS_OFF + 2 * 512 * 4;                // full[S], empty[S], tfull[2], tempty[2]
  static constexpr int SLOT_OFF = BAR_OFF + (2 * STAGES + 4) * 8;
  static constexpr int TOTAL = SLOT_OFF + 16 + 1024;
};

template <bool POOL, bool F16 = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(192, 1)
ptc2_conv_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TcArgs a) {
  using L = Ptc2Smem;
  constexpr int BN = 256, BKE = 64, STAGES = L::STAGES;
  constexpr int TMEM_COLS = 512;

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
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
  const int n_tiles = a.cout / BN;
  const int lg_bi = 7 - a.lg_bw - a.lg_bh;
  const int m_tiles = a.tiles_w * a.tiles_h * ((a.n_img + (1 << lg_bi) - 1) >> lg_bi);
  const int num_items = ((m_tiles + 1) >> 1) * n_tiles;      // item = (pair of pixel tiles, channel tile)
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
    mbar_init(bar_tempty, 8);          // one arrival per epilogue warp of both CTAs (only the leader's copy is used)
    mbar_init(bar_tempty + 8, 8);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc_2cta<TMEM_COLS>(smem_u32(const_cast<uint32_t*>(tmem_slot)));
  if (warp >= 2)
    for (int i = threadIdx.x - 64; i < a.cout; i += 128) {
      ss[2 * i] = a.scale[i];          // interleaved (scale, shift) pairs: the epilogue reads two channels per 16-byte load
      ss[2 * i + 1] = a.shift[i];
    }
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();                  // barrier inits + TMEM allocation visible to the peer
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // item -> channel tile and THIS CTA's pixel tile (a pixel tile past the end lands on an image index beyond the
  // tensor: TMA zero-fills, the epilogue masks it by n < n_img)
  auto item_coords = [&](int item, int* w0, int* h0, int* n0, int* col0) {
    const int nt = item % n_tiles, mt = (item / n_tiles) * 2 + static_cast<int>(rank);
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
      int s = 0, ph = 0;
      for (int item = cluster_id; item < num_items; item += num_clusters) {
        int w0, h0, n0, col0;
        item_coords(item, &w0, &h0, &n0, &col0);
        for (int kb = 0; kb < kb_total; ++kb) {
          mbar_wait(bar_empty + 8 * s, ph ^ 1);
          const uint32_t sa = base + s * L::STAGE_BYTES;
          const uint32_t bar = bar_full + 8 * s;
          if (leader) mbar_arrive_expect_tx(bar, 2 * L::STAGE_BYTES);      // own + peer's bytes
          const int tap = kb / a.kb_per_tap;
          const int cc = kb - tap * a.kb_per_tap;
          const int kh = tap / 3, kw = tap - kh * 3;
          tma_load_4d_2cta(sa, &tmA, bar, cc * BKE, w0 + kw - 1, h0 + kh - 1, n0);
          tma_load_2d_2cta(sa + L::A_BYTES, &tmB, bar, kb * BKE, col0 + 128 * static_cast<int>(rank));
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // Whole warp in uniform control flow, one elected lane around the MMAs: under `if (lane == 0)` the stage-dependent
    // descriptors live in vector registers and every UTCHMMA is wrapped in an ELECT + 8x R2UR.BROADCAST loop.
    if (leader) {
      constexpr uint32_t idesc = make_idesc_16<F16>(256, BN);
      int s = 0, ph = 0, it = 0;
      for (int item = cluster_id; item < num_items; item += num_clusters, ++it) {
        const int acc = it & 1;
        if (it >= 2) mbar_wait(bar_tempty + 8 * acc, ((it >> 1) - 1) & 1);
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < kb_total; ++kb) {
          mbar_wait(bar_full + 8 * s, ph);
          tcgen05_fence_after();
          const uint32_t sa = base + s * L::STAGE_BYTES;
          const uint64_t adesc = make_kmajor_desc<128>(sa);
          const uint64_t bdesc = make_kmajor_desc<128>(sa + L::A_BYTES);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16_ss_2cta(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
            umma_commit_2cta(bar_empty + 8 * s, 0x3);
          }
          __syncwarp();
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
        if (elect_one()) umma_commit_2cta(bar_tfull + 8 * acc, 0x3);
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
    for (int item = cluster_id; item < num_items; item += num_clusters, ++it) {
      const int acc = it & 1;
      mbar_wait(bar_tfull + 8 * acc, (it >> 1) & 1);
      tcgen05_fence_after();
      int w0, h0, n0, col0;
      item_coords(item, &w0, &h0, &n0, &col0);
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(g * 32) << 16) + acc * BN;
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
        if (c0 + 32 == BN) {             // accumulator fully read: hand the TMEM buffer back to the leader's MMA thread
          tcgen05_fence_before();
          __syncwarp();                  // one cluster-scope release per warp (a MEMBAR each), not one per thread
          if (lane == 0) mbar_arrive_cluster(bar_tempty + 8 * acc, 0);
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

  tcgen05_fence_before();
  cluster_sync_all();                  // the peer's smem / TMEM stay alive until the leader's last MMA has retired
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc_2cta<TMEM_COLS>(tmem_base);
  }
}

template <bool POOL, bool F16 = false>
inline cudaError_t launch_ptc2(int grid, cudaStream_t st, const CUtensorMap& a, const CUtensorMap& b, const TcArgs& args) {
  return ffh::launch_smem(ptc2_conv_kernel<POOL, F16>, dim3(grid), dim3(192), Ptc2Smem::TOTAL, st, true, a, b, args);
}

}  // namespace ff
