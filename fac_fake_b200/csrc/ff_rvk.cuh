// ff_rvk.cuh — kernels of ResVitKan inference (SURVEY.md §8f-1;
// /root/reference/CViT-main/ResVitKan/ResVitKan.py:185-240,284-329 and kan.py:90-206): input conversion, the stem
// (7x7 stride 2, no im2col), the 3x3 stride-2 max-pool, rvk_conv2_kernel (every bottleneck convolution; also used by
// the S3D and GGCA engines), the KAN head and the GGCA gate.
#pragma once
#include <type_traits>
#include "ff_c1.cuh"
#include "ff_small.cuh"

namespace ff {

// ---- input -> normalised bf16 NHWC4 [n,224,224,4] (channel 3 = 0), 8 bytes per pixel.
// IN_KIND 2: uint8 NHWC with (x/255-mean)/std (cvit_prediction.py:41-45 convention); IN_KIND 0: fp32 NCHW as given.
template <int IN_KIND, bool F16 = false>
__global__ void __launch_bounds__(256)
rvk_convert_kernel(const void* __restrict__ x, __nv_bfloat16* __restrict__ out, int n_img, float a0, float b0, float a1, float b1,
                   float a2, float b2) {
  const size_t total = static_cast<size_t>(n_img) * 224 * 224;
  if (IN_KIND == 2) {
    // 4 pixels per thread: 12 input bytes as three aligned 32-bit loads, 32 output bytes as two 16-byte stores
    const size_t i = (blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x) * 4;
    if (i >= total) return;                        // 224*224 is a multiple of 4: a quad never straddles the end
    const uint32_t* p = reinterpret_cast<const uint32_t*>(reinterpret_cast<const uint8_t*>(x) + i * 3);
    const uint32_t w0 = p[0], w1 = p[1], w2 = p[2];
    const uint32_t by[12] = {w0 & 255u, (w0 >> 8) & 255u, (w0 >> 16) & 255u, w0 >> 24, w1 & 255u, (w1 >> 8) & 255u,
                             (w1 >> 16) & 255u, w1 >> 24, w2 & 255u, (w2 >> 8) & 255u, (w2 >> 16) & 255u, w2 >> 24};
    uint32_t o[8];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float v0 = fmaf(static_cast<float>(by[3 * k]), a0, b0);
      const float v1 = fmaf(static_cast<float>(by[3 * k + 1]), a1, b1);
      const float v2 = fmaf(static_cast<float>(by[3 * k + 2]), a2, b2);
      o[2 * k] = pack16x2<F16>(v0, v1);
      o[2 * k + 1] = pack16x2<F16>(v2, 0.0f);
    }
    uint4* q = reinterpret_cast<uint4*>(out + i * 4);
    q[0] = make_uint4(o[0], o[1], o[2], o[3]);
    q[1] = make_uint4(o[4], o[5], o[6], o[7]);
  } else {
    const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
    if (i >= total) return;
    const size_t img = i / (224 * 224), pix = i % (224 * 224);
    const float* p = reinterpret_cast<const float*>(x) + img * 3 * 224 * 224 + pix;
    reinterpret_cast<uint2*>(out)[i] = make_uint2(pack16x2<F16>(p[0], p[224 * 224]), pack16x2<F16>(p[2 * 224 * 224], 0.0f));
  }
}

// ---- stem: Conv2d(3,64,7,stride 2,pad 3) + BN + ReLU (ResVitKan.py:191-193,229-231) with NO im2col.
// Output pixel w needs input pixels 2w-3..2w+3.  The NHWC4 patch (24 pixels x 37 rows, TMA, zero-filled borders) is
// read by non-swizzled K-major descriptors: row = output pixel (16 bytes = 2 input pixels apart), K window = the 8
// input pixels 2w-4..2w+3 (64 contiguous bytes = two K=16 steps; the first pixel meets zero weights), filter row kh
// = patch row 2*h_l + kh.  Tile = 8 x 16 output pixels, 14 tcgen05.mma (M=128, N=64) per tile.
struct RvkStemArgs {
  __nv_bfloat16* out;            // [n,112,112,64] bf16
  const __nv_bfloat16* w;        // [7 kh][64 cout][32 = 8 px x 4 ch] bf16 (px 0 and ch 3 are zero)
  int n_img;
  float scale[64];
  float shift[64];
};

constexpr int RVK_STEM_RING = 3, RVK_STEM_PSLOT = 7168;
constexpr int RVK_STEM_SMEM = RVK_STEM_RING * RVK_STEM_PSLOT + 7 * 4096 + 128;
static __global__ void __launch_bounds__(128, 4)
rvk_stem_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ RvkStemArgs a) {
  constexpr int OW = 112, TW = 8, TH = 16, TILES_W = OW / TW, TILES_H = OW / TH, TILES = TILES_W * TILES_H;
  constexpr int RING = RVK_STEM_RING, PROW = 192, PROWS = 37, PBYTES = PROWS * PROW, PSLOT = RVK_STEM_PSLOT;
  extern __shared__ uint8_t rvk_smem_raw[];
  uint8_t* sm = rvk_smem_raw + ((128u - (smem_u32(rvk_smem_raw) & 127u)) & 127u);
  uint8_t (*s_patch)[PSLOT] = reinterpret_cast<uint8_t (*)[PSLOT]>(sm);
  uint8_t* sB = sm + RING * PSLOT;
  __shared__ __align__(8) uint64_t s_bar[1 + RING];
  __shared__ uint32_t s_tmem;
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t bar_mma = smem_u32(&s_bar[0]);
  const uint32_t bar_raw = smem_u32(&s_bar[1]);
  // filter -> core-matrix layout: (n, 16-byte chunk c) at ((n/8)*4 + c)*128 + (n%8)*16
  for (int i = tid; i < 7 * 64 * 4; i += 128) {
    const int kh = i / 256, rem = i % 256, n = rem >> 2, c = rem & 3;
    *reinterpret_cast<uint4*>(sB + kh * 4096 + ((n >> 3) * 4 + c) * 128 + (n & 7) * 16) = reinterpret_cast<const uint4*>(a.w)[i];
  }
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
  const uint32_t sB_addr = smem_u32(sB);
  constexpr uint32_t idesc = make_idesc_bf16(128, 64);
  const int num_tiles = TILES * a.n_img;
  const int hl = tid >> 3, wl = tid & 7;
  auto issue = [&](int t, int slot) {
    const int n = t / TILES;
    const int rem = t - n * TILES;
    const int th = rem / TILES_W, tw = rem - th * TILES_W;
    mbar_arrive_expect_tx(bar_raw + 8 * slot, PBYTES);
    // patch = input pixels 2*w0-4 .. 2*w0+19 (96 bf16 = 192 B, 16-byte aligned start), rows 2*h0-3 .. 2*h0+33
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
            smem_u32(&s_patch[slot][0])),
        "l"(reinterpret_cast<uint64_t>(&tmX)), "r"(bar_raw + 8 * slot), "r"((2 * tw * TW - 4) * 4), "r"(2 * th * TH - 3), "r"(n)
        : "memory");
  };
  if (tid == 0)
    for (int s = 0; s < RING - 1; ++s) {
      const int t = blockIdx.x + s * gridDim.x;
      if (t < num_tiles) issue(t, s);
    }
  int it = 0;
  for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
    const int n = t / TILES;
    const int rem = t - n * TILES;
    const int th = rem / TILES_W, tw = rem - th * TILES_W;
    const int slot = it % RING;
    if (tid == 0) {
      const int tn = t + (RING - 1) * gridDim.x;        // slot (it+2)%3 was consumed by tile it-1 (all threads passed bar_mma)
      if (tn < num_tiles) issue(tn, (it + RING - 1) % RING);
      mbar_wait(bar_raw + 8 * slot, (it / RING) & 1);
      tcgen05_fence_after();
      const uint32_t patch = smem_u32(&s_patch[slot][0]);
#pragma unroll
      for (int kh = 0; kh < 7; ++kh) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          // A: row (h_l, w_l) at patch + (2*h_l + kh)*192 + w_l*16; K chunks 16 B apart; groups (h_l) 2 patch rows apart
          const uint64_t ad = make_kmajor_desc_noswz(patch + kh * PROW + 32 * j, 16, 2 * PROW);
          const uint64_t bd = make_kmajor_desc_noswz(sB_addr + kh * 4096 + 2 * j * 128, 128, 512);
          umma_bf16_ss(tmem, ad, bd, idesc, (kh > 0 || j > 0) ? 1u : 0u);
        }
      }
      umma_commit(bar_mma);
    }
    mbar_wait(bar_mma, it & 1);
    tcgen05_fence_after();
    __nv_bfloat16* o = a.out + ((static_cast<size_t>(n) * OW + (th * TH + hl)) * OW + (tw * TW + wl)) * 64;
#pragma unroll
    for (int p = 0; p < 2; ++p) {
      uint32_t v[32];
      tmem_ld_32x32(tmem + (static_cast<uint32_t>(warp * 32) << 16) + p * 32, v);
      tmem_ld_wait();
      uint32_t pk[16];
#pragma unroll
      for (int c = 0; c < 32; c += 2) {
        const float x0 = fmaf(__uint_as_float(v[c]), a.scale[p * 32 + c], a.shift[p * 32 + c]);
        const float x1 = fmaf(__uint_as_float(v[c + 1]), a.scale[p * 32 + c + 1], a.shift[p * 32 + c + 1]);
        pk[c >> 1] = pack_bf16x2_relu(x0, x1);
      }
      st_global_v8(o + p * 32, pk);
      st_global_v8(o + p * 32 + 16, pk + 8);
    }
    tcgen05_fence_before();
    __syncthreads();               // every thread has read TMEM and passed bar_mma before the next tile's MMAs / TMA reuse
  }
  __syncthreads();
  if (warp == 0) { tcgen05_fence_after(); tmem_dealloc<64>(tmem); }
}


// ---- persistent implicit-GEMM convolution for the ResNet bottlenecks (1x1 / 3x3, stride 1 / 2), the S3D convolutions
// and the GGCA variant's BN-less conv.  Same skeleton as ptc2_conv_kernel (ff_ptc2.cuh): one CTA per SM walks (2 pixel
// tiles x BN channels) work items, warp 0 = TMA producer, warp 1 = tcgen05 issuer, double-buffered TMEM accumulators;
//   * the tap loop and the TMA coordinates follow a.taps / a.stride (the strided maps carry elementStrides);
//   * a stride-1 1x1 convolution is run "flat": W = all pixels of the batch, H = N = 1, boxes of 128 pixels;
//   * 8 epilogue warps (two per TMEM lane group, each owning half of the BN columns): the bottleneck outputs are
//     HBM-bound (K is as small as 64), so ALL global traffic is on TMA.  ncu on the register-direct predecessor
//     (profiles/r01_ncu_rvk_conv.txt): per-thread 32-byte residual loads / output stores ran at 32 sectors per
//     request and held the LSU at 71 % while DRAM sat at 50 %.  Here the epilogue converts a sub-tile
// in a swizzled shared-memory panel ([128 pixels][64 channels] bf16, the TMA SW128 layout) and one elected thread
// stores it with cp.async.bulk.tensor; the residual arrives in the same panel by TMA (LA sub-tiles ahead) and is
// overwritten in place.  NSTG panels-sets rotate; set b is owned by elected thread E_b, which is the only one that
// stores from it and loads into it, so "my previous store has finished reading" is a plain bulk-group wait.
template <int BN, int STAGES, bool RESID>
struct Rvk2Smem {
  static constexpr int A_BYTES = 2 * 128 * 128;
  static constexpr int B_BYTES = BN * 128;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int NSTG = RESID ? 3 : 2;
  static constexpr int STG_BYTES = 128 * BN * 2;                      // one 128-pixel sub-tile: BN/64 panels of 16 KB
  static constexpr int STG_OFF = STAGES * STAGE_BYTES;
  static constexpr int SS_OFF = STG_OFF + NSTG * STG_BYTES;           // scale/shift floats [2][2048]
  static constexpr int BAR_OFF = SS_OFF + 2 * 2048 * 4;               // full[S], empty[S], tfull[2], tempty[2], res_full[3]
  static constexpr int SLOT_OFF = BAR_OFF + (2 * STAGES + 4 + 3) * 8;
  static constexpr int TOTAL = SLOT_OFF + 16 + 1024;
  static_assert(TOTAL <= 232448, "shared memory budget");
};

template <int BN, int STAGES, bool RESID, bool F16 = false>
__global__ void __launch_bounds__(320, 1)
rvk_conv2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmR, const TcArgs a) {
  using L = Rvk2Smem<BN, STAGES, RESID>;
  constexpr int MSUB = 2, BKE = 64, NSTG = L::NSTG, LA = NSTG - 1;
  constexpr int TMEM_COLS = 2 * MSUB * BN;
  constexpr int HALF = BN / 2, NCH = HALF / 32;
  constexpr int PANELS = BN / 64;
  static_assert(BN == 64 || BN == 128, "BN");
  static_assert(!(RESID && F16), "the residual epilogue reads bf16");

  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
  float* ss = reinterpret_cast<float*>(base_ptr + L::SS_OFF);
  const uint32_t bar_full = base + L::BAR_OFF;
  const uint32_t bar_empty = bar_full + STAGES * 8;
  const uint32_t bar_tfull = bar_empty + STAGES * 8;
  const uint32_t bar_tempty = bar_tfull + 16;
  const uint32_t bar_res = bar_tempty + 16;
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
    tma_prefetch_desc(&tmO);
    if (RESID) tma_prefetch_desc(&tmR);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    mbar_init(bar_tfull, 1);
    mbar_init(bar_tfull + 8, 1);
    mbar_init(bar_tempty, 256);
    mbar_init(bar_tempty + 8, 256);
    for (int b = 0; b < 3; ++b) mbar_init(bar_res + 8 * b, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<TMEM_COLS>(smem_u32(const_cast<uint32_t*>(tmem_slot)));
  if (warp >= 2)
    for (int i = threadIdx.x - 64; i < a.cout; i += 256) {
      ss[i] = a.scale[i];
      ss[2048 + i] = a.shift[i];
    }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) pdl_trigger();
  pdl_wait();

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
      const int sd = a.stride == 2 ? 2 : 1;
      int s = 0, ph = 0;
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
          // taps: 9 = 3x3 (pad 1); 28 = 7x7/2 over pixel pairs; 1 = 1x1; 3 / 7 = a 1-D filter along the H coordinate (pad 1 / 3), which is the time
          // axis of the S3D temporal convolutions — their stride applies to that axis only
          int dx = 0, dy = 0, sdw = sd;
          if (a.taps == 9) { const int kh = tap / 3; dy = kh - 1; dx = tap - kh * 3 - 1; }
          // 28 = 7x7 / stride 2 on a 32-channel tensor viewed as pixel PAIRS of 64 channels (S3D SRM stem): output pixel x
          // reads input pixels 2x-3 .. 2x+3 = pairs x-2 (second half) .. x+1, so the pair axis has unit stride
          else if (a.taps == 28) { const int kh = tap >> 2; dy = kh - 3; dx = (tap & 3) - 2; sdw = 1; }
          else if (a.taps == 3 || a.taps == 7) { dy = tap - (a.taps >> 1); sdw = 1; }
#pragma unroll
          for (int j = 0; j < MSUB; ++j) tma_load_4d(sa + j * 128 * 128, &tmA, bar, cc * BKE, sdw * w0[j] + dx, sd * h0[j] + dy, n0[j]);
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
    const int cb = ((warp - 2) >> 2) * HALF;     // this thread's first column inside the BN tile
    const int r = g * 32 + lane;                 // accumulator row == TMEM lane == pixel of the sub-tile
    const int et = threadIdx.x - 64;             // 0..255
    const int my_set = (et & 31) == 0 && (et >> 5) < NSTG ? (et >> 5) : -1;   // E_b = first lane of epilogue warp b
    const bool relu1 = a.conv_act == 0;
    const int total_q = ((num_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x) * MSUB;   // sub-tiles of this CTA
    // panel holding this thread's columns, and the byte offset of its row inside a panel
    const int panel = cb / 64, chunk0 = (cb % 64) / 8;      // 16-byte chunk index of the first column
    const uint32_t row_off = r * 128;
    const int rsw = r & 7;
    auto issue_resid = [&](int q) {              // by E_{q % NSTG}: residual of sub-tile q -> staging set q % NSTG
      if (q >= total_q) return;
      const int t = blockIdx.x + (q >> 1) * gridDim.x;
      int w0, h0, n0, col0;
      tile_coords(t, q & 1, &w0, &h0, &n0, &col0);
      const uint32_t dst = base + L::STG_OFF + (q % NSTG) * L::STG_BYTES;
      const uint32_t bar = bar_res + 8 * (q % NSTG);
      mbar_arrive_expect_tx(bar, L::STG_BYTES);
#pragma unroll
      for (int p = 0; p < PANELS; ++p) tma_load_4d(dst + p * 16384, &tmR, bar, col0 + p * 64, w0, h0, n0);
    };
    if (RESID && my_set >= 0 && my_set < LA) issue_resid(my_set);
    int it = 0, q = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
      const int acc = it & 1;
      mbar_wait(bar_tfull + 8 * acc, (it >> 1) & 1);
      tcgen05_fence_after();
#pragma unroll 1
      for (int j = 0; j < MSUB; ++j, ++q) {
        int w0, h0, n0, col0;
        tile_coords(t, j, &w0, &h0, &n0, &col0);
        const int set = q % NSTG;
        uint8_t* stg = base_ptr + L::STG_OFF + set * L::STG_BYTES + panel * 16384 + row_off;
        if (RESID) mbar_wait(bar_res + 8 * set, (q / NSTG) & 1);
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(g * 32) << 16) + (acc * MSUB + j) * BN + cb;
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
          uint32_t v[32];
          tmem_ld_32x32(taddr + ch * 32, v);
          tmem_ld_wait();
          if (j == MSUB - 1 && ch == NCH - 1) {
            tcgen05_fence_before();
            mbar_arrive(bar_tempty + 8 * acc);
          }
          const float* sc = ss + col0 + cb + ch * 32;
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) {       // four 16-byte chunks = 32 columns
            const int cidx = chunk0 + ch * 4 + k4;
            uint4* sp = reinterpret_cast<uint4*>(stg + ((cidx ^ rsw) << 4));
            uint32_t rr[4] = {0u, 0u, 0u, 0u};
            if (RESID) { const uint4 rv = *sp; rr[0] = rv.x; rr[1] = rv.y; rr[2] = rv.z; rr[3] = rv.w; }
            uint32_t p[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int c = k4 * 8 + e * 2;
              float x0 = fmaf(__uint_as_float(v[c]), sc[c], sc[2048 + c]);
              float x1 = fmaf(__uint_as_float(v[c + 1]), sc[c + 1], sc[2048 + c + 1]);
              if (relu1) { x0 = fmaxf(x0, 0.0f); x1 = fmaxf(x1, 0.0f); }
              if (RESID) {
                const float2 rf = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&rr[e]));
                x0 = fmaxf(x0 + rf.x, 0.0f);
                x1 = fmaxf(x1 + rf.y, 0.0f);
              }
              p[e] = pack16x2<F16>(x0, x1);
            }
            *sp = make_uint4(p[0], p[1], p[2], p[3]);
          }
        }
        fence_proxy_async_smem();                // generic-proxy writes of this thread -> visible to the TMA store
        // E of the set that sub-tile q+LA will use: its last store (sub-tile q+LA-NSTG = q-1) was issued one sub-tile
        // ago; once it has finished reading, the set is free (for the next residual load / for direct writes)
        if (my_set == (q + LA) % NSTG) {
          bulk_wait_group_read<0>();
          if (RESID) issue_resid(q + LA);
        }
        named_bar_sync(1, 256);
        if (my_set == set) {
          const uint32_t src = base + L::STG_OFF + set * L::STG_BYTES;
#pragma unroll
          for (int p = 0; p < PANELS; ++p) tma_store_4d(&tmO, src + p * 16384, col0 + p * 64, w0, h0, n0);
          bulk_commit_group();
        }
      }
    }
    if (my_set >= 0) bulk_wait_group<0>();       // stores performed before the CTA retires
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

// pixel-tile box (bw x bh pixels x bi images = 128 rows) of rvk_conv2_kernel per output size
inline void rvk_tile_geometry(int out_hw, int* bw, int* bh, int* bi) {
  if (out_hw == 56) { *bw = 8; *bh = 8; *bi = 2; }
  else if (out_hw == 28) { *bw = 4; *bh = 4; *bi = 8; }
  else if (out_hw == 14) { *bw = 2; *bh = 2; *bi = 32; }
  else { *bw = 8; *bh = 8; *bi = 2; }                         // 7x7: one masked 8x8 box per image
}

template <int BN, int STAGES, bool RESID, bool F16 = false>
inline cudaError_t launch_rvk_conv2(int grid, cudaStream_t st, const CUtensorMap& a, const CUtensorMap& b, const CUtensorMap& o,
                                    const CUtensorMap& r, const TcArgs& args) {
  return ffh::launch_smem(rvk_conv2_kernel<BN, STAGES, RESID, F16>, dim3(grid), dim3(320), Rvk2Smem<BN, STAGES, RESID>::TOTAL, st, true, a, b, o, r, args);
}

// ---- MaxPool2d(kernel 3, stride 2, pad 1) on NHWC bf16 with C = 64 (ResVitKan.py:194,232): [n,112,112,64] -> [n,56,56,64]
static __global__ void __launch_bounds__(256)
rvk_maxpool_kernel(const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ out, int n_img) {
  constexpr int IW = 112, OW = 56, C8 = 8;     // 8 chunks of 8 channels
  const size_t total = static_cast<size_t>(n_img) * OW * OW * C8;
  const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (i >= total) return;
  const int q = static_cast<int>(i % C8);
  size_t t = i / C8;
  const int ow = static_cast<int>(t % OW); t /= OW;
  const int oh = static_cast<int>(t % OW);
  const size_t n = t / OW;
  uint4 m = make_uint4(0xff80ff80u, 0xff80ff80u, 0xff80ff80u, 0xff80ff80u);   // bf16 -inf pairs
  auto mx = [](uint32_t a0, uint32_t b0) {
    __nv_bfloat162 r2 = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&a0), *reinterpret_cast<__nv_bfloat162*>(&b0));
    return *reinterpret_cast<uint32_t*>(&r2);
  };
  for (int dy = -1; dy <= 1; ++dy) {
    const int ih = 2 * oh + dy;
    if (ih < 0 || ih >= IW) continue;
    for (int dx = -1; dx <= 1; ++dx) {
      const int iw = 2 * ow + dx;
      if (iw < 0 || iw >= IW) continue;
      const uint4 v = *reinterpret_cast<const uint4*>(in + ((n * IW + ih) * IW + iw) * 64 + q * 8);
      m.x = mx(m.x, v.x); m.y = mx(m.y, v.y); m.z = mx(m.z, v.z); m.w = mx(m.w, v.w);
    }
  }
  *reinterpret_cast<uint4*>(out + ((n * OW + oh) * OW + ow) * 64 + q * 8) = m;
}

// ---- KAN head: KAN([2048, 64, 2]) on the ReLU'd hidden vector (ResVitKan.py:302-307, kan.py:189-206).
// KANLinear(x) = silu(x) W_base^T + Bspline(x) W_spline_scaled^T, cubic B-splines on each input feature's 12 knots
// (g0[2048][12], g1[64][12]) with half-open intervals (kan.py:115), Cox-de Boor recursion in the reference's order
// (kan.py:116-126).  Weights are pre-packed on the host as w0[in=2048][9][64] and w1[in=64][9][2]
// (slot 0 = base_weight, 1..8 = spline_weight * spline_scaler).
__device__ __forceinline__ void kan_bases(float x, const float* __restrict__ grid, float b[8]) {
  float g[12];
#pragma unroll
  for (int i = 0; i < 12; ++i) g[i] = grid[i];     // this input feature's knots (kan.py:44-52; update_grid may move them)
  float t[11];
#pragma unroll
  for (int i = 0; i < 11; ++i) t[i] = (x >= g[i] && x < g[i + 1]) ? 1.0f : 0.0f;
#pragma unroll
  for (int k = 1; k <= 3; ++k) {
#pragma unroll
    for (int i = 0; i < 11 - k; ++i)
      t[i] = (x - g[i]) / (g[i + k] - g[i]) * t[i] + (g[i + k + 1] - x) / (g[i + k + 1] - g[i + 1]) * t[i + 1];
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) b[i] = t[i];
}
__device__ __forceinline__ float silu_f(float x) { return x / (1.0f + expf(-x)); }

// Layer 0 (2048 -> 64) as a split-K product of the [n][2048*9] feature matrix (silu + 8 bases per input, built on
// the fly in shared memory) with w0: block = 64 inputs x 32 samples, so w0 is streamed once per 32 samples instead
// of once per sample; the 32 K-chunks write private slabs part[chunk][n_cap][64] that layer 1 sums in a fixed order.
constexpr int KAN_KC = 64, KAN_SUB = 16, KAN_SG = 32, KAN_CHUNKS = 2048 / KAN_KC;
static __global__ void __launch_bounds__(256)
kan_l0_kernel(const float* __restrict__ hid, const float* __restrict__ w0, const float* __restrict__ g0,
              float* __restrict__ part, int n, int n_cap) {
  __shared__ float s_feat[KAN_SG][KAN_SUB * 9 + 1];
  const int tid = threadIdx.x;
  const int i_base = blockIdx.x * KAN_KC, s_base = blockIdx.y * KAN_SG;
  const int o = tid & 63, q = tid >> 6;
  float acc[8];
#pragma unroll
  for (int s = 0; s < 8; ++s) acc[s] = 0.0f;
  for (int sub = 0; sub < KAN_KC / KAN_SUB; ++sub) {
    const int i0 = i_base + sub * KAN_SUB;
#pragma unroll
    for (int rep = 0; rep < KAN_SG * KAN_SUB / 256; ++rep) {
      const int idx = tid + rep * 256;
      const int s = idx / KAN_SUB, ii = idx % KAN_SUB;
      const int smp = s_base + s;
      const float x = smp < n ? hid[static_cast<size_t>(smp) * 2048 + i0 + ii] : 0.0f;
      float f[9];
      f[0] = silu_f(x);
      kan_bases(x, g0 + (i0 + ii) * 12, f + 1);
#pragma unroll
      for (int k = 0; k < 9; ++k) s_feat[s][ii * 9 + k] = f[k];
    }
    __syncthreads();
    const float* w = w0 + static_cast<size_t>(i0) * 9 * 64 + o;
#pragma unroll 4
    for (int kk = 0; kk < KAN_SUB * 9; ++kk) {
      const float wv = w[kk * 64];
#pragma unroll
      for (int s = 0; s < 8; ++s) acc[s] = fmaf(s_feat[q * 8 + s][kk], wv, acc[s]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int s = 0; s < 8; ++s) {
    const int smp = s_base + q * 8 + s;
    if (smp < n) part[(static_cast<size_t>(blockIdx.x) * n_cap + smp) * 64 + o] = acc[s];
  }
}

// Layer 1 (64 -> 2): one warp per sample; sums the K-chunk slabs of layer 0, then silu / B-spline features of the 64
// hidden values against w1.
static __global__ void __launch_bounds__(256)
kan_l1_kernel(const float* __restrict__ part, const float* __restrict__ w1, const float* __restrict__ g1,
              float* __restrict__ logits, int n, int n_cap) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x * 8 + warp;
  if (b >= n) return;
  float o0 = 0.0f, o1 = 0.0f;
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int i = lane + 32 * half;
    float x = 0.0f;
    for (int c = 0; c < KAN_CHUNKS; ++c) x += part[(static_cast<size_t>(c) * n_cap + b) * 64 + i];
    float f[9];
    f[0] = silu_f(x);
    kan_bases(x, g1 + i * 12, f + 1);
    const float* w = w1 + i * 9 * 2;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      o0 = fmaf(f[k], w[k * 2], o0);
      o1 = fmaf(f[k], w[k * 2 + 1], o1);
    }
  }
  o0 = warp_sum(o0);
  o1 = warp_sum(o1);
  if (lane == 0) {
    logits[2 * b] = o0;
    logits[2 * b + 1] = o1;
  }
}

// ---- GGCA(512, 7, 7) gate of the cvit_GGCA_ADD_DEConv_RepBn8 variant (SURVEY.md §8f-4;
// /root/reference/CViT-main/model/cvit_GGCA_ADD_DEConv_RepBn8.py:143-213 and forward :447-448):
//   att_h[c][h] = sigmoid(S(mean_w x) + S(max_w x)),  att_w[c][w] likewise over h,  S = shared 1x1 convs per group of
//   128 channels: 128 -> 8 (+BN, folded on the host) -> ReLU -> 128;  out = x * (x * att_h * att_w).
// One block per crop, one thread per channel; the 7x7 map of the thread's channel lives in registers, the pooled
// vectors and the hidden units go through shared memory.  In place on the NHWC feature map [n][7][7][512]: read in the
// conv stack's 16-bit type (fp16 when in_f16 != 0), written as bf16 — the patch-embedding GEMM that follows is bf16.
// KIND 0: bf16 map;  1: fp16 in, bf16 out;  2: fp32 map (FF_COMPUTE_FP32 path).
template <int KIND>
__global__ void __launch_bounds__(512)
ggca_gate_kernel(void* __restrict__ feat_v, const float* __restrict__ w1, const float* __restrict__ b1,
                 const float* __restrict__ w2, const float* __restrict__ b2, int n) {
  using T = typename std::conditional<KIND == 2, float, __nv_bfloat16>::type;
  T* feat = reinterpret_cast<T*>(feat_v);
  __shared__ float s_pool[7][512];             // pooled vectors of one type: [pos][channel]
  __shared__ float s_hid[28][4][8];            // [type*7 + pos][group][hidden unit]; type 0 = h_avg, 1 = h_max, 2 = w_avg, 3 = w_max
  const int b = blockIdx.x;
  if (b >= n) return;
  const int c = threadIdx.x, grp = c >> 7, cg = c & 127;
  T* f = feat + static_cast<size_t>(b) * 49 * 512 + c;
  float x[49];
#pragma unroll
  for (int i = 0; i < 49; ++i) {
    if constexpr (KIND == 2) x[i] = f[i * 512];
    else if constexpr (KIND == 1) x[i] = __half2float(reinterpret_cast<const __half*>(f)[i * 512]);
    else x[i] = __bfloat162float(f[i * 512]);
  }
#pragma unroll
  for (int type = 0; type < 4; ++type) {
#pragma unroll
    for (int r = 0; r < 7; ++r) {
      float sum = 0.f, mx = -INFINITY;
#pragma unroll
      for (int k = 0; k < 7; ++k) {
        const float v = type < 2 ? x[r * 7 + k] : x[k * 7 + r];   // types 0/1: row r pooled over w; 2/3: column r over h
        sum += v;
        mx = fmaxf(mx, v);
      }
      s_pool[r][c] = (type & 1) ? mx : sum / 7.0f;
    }
    __syncthreads();
    // hidden units of this type: 7 positions x 4 groups x 8 units = 224 dot products of length 128
    if (threadIdx.x < 224) {
      const int u = threadIdx.x & 7, g = (threadIdx.x >> 3) & 3, pos = threadIdx.x >> 5;
      const float* pv = &s_pool[pos][g * 128];
      const float* pw = w1 + u * 128;
      float acc = b1[u];
      for (int k = 0; k < 128; ++k) acc = fmaf(pw[k], pv[k], acc);
      s_hid[type * 7 + pos][g][u] = fmaxf(acc, 0.0f);
    }
    __syncthreads();
  }
  float att[14];                               // att_h[0..6], att_w[0..6] of this channel
  float wv[8];
#pragma unroll
  for (int u = 0; u < 8; ++u) wv[u] = w2[cg * 8 + u];
  const float bb = b2[cg];
#pragma unroll
  for (int d = 0; d < 2; ++d)
#pragma unroll
    for (int r = 0; r < 7; ++r) {
      float ya = bb, ym = bb;
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        ya = fmaf(wv[u], s_hid[d * 14 + r][grp][u], ya);
        ym = fmaf(wv[u], s_hid[d * 14 + 7 + r][grp][u], ym);
      }
      att[d * 7 + r] = 1.0f / (1.0f + expf(-(ya + ym)));
    }
#pragma unroll
  for (int hh = 0; hh < 7; ++hh)
#pragma unroll
    for (int ww = 0; ww < 7; ++ww) {
      const float v = x[hh * 7 + ww];
      const float g = v * (v * att[hh] * att[7 + ww]);
      if constexpr (KIND == 2) f[(hh * 7 + ww) * 512] = g;
      else f[(hh * 7 + ww) * 512] = __float2bfloat16(g);
    }
}

}  // namespace ff
