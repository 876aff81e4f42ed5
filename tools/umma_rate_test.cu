// Microbenchmark: cycles per tcgen05.mma (cta_group::1, kind::f16, M=128, K=16, SS mode) as a function of N,
// swizzle width (SW64 / SW128), the A-side stride-byte-offset (8 rows = dense tile, 10 rows = halo patch) and the
// number of CTAs.  Operands are whatever is in shared memory (zeros); only timing matters.
#include <cstdio>
#include <vector>
#include "../fac_fake_b200/csrc/ff_ptx.cuh"
using namespace ff;

template <int ROWB, int BN>
__global__ void __launch_bounds__(128, 1) rate_kernel(int iters, int sbo_rows, int ksteps, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* bp = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t sa = base, sb = base + 200 * ROWB + 1024;
  const uint32_t sb_al = (sb + 1023u) & ~1023u;
  const uint32_t bar = sb_al + 256 * ROWB;
  volatile uint32_t* slot = reinterpret_cast<volatile uint32_t*>(bp + (bar - base) + 16);
  for (int i = threadIdx.x; i < (int)((bar - base) / 4); i += blockDim.x) reinterpret_cast<uint32_t*>(bp)[i] = 0;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  if (warp == 1) tmem_alloc<256>(smem_u32(const_cast<uint32_t*>(slot)));
  fence_proxy_async_smem();
  tcgen05_fence_before(); __syncthreads(); tcgen05_fence_after();
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    constexpr uint32_t idesc = make_idesc_bf16(128, BN);
    uint64_t adesc = make_kmajor_desc<ROWB>(sa);
    adesc &= ~(0x3FFFull << 32);
    adesc |= (uint64_t)((sbo_rows * ROWB) >> 4) << 32;
    const uint64_t bdesc = make_kmajor_desc<ROWB>(sb_al);
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      for (int k = 0; k < ksteps; ++k) umma_bf16_ss(tmem, adesc + 2 * (k % (ROWB / 32)) + (k / (ROWB / 32)) * (ROWB >> 4), bdesc + 2 * (k % (ROWB / 32)), idesc, 1u);
    }
    umma_commit(bar);
    mbar_wait(bar, 0);
    const long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  tcgen05_fence_before(); __syncthreads();
  if (warp == 1) { tcgen05_fence_after(); tmem_dealloc<256>(tmem); }
}

template <int ROWB, int BN>
void run(int sbo, int grid, long long* d) {
  const int smem = 200 * ROWB + 1024 + 256 * ROWB + 64 + 2048;
  cudaFuncSetAttribute(rate_kernel<ROWB, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int iters = 500, ksteps = 18;
  rate_kernel<ROWB, BN><<<grid, 128, smem>>>(iters, sbo, ksteps, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long c = 0;
  cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
  printf("ROWB %3d N %3d sbo_rows %2d grid %3d: %7.1f cycles/MMA  (%s)  -> %.0f MAC/cycle/SM\n", ROWB, BN, sbo, grid,
         (double)c / (iters * ksteps), cudaGetErrorString(e), 128.0 * BN * 16 / ((double)c / (iters * ksteps)));
}

// Second experiment: the operand patterns of the im2col-free kernels.
//   mode 0: SW128 A, window starting `off` bytes into the row and running across the following rows (ws2conv: off = 64)
//   mode 1: non-swizzled A with overlapping 32-byte windows 16 bytes apart (conv1_pair / c12: LBO 16, SBO = row pitch)
__device__ __forceinline__ uint64_t desc_noswz(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((addr & 0x3FFFF) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
template <int BN>
__global__ void __launch_bounds__(128, 1) pattern_kernel(int iters, int mode, int off, int ksteps, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* bp = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t sa = base, sb = base + 64 * 1024, bar = sb + 32 * 1024;
  volatile uint32_t* slot = reinterpret_cast<volatile uint32_t*>(bp + (bar - base) + 16);
  for (int i = threadIdx.x; i < (int)((bar - base) / 4); i += blockDim.x) reinterpret_cast<uint32_t*>(bp)[i] = 0;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  if (warp == 1) tmem_alloc<256>(smem_u32(const_cast<uint32_t*>(slot)));
  fence_proxy_async_smem();
  tcgen05_fence_before(); __syncthreads(); tcgen05_fence_after();
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    constexpr uint32_t idesc = make_idesc_bf16(128, BN);
    uint64_t ad[24], bd[24];                       // descriptors built BEFORE the timed loop (fully unrolled -> registers)
#pragma unroll
    for (int k = 0; k < 24; ++k) {
      if (mode == 0) {
        ad[k] = make_kmajor_desc<128>(sa + off + 32 * (k % 8) + (k / 8) * 1280);
        ad[k] &= ~(0x3FFFull << 32);
        ad[k] |= (uint64_t)(1280 >> 4) << 32;
        bd[k] = make_kmajor_desc<128>(sb + ((k / 4) & 1) * BN * 128) + 2 * (k % 4);
      } else {
        ad[k] = desc_noswz(sa + (k % 3) * 192 + off, 16, 192);
        bd[k] = desc_noswz(sb + (k % 3) * 2048, 128, 256);
      }
    }
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int k = 0; k < 24; ++k) umma_bf16_ss(tmem, ad[k], bd[k], idesc, 1u);
    }
    umma_commit(bar);
    mbar_wait(bar, 0);
    const long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  tcgen05_fence_before(); __syncthreads();
  if (warp == 1) { tcgen05_fence_after(); tmem_dealloc<256>(tmem); }
}
template <int BN>
void run_pattern(int mode, int off, int grid, long long* d) {
  const int smem = 64 * 1024 + 32 * 1024 + 64 + 2048;
  cudaFuncSetAttribute(pattern_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int iters = 400, ksteps = 24;
  pattern_kernel<BN><<<grid, 128, smem>>>(iters, mode, off, ksteps, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long c = 0;
  cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
  printf("pattern %s N %3d off %3d grid %3d: %7.1f cycles/MMA  (%s)\n", mode == 0 ? "SW128 shifted window (ws2conv)" : "no-swizzle overlapping (conv1) ",
         BN, off, grid, (double)c / (iters * ksteps), cudaGetErrorString(e));
}

int main() {
  long long* d; cudaMalloc(&d, 8);
  for (int grid : {1, 148}) {
    for (int off : {0, 64}) { run_pattern<64>(0, off, grid, d); run_pattern<128>(0, off, grid, d); }
    for (int off : {0, 32}) run_pattern<64>(1, off, grid, d);
  }
  for (int grid : {1, 148}) for (int sbo : {8, 10}) {
    run<64, 32>(sbo, grid, d); run<64, 64>(sbo, grid, d); run<64, 128>(sbo, grid, d);
    run<128, 32>(sbo, grid, d); run<128, 64>(sbo, grid, d); run<128, 128>(sbo, grid, d); run<128, 256>(sbo, grid, d);
  }
  return 0;
}
