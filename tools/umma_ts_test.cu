// Microbenchmark 2: cycles per tcgen05.mma (kind::f16, K = 16) when the A operand comes from TENSOR MEMORY instead of
// shared memory, for M = 64 / 128 (cta_group::1) and M = 128 / 256 (cta_group::2), as a function of N.
// Question it answers: does taking A out of the shared-memory port lift the 58-cycle floor of the N <= 64 instructions?
// Operands are zeros; only timing matters.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_ts_test umma_ts_test.cu
#include <cstdio>
#include <cooperative_groups.h>
#include "../fac_fake_b200/csrc/ff_ptx.cuh"
using namespace ff;
namespace cg = cooperative_groups;

__device__ __forceinline__ void umma_ts(uint32_t d, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_ts_2cta(uint32_t d, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}

// MODE 0: SS (A, B in smem)   1: TS (A in TMEM columns 256.., B in smem)
template <int MODE, int BM, int BN>
__global__ void __launch_bounds__(128, 1) rate1_kernel(int iters, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* bp = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t sa = base, sb = base + 32 * 1024, bar = sb + 64 * 1024;
  volatile uint32_t* slot = reinterpret_cast<volatile uint32_t*>(bp + (bar - base) + 16);
  for (int i = threadIdx.x; i < (int)((bar - base) / 4); i += blockDim.x) reinterpret_cast<uint32_t*>(bp)[i] = 0;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  if (warp == 1) tmem_alloc<512>(smem_u32(const_cast<uint32_t*>(slot)));
  fence_proxy_async_smem();
  tcgen05_fence_before(); __syncthreads(); tcgen05_fence_after();
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    constexpr uint32_t idesc = make_idesc_bf16(BM, BN);
    const uint64_t adesc = make_kmajor_desc<128>(sa);
    const uint64_t bdesc = make_kmajor_desc<128>(sb);
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        if (MODE == 0) umma_bf16_ss(tmem, adesc + 2 * (k % 4) + (k / 4) * 1024, bdesc + 2 * (k % 4) + ((k / 4) & 1) * (BN * 8), idesc, 1u);
        else umma_ts(tmem, tmem + 256 + 8 * k, bdesc + 2 * (k % 4) + ((k / 4) & 1) * (BN * 8), idesc, 1u);
      }
    }
    umma_commit(bar);
    mbar_wait(bar, 0);
    const long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  tcgen05_fence_before(); __syncthreads();
  if (warp == 1) { tcgen05_fence_after(); tmem_dealloc<512>(tmem); }
}

template <int MODE, int BM, int BN>
void run1(int grid, long long* d) {
  const int smem = 32 * 1024 + 64 * 1024 + 64 + 2048;
  cudaFuncSetAttribute(rate1_kernel<MODE, BM, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int iters = 400;
  rate1_kernel<MODE, BM, BN><<<grid, 128, smem>>>(iters, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long c = 0;
  cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
  const double cyc = (double)c / (iters * 16);
  printf("cta_group::1 %s M %3d N %3d grid %3d: %7.1f cycles/MMA (%s) -> %.0f MAC/cycle/SM\n", MODE ? "TS" : "SS", BM, BN, grid, cyc,
         cudaGetErrorString(e), (double)BM * BN * 16 / cyc);
}

// CTA pair: BM = 128 or 256 rows across the two CTAs; B rows split between the two CTAs' shared memories.
template <int MODE, int BM, int BN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) rate2_timed(int iters, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  cg::cluster_group cluster = cg::this_cluster();
  const uint32_t rank = cluster.block_rank();
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* bp = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t sa = base, sb = base + 32 * 1024, bar = sb + 64 * 1024;
  volatile uint32_t* slot = reinterpret_cast<volatile uint32_t*>(bp + (bar - base) + 16);
  for (int i = threadIdx.x; i < (int)((bar - base) / 4); i += blockDim.x) reinterpret_cast<uint32_t*>(bp)[i] = 0;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(const_cast<uint32_t*>(slot))) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  fence_proxy_async_smem();
  tcgen05_fence_before();
  cluster.sync();
  tcgen05_fence_after();
  const uint32_t tmem = *slot;
  long long t0 = 0;
  if (rank == 0 && threadIdx.x == 0) {
    constexpr uint32_t idesc = make_idesc_bf16(BM, BN);
    const uint64_t adesc = make_kmajor_desc<128>(sa);
    const uint64_t bdesc = make_kmajor_desc<128>(sb);
    t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        if (MODE == 0) umma_bf16_ss_2cta(tmem, adesc + 2 * (k % 4) + (k / 4) * 1024, bdesc + 2 * (k % 4) + ((k / 4) & 1) * (BN * 4), idesc, 1u);
        else umma_ts_2cta(tmem, tmem + 256 + 8 * k, bdesc + 2 * (k % 4) + ((k / 4) & 1) * (BN * 4), idesc, 1u);
      }
    }
    umma_commit_2cta(bar, 3);
  }
  if (threadIdx.x == 0) {
    mbar_wait(bar, 0);
    if (rank == 0 && blockIdx.x == 0) out[0] = clock64() - t0;
  }
  tcgen05_fence_before();
  cluster.sync();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
  }
}

template <int MODE, int BM, int BN>
void run2(int grid, long long* d) {
  const int smem = 32 * 1024 + 64 * 1024 + 64 + 2048;
  cudaFuncSetAttribute(rate2_timed<MODE, BM, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int iters = 400;
  rate2_timed<MODE, BM, BN><<<grid, 128, smem>>>(iters, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long c = 0;
  cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
  const double cyc = (double)c / (iters * 16);
  printf("cta_group::2 %s M %3d N %3d grid %3d: %7.1f cycles/MMA (%s) -> %.0f MAC/cycle/SM\n", MODE ? "TS" : "SS", BM, BN, grid, cyc,
         cudaGetErrorString(e), (double)BM * BN * 16 / cyc / 2);
}

// Third experiment: does a tcgen05.commit after every CE MMAs (the per-stage "empty" arrival of a pipelined mainloop)
// slow the tensor pipe down?  Commits go to mbarriers nobody waits on (count 1, phases just flip).
template <int BN, int CE>
__global__ void __launch_bounds__(128, 1) commit_kernel(int iters, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* bp = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t sa = base, sb = base + 32 * 1024, bar = sb + 64 * 1024;
  volatile uint32_t* slot = reinterpret_cast<volatile uint32_t*>(bp + (bar - base) + 64);
  for (int i = threadIdx.x; i < (int)((bar - base) / 4); i += blockDim.x) reinterpret_cast<uint32_t*>(bp)[i] = 0;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { for (int i = 0; i < 5; ++i) mbar_init(bar + 8 * i, 1); fence_mbar_init(); }
  if (warp == 1) tmem_alloc<512>(smem_u32(const_cast<uint32_t*>(slot)));
  fence_proxy_async_smem();
  tcgen05_fence_before(); __syncthreads(); tcgen05_fence_after();
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    constexpr uint32_t idesc = make_idesc_bf16(128, BN);
    const uint64_t adesc = make_kmajor_desc<128>(sa);
    const uint64_t bdesc = make_kmajor_desc<128>(sb);
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        umma_bf16_ss(tmem, adesc + 2 * (k % 4) + (k / 4) * 1024, bdesc + 2 * (k % 4) + ((k / 4) & 1) * (BN * 8), idesc, 1u);
        if (CE > 0 && (k % CE) == CE - 1) umma_commit(bar + 8 * (1 + (k / CE) % 4));
      }
    }
    umma_commit(bar);
    mbar_wait(bar, 0);
    const long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  tcgen05_fence_before(); __syncthreads();
  if (warp == 1) { tcgen05_fence_after(); tmem_dealloc<512>(tmem); }
}
template <int BN, int CE>
void run3(long long* d) {
  const int smem = 32 * 1024 + 64 * 1024 + 128 + 2048;
  cudaFuncSetAttribute(commit_kernel<BN, CE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int iters = 400;
  commit_kernel<BN, CE><<<1, 128, smem>>>(iters, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long c = 0;
  cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
  printf("SS M 128 N %3d, commit every %2d MMAs: %7.1f cycles/MMA (%s)\n", BN, CE, (double)c / (iters * 16), cudaGetErrorString(e));
}

int main() {
  long long* d; cudaMalloc(&d, 16);
  run3<64, 0>(d); run3<64, 4>(d); run3<64, 1>(d); run3<128, 4>(d); run3<192, 4>(d); run3<256, 4>(d);
  for (int grid : {1, 148}) {
    run1<0, 128, 32>(grid, d); run1<0, 128, 64>(grid, d); run1<0, 128, 128>(grid, d); run1<0, 128, 256>(grid, d);
    run1<1, 128, 32>(grid, d); run1<1, 128, 64>(grid, d); run1<1, 128, 128>(grid, d); run1<1, 128, 256>(grid, d);
    run1<0, 64, 64>(grid, d); run1<0, 64, 128>(grid, d); run1<0, 64, 256>(grid, d);
    run1<1, 64, 64>(grid, d); run1<1, 64, 128>(grid, d); run1<1, 64, 256>(grid, d);
  }
  for (int grid : {2, 148}) {
    run2<0, 256, 64>(grid, d); run2<0, 256, 128>(grid, d); run2<0, 256, 256>(grid, d);
    run2<1, 256, 64>(grid, d); run2<1, 256, 128>(grid, d); run2<1, 256, 256>(grid, d);
    run2<0, 128, 64>(grid, d); run2<0, 128, 128>(grid, d); run2<0, 128, 256>(grid, d);
    run2<1, 128, 64>(grid, d); run2<1, 128, 128>(grid, d); run2<1, 128, 256>(grid, d);
  }
  return 0;
}
