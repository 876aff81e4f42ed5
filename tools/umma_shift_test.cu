// Experiment: can a tcgen05 K-major swizzled smem descriptor start at an arbitrary ROW offset inside a TMA-written
// buffer (start address not aligned to the swizzle atom) and use a group stride (SBO) that is not a multiple of
// the atom?  This decides whether the 3x3 conv can read all 9 taps from ONE halo patch in shared memory.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/umma_shift_test tools/umma_shift_test.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../fac_fake_b200/csrc/ff_ptx.cuh"
using namespace ff;

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int ROWB>
__global__ void __launch_bounds__(128, 1)
shift_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int shift_rows, int sbo_rows,
             int base_mode, float* out) {
  constexpr int AROWS = 256, BN = 32;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* bp = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t sa = base, sb = base + AROWS * ROWB;
  const uint32_t bar = sb + BN * ROWB;          // 8-byte aligned
  const uint32_t bar2 = bar + 8;
  volatile uint32_t* slot = reinterpret_cast<volatile uint32_t*>(bp + AROWS * ROWB + BN * ROWB + 16);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(bar2, 1); fence_mbar_init(); }
  if (warp == 1) tmem_alloc<32>(smem_u32(const_cast<uint32_t*>(slot)));
  tcgen05_fence_before(); __syncthreads(); tcgen05_fence_after();
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(bar, AROWS * ROWB + BN * ROWB);
    tma_load_2d(sa, &tmA, bar, 0, 0);
    tma_load_2d(sb, &tmB, bar, 0, 0);
    mbar_wait(bar, 0);
    tcgen05_fence_after();
    const uint32_t a_start = sa + shift_rows * ROWB;
    uint64_t adesc = make_kmajor_desc<ROWB>(a_start);
    // override SBO
    adesc &= ~(0x3FFFull << 32);
    adesc |= (uint64_t)((sbo_rows * ROWB) >> 4) << 32;
    if (base_mode == 1) adesc |= (uint64_t)((a_start >> 7) & 7) << 49;
    const uint64_t bdesc = make_kmajor_desc<ROWB>(sb);
    constexpr uint32_t idesc = make_idesc_bf16(128, BN);
    for (int k = 0; k < ROWB / 32; ++k) umma_bf16_ss(tmem, adesc + 2 * k, bdesc + 2 * k, idesc, k > 0);
    umma_commit(bar2);
  }
  mbar_wait(bar2, 0);
  tcgen05_fence_after();
  uint32_t v[32];
  tmem_ld_32x32(tmem + ((uint32_t)(warp * 32) << 16), v);
  tmem_ld_wait();
  for (int i = 0; i < 32; ++i) out[(warp * 32 + lane) * 32 + i] = __uint_as_float(v[i]);
  tcgen05_fence_before(); __syncthreads();
  if (warp == 1) { tcgen05_fence_after(); tmem_dealloc<32>(tmem); }
}

int main() {
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  EncodeTiledFn enc = (EncodeTiledFn)fn;
  for (int rowb : {128, 64}) {
    const int KE = rowb / 2, AROWS = 256, BN = 32;
    std::vector<__nv_bfloat16> A(AROWS * KE), B(BN * KE);
    std::vector<float> Af(AROWS * KE), Bf(BN * KE);
    srand(1);
    for (int i = 0; i < AROWS * KE; ++i) { Af[i] = (float)(rand() % 7 - 3); A[i] = __float2bfloat16(Af[i]); }
    for (int i = 0; i < BN * KE; ++i) { Bf[i] = (float)(rand() % 5 - 2); B[i] = __float2bfloat16(Bf[i]); }
    __nv_bfloat16 *dA, *dB; float* dO;
    cudaMalloc(&dA, A.size() * 2); cudaMalloc(&dB, B.size() * 2); cudaMalloc(&dO, 128 * 32 * 4);
    cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice);
    CUtensorMap tmA, tmB;
    CUtensorMapSwizzle sw = rowb == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    cuuint64_t dimsA[2] = {(cuuint64_t)KE, (cuuint64_t)AROWS}, strA[1] = {(cuuint64_t)KE * 2};
    cuuint32_t boxA[2] = {(cuuint32_t)KE, (cuuint32_t)AROWS}, es[2] = {1, 1};
    cuuint64_t dimsB[2] = {(cuuint64_t)KE, (cuuint64_t)BN};
    cuuint32_t boxB[2] = {(cuuint32_t)KE, (cuuint32_t)BN};
    if (enc(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dA, dimsA, strA, boxA, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) ||
        enc(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dB, dimsB, strA, boxB, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)) { printf("encode failed\n"); return 1; }
    const int smem = AROWS * rowb + BN * rowb + 64 + 1024;
    if (rowb == 128) cudaFuncSetAttribute(shift_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    else cudaFuncSetAttribute(shift_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int shifts[] = {0, 1, 2, 3, 5, 8, 10, 11, 12, 21, 22};
    for (int sbo : {8, 10, 18}) for (int s : shifts) for (int mode : {0, 1}) {
      if (s + 15 * sbo + 8 > AROWS) continue;
      cudaMemset(dO, 0, 128 * 32 * 4);
      if (rowb == 128) shift_kernel<128><<<1, 128, smem>>>(tmA, tmB, s, sbo, mode, dO);
      else shift_kernel<64><<<1, 128, smem>>>(tmA, tmB, s, sbo, mode, dO);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("rowb %d sbo %d shift %d mode %d: CUDA error %s\n", rowb, sbo, s, mode, cudaGetErrorString(e)); return 2; }
      std::vector<float> O(128 * 32);
      cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost);
      double maxerr = 0; int bad = 0;
      for (int r = 0; r < 128; ++r) {
        const int src = s + (r / 8) * sbo + (r % 8);
        for (int n = 0; n < BN; ++n) {
          float ref = 0;
          for (int k = 0; k < KE; ++k) ref += Af[src * KE + k] * Bf[n * KE + k];
          const double d = fabs(ref - O[r * 32 + n]);
          if (d > maxerr) maxerr = d;
          if (d > 1e-3) ++bad;
        }
      }
      printf("rowb %3d sbo_rows %2d shift %2d base_mode %d : max_err %8.3f bad %4d %s\n", rowb, sbo, s, mode, maxerr, bad, bad ? "FAIL" : "ok");
    }
    cudaFree(dA); cudaFree(dB); cudaFree(dO);
  }
  return 0;
}
