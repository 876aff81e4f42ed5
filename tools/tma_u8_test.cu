// Isolated test of the uint8 3-D TMA box load used by conv1_tma_kernel.
#include <cstdio>
#include <vector>
#include "../fac_fake_b200/csrc/ff_ptx.cuh"
using namespace ff;
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__global__ void k(const __grid_constant__ CUtensorMap tm, int x, int y, int n, uint8_t* out) {
  __shared__ __align__(128) uint8_t raw[896];
  __shared__ __align__(8) uint64_t bar;
  const uint32_t b = smem_u32(&bar);
  if (threadIdx.x == 0) { mbar_init(b, 1); fence_mbar_init(); }
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(b, 864);
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(smem_u32(raw)), "l"(reinterpret_cast<uint64_t>(&tm)), "r"(b), "r"(x), "r"(y), "r"(n) : "memory");
  }
  mbar_wait(b, 0);
  for (int i = threadIdx.x; i < 864; i += blockDim.x) out[i] = raw[i];
}
int main() {
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  EncodeTiledFn enc = (EncodeTiledFn)fn;
  const int N = 3;
  std::vector<uint8_t> h((size_t)N * 224 * 672);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (uint8_t)((i * 7 + (i >> 9)) & 0xff);
  uint8_t *d, *o; cudaMalloc(&d, h.size()); cudaMalloc(&o, 864);
  cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
  for (int promo = 0; promo < 2; ++promo) {
    CUtensorMap tm;
    cuuint64_t dims[3] = {672, 224, N}, strides[2] = {672, 224 * 672};
    cuuint32_t box[3] = {48, 18, 1}, es[3] = {1, 1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, promo ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_NONE,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("promo %d encode rc %d\n", promo, (int)r);
    const int cases[4][3] = {{32, 15, 1}, {-16, -1, 0}, {640, 207, 2}, {656, 100, 1}};
    for (auto& c : cases) {
      cudaMemset(o, 0xEE, 864);
      k<<<1, 128>>>(tm, c[0], c[1], c[2], o);
      cudaError_t e = cudaDeviceSynchronize();
      std::vector<uint8_t> got(864);
      cudaMemcpy(got.data(), o, 864, cudaMemcpyDeviceToHost);
      int bad = 0;
      for (int r2 = 0; r2 < 18; ++r2) for (int b = 0; b < 48; ++b) {
        const int gy = c[1] + r2, gx = c[0] + b;
        uint8_t want = 0;
        if (gy >= 0 && gy < 224 && gx >= 0 && gx < 672) want = h[((size_t)c[2] * 224 + gy) * 672 + gx];
        if (got[r2 * 48 + b] != want) ++bad;
      }
      printf("  x %4d y %4d n %d: %s bad=%d\n", c[0], c[1], c[2], cudaGetErrorString(e), bad);
      if (e != cudaSuccess) return 1;
    }
  }
  return 0;
}
