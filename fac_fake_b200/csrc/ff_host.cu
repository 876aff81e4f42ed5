// ff_host.cu — definitions of the helpers declared in ff_host.h.
#include "ff_host.h"

#include <cuda_fp16.h>

#include <mutex>
#include <set>

namespace ffh {

EncodeTiledFn encode_tiled() {
  static std::once_flag once;
  static EncodeTiledFn fn = nullptr;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && p)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

cudaError_t ensure_dyn_smem(const void* kernel, int bytes) {
  static std::mutex mu;
  static std::set<std::pair<int, const void*>> done;
  int dev = -1;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  std::lock_guard<std::mutex> lk(mu);
  if (done.count({dev, kernel})) return cudaSuccess;
  e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess) done.insert({dev, kernel});
  return e;
}

std::vector<bf16> to_bf16(const std::vector<float>& v) {
  std::vector<bf16> o(v.size());
  for (size_t i = 0; i < v.size(); ++i) o[i] = __float2bfloat16(v[i]);
  return o;
}

std::vector<bf16> to_f16_bits(const std::vector<float>& v) {
  std::vector<bf16> o(v.size());
  for (size_t i = 0; i < v.size(); ++i) {
    const __half hv = __float2half_rn(v[i]);
    memcpy(&o[i], &hv, sizeof(hv));
  }
  return o;
}

}  // namespace ffh
