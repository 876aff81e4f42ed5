// ff_host.h — host-side helpers shared by every translation unit of libfacfake.so (ff_cvit.cu, ff_resvitkan.cu,
// ff_ggca.cu, ff_s3d.cu, ff_blaze.cu): the driver entry point for tensor maps, a device guard, per-device
// kernel-attribute bookkeeping and the PDL launch helper.  Nothing here is specific to one engine.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstring>
#include <utility>
#include <vector>

namespace ffh {

typedef __nv_bfloat16 bf16;

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
// cuTensorMapEncodeTiled through the runtime's driver entry point (no -lcuda); nullptr if the driver lacks it.
EncodeTiledFn encode_tiled();

// Every C-ABI entry point runs on the handle's device and leaves the caller's current device as it found it
// (ctypes callers share the process with torch, whose current device must not change under it).
struct DeviceGuard {
  int prev = -1;
  bool switched = false;
  cudaError_t status = cudaSuccess;
  explicit DeviceGuard(int device) {
    status = cudaGetDevice(&prev);
    if (status == cudaSuccess && prev != device) {
      status = cudaSetDevice(device);
      switched = status == cudaSuccess;
    }
  }
  ~DeviceGuard() {
    if (switched) cudaSetDevice(prev);
  }
  DeviceGuard(const DeviceGuard&) = delete;
  DeviceGuard& operator=(const DeviceGuard&) = delete;
};

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-DEVICE attribute: remember (device, kernel) pairs, not kernels.
cudaError_t ensure_dyn_smem(const void* kernel, int bytes);

std::vector<bf16> to_bf16(const std::vector<float>& v);
// fp16 bit patterns carried in the same 16-bit element type (the tensor maps only move bytes; the MMA instruction
// descriptor says how to read them)
std::vector<bf16> to_f16_bits(const std::vector<float>& v);
inline int ilog2(int v) { int l = 0; while ((1 << l) < v) ++l; return l; }

// Launch with the programmatic-stream-serialization attribute (PDL): the kernel may start while its predecessor in
// the stream is still draining; every kernel launched this way executes griddepcontrol.wait before it touches global
// data produced (or still read) by the predecessor.
template <typename... KArgs, typename... Args>
cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl, Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}
// Same, after opting the kernel in to `smem` bytes of dynamic shared memory on the current device.
template <typename... KArgs, typename... Args>
cudaError_t launch_smem(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl, Args&&... args) {
  cudaError_t e = ensure_dyn_smem(reinterpret_cast<const void*>(kernel), static_cast<int>(smem));
  if (e != cudaSuccess) return e;
  return launch_k(kernel, grid, block, smem, st, pdl, std::forward<Args>(args)...);
}

}  // namespace ffh
