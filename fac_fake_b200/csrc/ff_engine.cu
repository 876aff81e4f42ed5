// ff_engine.cu — host side of libfacfake.so: weight folding/layout, workspace, TMA descriptors, the launch
// schedule of the CViT forward and the C-ABI declared in include/facfake.h.
//
// Reference path being replaced (all under /root/reference/CViT-main/):
//   model/cvit.py:80-179 (CViT), cvit_prediction.py:209-242 (model half of predict()), :258-281 (reduction).
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <utility>
#include <vector>

#include "../../include/facfake.h"
#include "ff_fp32.cuh"
#include "ff_pre.cuh"
#include "ff_small.cuh"
#include "ff_tc.cuh"
#include "ff_ws.cuh"
#include "ff_c1.cuh"
#include "ff_rvk.cuh"
#include "ff_c12.cuh"
#include "ff_xf.cuh"
#include "ff_ptc2.cuh"

namespace {

using namespace ff;
typedef __nv_bfloat16 bf16;

// ------------------------------------------------------------------------------------------------ model plan
struct ConvPlan { int cin, cout, hw; bool pool; int conv_idx; };
const ConvPlan kConv[17] = {
    {3, 32, 224, false, 0},    {32, 32, 224, false, 3},   {32, 32, 224, true, 6},
    {32, 64, 112, false, 10},  {64, 64, 112, false, 13},  {64, 64, 112, true, 16},
    {64, 128, 56, false, 20},  {128, 128, 56, false, 23}, {128, 128, 56, true, 26},
    {128, 256, 28, false, 30}, {256, 256, 28, false, 33}, {256, 256, 28, false, 36}, {256, 256, 28, true, 39},
    {256, 512, 14, false, 43}, {512, 512, 14, false, 46}, {512, 512, 14, false, 49}, {512, 512, 14, true, 52},
};
constexpr int DIM = 1024, DEPTH = 6, MLP = 2048, PATCH = 25088, SLOTS = 32;
constexpr float BN_EPS = 1e-5f;
constexpr int EMBED_SPLITS = 8;   // split-K of the 25088-deep patch embedding (392 k-blocks = 8 x 49)

std::string g_create_error;

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;

struct ConvLayerDev {
  bf16* w = nullptr;        // [cout][3][3][cin] bf16
  float* wf = nullptr;      // fp32 path: [cout][3][3][cin]
  float* scale = nullptr;
  float* shift = nullptr;
  CUtensorMap tmA, tmB;
  int rowb = 128, bn = 128;
  int bw = 16, bh = 8, bi = 1;
  bool ptc2 = false;        // CTA-pair kernel (ff_ptc2.cuh): half filter tile per CTA
  bool ptc2m = false;       // unvalidated Cout = 128 pair kernel (tmB_half then has 64-row boxes)
  CUtensorMap tmB_half;     // box {64, 128}
  bool ws = false;          // persistent weight-stationary halo kernel (ff_ws.cuh)
  CUtensorMap tmA_ws, tmW_ws;
  WsEpi epi;                // host copy of (scale, shift) passed by value to the ws kernel
  bool ws2x = false;        // pixel-pair formulation on a CTA pair (Cin = 64): N = 128, cta_group::2
  bf16* w2x = nullptr;      // pair-expanded filter [128][768]
  CUtensorMap tmA_ws2x, tmW_ws2x;
  bool ws4 = false;         // pixel-quad formulation (32 -> 32): N = 4*Cout = 128, compact expanded filter
  bf16* w4 = nullptr;       // [3 kh][64 + 128 + 64 rows][64] (ff_ws.cuh: Ws4Smem)
  CUtensorMap tmA_ws4, tmW_ws4;
  bool ws2 = false;         // pixel-pair formulation (Cin = 32): N = 2*Cout
  bf16* w2 = nullptr;       // pair-expanded filter [2*Cout][384]
  CUtensorMap tmA_ws2, tmW_ws2;
};
struct LinearDev {
  bf16* w = nullptr;        // [out][in] bf16
  float* wf = nullptr;      // fp32 copy (fp32 path / head2)
  float* b = nullptr;
  int out_f = 0, in_f = 0;
  CUtensorMap tmB;          // box {64, bn}
  int bn = 128;
};
struct XfLayerDev {
  float *ln1_g, *ln1_b, *ln2_g, *ln2_b;
  LinearDev qkv, out, ff1, ff2;
};

}  // namespace

struct ff_cvit {
  int device = 0;
  int cap = 0;             // crops per pass (multiple of 32)
  int rows_cap = 0;        // token rows capacity (multiple of 128)
  int s12 = 64;            // crops per stage-1/2 sub-pass
  int s12_cap = 256;
  int compute = FF_COMPUTE_BF16;
  int variant = 0;         // tile-shape variant (tuning)
  int use_ws = 1;          // feature layers 2..6 on the weight-stationary halo kernel
  int ws_ctas_per_sm = 2;   // CTAs per SM for the Cin=32 weight-stationary kernels (Cin=64 always 1: smem)
  int use_ptc = 1;         // feature layers 7..17 on the persistent implicit-GEMM kernel
  int use_ptc2 = 1;        // layers 10..17 on the CTA-pair kernel (ff_ptc2.cuh); FF_PTC2=0: single-CTA persistent kernel
  int use_ptc2m = 0;       // NOT VALIDATED ON HARDWARE (FF_PTC2_128=1): layers 7..9 on the Cout = 128 pair kernel
  int use_c12 = 1;         // feature layers 1+2 fused in one kernel (ff_c12.cuh) on the uint8 path; FF_C12=0 -> separate kernels
  int use_ws4 = 0;         // 32 -> 32 layers (2, 3) in the pixel-quad formulation (FF_WS4=1; measured equal to the pair kernel)
  int use_ws2 = 1;         // Cin = 32 layers in the pixel-pair formulation
  int use_ws2x = 1;        // Cin = 64 layers (5, 6) in the pixel-pair formulation on CTA pairs (needs use_ws2)
  int use_c1_tc = 3;       // feature layer 1: 3 = pixel-pair GEMM out of the patch (no im2col), 2 = TMA-fed im2col rows,
                           // 1 = register-prefetched im2col rows, 0 = CUDA cores
  int c1_ctas_per_sm = 8;
  bf16* c1_w = nullptr;    // [32][64] bf16, k = kh*16 + kw*4 + cin
  bf16* c1_lut = nullptr;  // [3][256] bf16 normalisation table
  bf16* c1_wp = nullptr;   // [3][64][16] pair-expanded conv1 filter (conv1_pair_kernel)
  float c1_na[3] = {0, 0, 0}, c1_nb[3] = {0, 0, 0};   // FMA form of the table (valid iff it reproduces all 768 entries)
  int num_sms = 148;
  bool finalized = false;
  std::mutex mu;
  mutable std::string err;
  int64_t launches = 0;
  cudaEvent_t done_ev = nullptr;

  std::map<std::string, std::vector<float>> host_w;
  std::map<std::string, std::vector<int64_t>> host_shape;

  Conv1Params conv1;
  ConvLayerDev conv[17];
  ConvLayerDev conv_alt[6];   // layers 1..6 with TMA descriptors over the second ping-pong set
  LinearDev embed, head1, head2;
  XfLayerDev xf[DEPTH];
  float *pos = nullptr, *cls = nullptr;

  // workspace
  bf16 *bufA = nullptr, *bufB = nullptr;   // stage 1/2 ping-pong, s12_cap crops
  bf16 *bufA2 = nullptr, *bufB2 = nullptr; // second ping-pong set: odd sub-passes run on aux_stream (dual-stream overlap)
  cudaStream_t aux_stream = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  // ---- GGCA / DEConv / RepBN variant (kind == 2): the CViT plan + one BN-less linear conv + the gate
  bf16* bufR = nullptr;                  // output of the extra Conv2d(128,128) (features1.26), read by layer 9
  float *ggca_w1 = nullptr, *ggca_b1 = nullptr, *ggca_w2 = nullptr, *ggca_b2 = nullptr;
  float ln2_eps = 1e-5f;                 // eps of the MLP-branch LayerNorm (1e-6 for LinearNorm.norm1)
  // ---- ResVitKan (kind == 1): ResNet-50 features + the same ViT + KAN head
  int kind = 0;
  struct RvkOp {
    int type = 0;            // 0 = 1x1 stride-1 conv as GEMM, 1 = implicit-GEMM conv (3x3 any stride, or 1x1 stride 2)
    int cin = 0, cout = 0, taps = 9, stride = 1, in_hw = 0, out_hw = 0;
    int act = 0;             // 1 = ReLU after BN
    int resid = -1;          // buffer index of the residual (conv3) or -1
    int in_buf = 0, out_buf = 0;
    int bn = 64, bw = 8, bh = 8, bi = 2;
    bf16* w = nullptr;
    bf16* out_ptr = nullptr;   // explicit output (kind 2); otherwise rvk_buf[out_buf] / feat
    float *scale = nullptr, *shift = nullptr;
    CUtensorMap tmA, tmB;
    CUtensorMap tmO, tmR;    // TMA-store epilogue: output / residual tiles in the output geometry
    CUtensorMap tmA_flat;    // 1x1 stride-1 ops for the persistent kernel: [1][1][pixels][cin], boxes of 128 pixels
    std::string name;
  };
  std::vector<RvkOp> rvk_ops;            // kind 2 keeps its single extra conv (features1.26) here
  int rvk_persist = 2;                   // 2 = rvk_conv2_kernel (persistent, TMA epilogue), 1 = rvk_conv_kernel, 0 = tc_kernel per tile
  int rvk_layer_end[4] = {0, 0, 0, 0};   // index of the last op of layer1..4 (debug taps)
  bf16* rvk_buf[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  bf16* rvk_x4 = nullptr;                // normalised bf16 NHWC4 input
  bf16* rvk_stem_w = nullptr;
  float rvk_stem_scale[64], rvk_stem_shift[64];
  CUtensorMap rvk_tm_x4;
  float *kan_w0 = nullptr, *kan_w1 = nullptr, *kan_g0 = nullptr, *kan_g1 = nullptr;
  float* kan_part = nullptr;             // layer-0 split-K slabs [KAN_CHUNKS][cap][64]
  int gemm_bn_wide = 64;   // N tile of the wide transformer linears (qkv, ff1): 64 or 128
  int use_dual = 1;        // overlap consecutive stage-1/2 sub-passes on two streams (hides launch tails/prologues)
  bf16 *P = nullptr, *Q = nullptr;         // stage 3..5 ping-pong, cap crops
  bf16* feat = nullptr;                    // [cap_rows128][25088]
  float* emb = nullptr;                    // [cap][1024]
  float* x = nullptr;                      // [rows_cap][1024] residual stream
  bf16* xn = nullptr;                      // [rows_cap][1024]
  float* qkv = nullptr;                    // [rows_cap][3072] (fp32 path)
  bf16* qkvb = nullptr;                    // [rows_cap][3072] (bf16 path)
  bf16* att = nullptr;                     // [rows_cap][1024]
  bf16* ffh = nullptr;                     // [rows_cap][2048]
  bf16* clsb = nullptr;                    // [cap128][1024]
  float* hid = nullptr;                    // [cap128][2048]
  CUtensorMap tm_feat, tm_xn, tm_att, tm_ffh, tm_cls;
  // whole-encoder cooperative kernel (ff_xf.cuh): device copy of the tensor maps it indexes, readiness, opt-out (FF_XF=0)
  CUtensorMap* xf_maps = nullptr;
  unsigned int* xf_sync = nullptr;   // group-barrier counters
  int xf_groups = 0;                 // co-resident groups of 16 CTAs
  int use_xf = 1;
  bool xf_ready = false;
  // fp32-path workspace
  float *fA = nullptr, *fB = nullptr;
  // grow-only scratch for predict()
  int32_t* slot_buf = nullptr; size_t slot_cap = 0;
  uint8_t* xin_buf = nullptr; size_t xin_cap = 0;
  float* logit_buf = nullptr; size_t logit_cap = 0;
  int32_t* off_buf = nullptr; size_t off_cap = 0;
  float* score_buf = nullptr; size_t score_cap = 0;
  CropDesc* crop_desc = nullptr; size_t crop_desc_cap = 0;
  std::vector<void*> allocs;
  // pipelined host->device input copy (ff_cvit_predict_host): chunk j of `h2d_chunk` crops is copied on
  // copy_stream and published with h2d_ready[j]; the forward waits on it right before it first reads the chunk
  cudaStream_t copy_stream = nullptr;
  std::vector<cudaEvent_t> h2d_ready;
  int h2d_chunks_pending = 0;   // > 0 while a predict_host call is in flight
  int h2d_chunk = 0;
  // optional per-launch timing (bench.py roofline): event pairs tagged with a kernel class
  bool profiling = false;
  bool prof_coarse = false;   // true: only 3 phase boundaries per pass are timed (slots 0/1/2), launches stay PDL-chained
  std::vector<cudaEvent_t> ev_pool;
  size_t ev_used = 0;
  std::vector<int> ev_class;       // class of pair i (events 2i, 2i+1)
  double prof_ms[21] = {0};
  int64_t prof_launches[21] = {0};
};

namespace {

int fail(const ff_cvit* h, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (h) h->err = buf; else g_create_error = buf;
  return code;
}

// profile slots: 0 = conv1, 1..16 = tcgen05 conv layer (index li), 17 = embed GEMM, 18 = transformer GEMMs,
// 19 = head GEMM, 20 = small kernels
enum { KC_CONV1 = 0, KC_TC_CONV = 1, KC_GEMM_EMBED = 17, KC_GEMM_XF = 18, KC_GEMM_HEAD = 19, KC_SMALL = 20, KC_COUNT = 21 };

void prof_mark(ff_cvit* h, cudaStream_t st, int cls, bool begin, bool coarse = false) {
  if (!h->profiling || (h->prof_coarse != coarse)) return;
  if (h->ev_used >= h->ev_pool.size()) {
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) return;
    h->ev_pool.push_back(e);
  }
  cudaEventRecord(h->ev_pool[h->ev_used++], st);
  if (begin) h->ev_class.push_back(cls);
}
struct ProfScope {
  ff_cvit* h; cudaStream_t st; int cls;
  ProfScope(ff_cvit* h_, cudaStream_t st_, int cls_) : h(h_), st(st_), cls(cls_) { prof_mark(h, st, cls, true); }
  ~ProfScope() { prof_mark(h, st, cls, false); }
};

#define FF_CUDA(h, call)                                                                             \
  do {                                                                                               \
    cudaError_t e_ = (call);                                                                         \
    if (e_ != cudaSuccess) return fail(h, FF_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

#define FF_LAUNCH_CHECK(h, what)                                                                     \
  do {                                                                                               \
    cudaError_t e_ = cudaGetLastError();                                                             \
    if (e_ != cudaSuccess) return fail(h, FF_ERR_CUDA, "launch of %s failed: %s", what, cudaGetErrorString(e_)); \
    ++(h)->launches;                                                                                 \
  } while (0)

template <typename T>
int dev_alloc(ff_cvit* h, T** p, size_t count) {
  void* q = nullptr;
  FF_CUDA(h, cudaMalloc(&q, std::max<size_t>(count, 1) * sizeof(T)));
  h->allocs.push_back(q);
  *p = reinterpret_cast<T*>(q);
  return FF_OK;
}
template <typename T>
int dev_upload(ff_cvit* h, T** p, const std::vector<T>& v) {
  int rc = dev_alloc(h, p, v.size());
  if (rc) return rc;
  FF_CUDA(h, cudaMemcpy(*p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  return FF_OK;
}
template <typename T>
int grow(ff_cvit* h, T** p, size_t* cap, size_t need) {
  if (need <= *cap) return FF_OK;
  if (*p) cudaFree(*p);
  *p = nullptr;
  *cap = 0;
  void* q = nullptr;
  FF_CUDA(h, cudaMalloc(&q, need * sizeof(T)));
  *p = reinterpret_cast<T*>(q);
  *cap = need;
  return FF_OK;
}

std::vector<bf16> to_bf16(const std::vector<float>& v) {
  std::vector<bf16> o(v.size());
  for (size_t i = 0; i < v.size(); ++i) o[i] = __float2bfloat16(v[i]);
  return o;
}

// ------------------------------------------------------------------------------------------------ TMA descriptors
int tmap_2d(ff_cvit* h, CUtensorMap* m, const void* base, uint64_t inner, uint64_t rows, uint32_t box_inner,
            uint32_t box_rows) {
  cuuint64_t dims[2] = {inner, rows};
  cuuint64_t strides[1] = {inner * 2};
  cuuint32_t box[2] = {box_inner, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMapSwizzle sw = box_inner * 2 == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(h, FF_ERR_CUDA, "cuTensorMapEncodeTiled(2d %llu x %llu) failed: %d",
                                     (unsigned long long)inner, (unsigned long long)rows, (int)r);
  return FF_OK;
}
int tmap_4d(ff_cvit* h, CUtensorMap* m, const void* base, int C, int W, int H, int N, int boxC, int bw, int bh, int bi,
            int estride = 1) {
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  // with element strides the box is given in input-space extents: ceil(box/stride) elements are loaded per dim
  cuuint32_t box[4] = {(cuuint32_t)boxC, (cuuint32_t)(bw * estride), (cuuint32_t)(bh * estride), (cuuint32_t)bi};
  cuuint32_t estr[4] = {1, (cuuint32_t)estride, (cuuint32_t)estride, 1};
  CUtensorMapSwizzle sw = boxC * 2 == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(h, FF_ERR_CUDA, "cuTensorMapEncodeTiled(4d C%d W%d H%d N%d) failed: %d", C, W, H, N, (int)r);
  return FF_OK;
}

// ------------------------------------------------------------------------------------------------ launch helper
// Launch with the programmatic-stream-serialization attribute (PDL): the kernel may start while its predecessor
// in the stream is still draining; every kernel launched this way executes griddepcontrol.wait before it touches
// global data produced (or still read) by the predecessor.
bool g_use_pdl = true;
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
  cfg.numAttrs = (pdl && g_use_pdl) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

// ------------------------------------------------------------------------------------------------ tc launch dispatch
template <int MODE, int ROWB, int BN, bool POOL, int STAGES>
cudaError_t launch_tc_t(dim3 grid, cudaStream_t st, const CUtensorMap& a, const CUtensorMap& b, const TcArgs& args) {
  using L = TcSmem<ROWB, BN, STAGES>;
  auto k = tc_kernel<MODE, ROWB, BN, POOL, STAGES>;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL);
    if (e != cudaSuccess) return e;
    attr_done = true;
  }
  return launch_k(k, grid, dim3(192), L::TOTAL, st, true, a, b, args);
}

cudaError_t launch_conv(int rowb, int bn, bool pool, int variant, dim3 grid, cudaStream_t st, const CUtensorMap& a,
                        const CUtensorMap& b, const TcArgs& args) {
#define FF_CONV_CASE(R, N, S)                                                              \
  if (rowb == R && bn == N)                                                                \
    return pool ? launch_tc_t<MODE_CONV, R, N, true, S>(grid, st, a, b, args)              \
                : launch_tc_t<MODE_CONV, R, N, false, S>(grid, st, a, b, args);
  FF_CONV_CASE(64, 32, 4)
  FF_CONV_CASE(64, 64, 4)
  FF_CONV_CASE(128, 64, 4)
  if (variant == 2) { FF_CONV_CASE(128, 256, 2) }
  FF_CONV_CASE(128, 128, 3)
  FF_CONV_CASE(128, 256, 4)
#undef FF_CONV_CASE
  return cudaErrorInvalidValue;
}

template <int ROWB, int BN, bool POOL, int STAGES>
cudaError_t launch_ws_t(int grid, cudaStream_t st, const CUtensorMap& a, const CUtensorMap& w, const TcArgs& args,
                        const WsEpi& epi) {
  using L = WsSmem<ROWB, BN, STAGES>;
  auto k = wsconv_kernel<ROWB, BN, POOL, STAGES>;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL);
    if (e != cudaSuccess) return e;
    attr_done = true;
  }
  return launch_k(k, dim3(grid), dim3(192), L::TOTAL, st, true, a, w, args, epi);
}

cudaError_t launch_ws(int cin, int cout, bool pool, int cps, int grid, cudaStream_t st, const CUtensorMap& a,
                      const CUtensorMap& w, const TcArgs& args, const WsEpi& epi) {
  if (cps >= 2 && cin == 32) {   // shallower patch ring so that 2-3 CTAs share an SM
    if (cout == 32)
      return pool ? launch_ws_t<64, 32, true, 3>(grid, st, a, w, args, epi) : launch_ws_t<64, 32, false, 3>(grid, st, a, w, args, epi);
    return launch_ws_t<64, 64, false, 3>(grid, st, a, w, args, epi);
  }
  if (cin == 32 && cout == 32)
    return pool ? launch_ws_t<64, 32, true, 8>(grid, st, a, w, args, epi) : launch_ws_t<64, 32, false, 8>(grid, st, a, w, args, epi);
  if (cin == 32 && cout == 64) return launch_ws_t<64, 64, false, 8>(grid, st, a, w, args, epi);
  if (cin == 64 && cout == 64)
    return pool ? launch_ws_t<128, 64, true, 5>(grid, st, a, w, args, epi) : launch_ws_t<128, 64, false, 5>(grid, st, a, w, args, epi);
  return cudaErrorInvalidValue;
}

template <int BN, bool POOL, int STAGES>
cudaError_t launch_ws2_t(int grid, cudaStream_t st, const CUtensorMap& a, const CUtensorMap& w, const TcArgs& args,
                         const WsEpi& epi) {
  using L = Ws2Smem<BN, STAGES>;
  auto k = ws2conv_kernel<BN, POOL, STAGES>;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL);
    if (e != cudaSuccess) return e;
    attr_done = true;
  }
  return launch_k(k, dim3(grid), dim3(192), L::TOTAL, st, true, a, w, args, epi);
}

template <bool POOL>
cudaError_t launch_ws4_t(int grid, cudaStream_t st, const CUtensorMap& a, const CUtensorMap& w, const TcArgs& args, const WsEpi& epi) {
  auto k = ws4conv_kernel<POOL>;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, Ws4Smem::TOTAL);
    if (e != cudaSuccess) return e;
    attr_done = true;
  }
  return launch_k(k, dim3(grid), dim3(192), Ws4Smem::TOTAL, st, true, a, w, args, epi);
}

template <int BN, int MSUB, bool POOL, int STAGES>
cudaError_t launch_ptc_t(int grid, cudaStream_t st, const CUtensorMap& a, const CUtensorMap& b, const TcArgs& args) {
  using L = PtcSmem<BN, MSUB, STAGES>;
  auto k = ptc_conv_kernel<BN, MSUB, POOL, STAGES>;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL);
    if (e != cudaSuccess) return e;
    attr_done = true;
  }
  return launch_k(k, dim3(grid), dim3(192), L::TOTAL, st, true, a, b, args);
}

template <bool POOL>
cudaError_t launch_ptc2_t(int grid, cudaStream_t st, const CUtensorMap& a, const CUtensorMap& b, const TcArgs& args) {
  auto k = ptc2_conv_kernel<POOL>;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, Ptc2Smem::TOTAL);
    if (e != cudaSuccess) return e;
    attr_done = true;
  }
  return launch_k(k, dim3(grid), dim3(192), Ptc2Smem::TOTAL, st, true, a, b, args);
}

template <bool POOL>
cudaError_t launch_ptc2m_t(int grid, cudaStream_t st, const CUtensorMap& a, const CUtensorMap& b, const TcArgs& args) {
  auto k = ptc2m_conv_kernel<POOL>;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, Ptc2mSmem::TOTAL);
    if (e != cudaSuccess) return e;
    attr_done = true;
  }
  return launch_k(k, dim3(grid), dim3(192), Ptc2mSmem::TOTAL, st, true, a, b, args);
}

template <bool POOL>
cudaError_t launch_ws2x_t(int grid, cudaStream_t st, const CUtensorMap& a, const CUtensorMap& w, const TcArgs& args,
                          const WsEpi& epi) {
  auto k = ws2x_conv_kernel<POOL>;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, Ws2xSmem::TOTAL);
    if (e != cudaSuccess) return e;
    attr_done = true;
  }
  return launch_k(k, dim3(grid), dim3(192), Ws2xSmem::TOTAL, st, true, a, w, args, epi);
}

int conv_bn_for(int cout, int variant) {
  const int bn_max = (variant == 1) ? 128 : 256;
  return std::min(cout, bn_max);
}

// ------------------------------------------------------------------------------------------------ weights
const std::vector<float>* get_w(ff_cvit* h, const std::string& key, std::initializer_list<int64_t> shape) {
  auto it = h->host_w.find(key);
  if (it == h->host_w.end()) {
    fail(h, FF_ERR_STATE, "missing weight '%s'", key.c_str());
    return nullptr;
  }
  const auto& s = h->host_shape[key];
  if (s.size() != shape.size() || !std::equal(s.begin(), s.end(), shape.begin())) {
    fail(h, FF_ERR_SHAPE, "weight '%s' has the wrong shape", key.c_str());
    return nullptr;
  }
  return &it->second;
}

int xf_setup(ff_cvit* h);   // encoder cooperative kernel (defined next to its launch)

int upload_linear(ff_cvit* h, LinearDev* L, const std::string& name, int out_f, int in_f, bool bias, int bn) {
  const auto* w = get_w(h, name + ".weight", {out_f, in_f});
  if (!w) return h->err.find("missing") != std::string::npos ? FF_ERR_STATE : FF_ERR_SHAPE;
  L->out_f = out_f;
  L->in_f = in_f;
  L->bn = bn;
  int rc;
  if (out_f >= 32) {
    if ((rc = dev_upload(h, &L->w, to_bf16(*w)))) return rc;
    if ((rc = tmap_2d(h, &L->tmB, L->w, in_f, out_f, 64, bn))) return rc;
  }
  if (h->compute == FF_COMPUTE_FP32 || out_f < 32)
    if ((rc = dev_upload(h, &L->wf, *w))) return rc;
  if (bias) {
    const auto* b = get_w(h, name + ".bias", {out_f});
    if (!b) return FF_ERR_STATE;
    if ((rc = dev_upload(h, &L->b, *b))) return rc;
  }
  return FF_OK;
}

int upload_vec(ff_cvit* h, float** p, const std::string& key, int64_t n) {
  const auto* v = get_w(h, key, {n});
  if (!v) return FF_ERR_STATE;
  return dev_upload(h, p, *v);
}

void conv_tile_geometry(int hw, int* bw, int* bh, int* bi) {
  if (hw >= 112) { *bw = 16; *bh = 8; *bi = 1; }
  else if (hw == 56) { *bw = 8; *bh = 8; *bi = 2; }
  else if (hw == 28) { *bw = 4; *bh = 4; *bi = 8; }
  else { *bw = 2; *bh = 2; *bi = 32; }
}
int ilog2(int v) { int l = 0; while ((1 << l) < v) ++l; return l; }

// Which buffer conv layer li (1..16; layer 0 is conv1) reads, per the ping-pong schedule in forward_pass().
const bf16* conv_input_buffer(const ff_cvit* h, int li, int set = 0) {
  const bf16* A = set ? h->bufA2 : h->bufA;
  const bf16* B = set ? h->bufB2 : h->bufB;
  switch (li) {
    case 1: return A; case 2: return B; case 3: return A; case 4: return B; case 5: return A;
    case 6: return h->P; case 7: return h->Q; case 8: return h->kind == 2 ? h->bufR : h->P;
    case 9: return h->Q; case 10: return h->P; case 11: return h->Q; case 12: return h->P;
    case 13: return h->Q; case 14: return h->P; case 15: return h->Q; case 16: return h->P;
  }
  return nullptr;
}
bf16* conv_output_buffer(const ff_cvit* h, int li, int set = 0) {
  bf16* A = set ? h->bufA2 : h->bufA;
  bf16* B = set ? h->bufB2 : h->bufB;
  switch (li) {
    case 0: return A; case 1: return B; case 2: return A; case 3: return B; case 4: return A;
    case 5: return h->P;
    case 6: return h->Q; case 7: return h->P; case 8: return h->Q;
    case 9: return h->P; case 10: return h->Q; case 11: return h->P; case 12: return h->Q;
    case 13: return h->P; case 14: return h->Q; case 15: return h->P; case 16: return h->feat;
  }
  return nullptr;
}

int build_conv_maps(ff_cvit* h) {
  for (int pass = 0; pass < 2; ++pass)
  for (int li = 1; li < (pass == 0 ? 17 : 6); ++li) {
    const ConvPlan& p = kConv[li];
    if (pass == 1) h->conv_alt[li] = h->conv[li];
    ConvLayerDev& L = pass == 0 ? h->conv[li] : h->conv_alt[li];
    const int set = pass;
    L.rowb = p.cin == 32 ? 64 : 128;
    L.bn = conv_bn_for(p.cout, h->variant);
    conv_tile_geometry(p.hw, &L.bw, &L.bh, &L.bi);
    const int ncap = li <= 5 ? h->s12_cap : h->cap;
    int rc = tmap_4d(h, &L.tmA, conv_input_buffer(h, li, set), p.cin, p.hw, p.hw, ncap, L.rowb / 2, L.bw, L.bh, L.bi);
    if (rc) return rc;
    rc = tmap_2d(h, &L.tmB, L.w, (uint64_t)9 * p.cin, p.cout, L.rowb / 2, L.bn);
    if (rc) return rc;
    L.ptc2 = h->use_ptc2 && L.bn == 256 && L.rowb == 128;
    if (L.ptc2 && (rc = tmap_2d(h, &L.tmB_half, L.w, (uint64_t)9 * p.cin, p.cout, 64, 128))) return rc;
    L.ptc2m = h->use_ptc2m && L.bn == 128 && L.rowb == 128 && p.cout == 128 && li >= 6;
    if (L.ptc2m && (rc = tmap_2d(h, &L.tmB_half, L.w, (uint64_t)9 * p.cin, p.cout, 64, 64))) return rc;
    L.ws2x = h->use_ws && h->use_ws2 && h->use_ws2x && p.cin == 64 && p.cout == 64 && li <= 5;
    if (L.ws2x) {
      rc = tmap_4d(h, &L.tmA_ws2x, conv_input_buffer(h, li, set), 64, p.hw / 2, p.hw, 2 * ncap, 64, 10, 18, 1);
      if (rc) return rc;
      rc = tmap_2d(h, &L.tmW_ws2x, L.w2x, 768, 128, 64, 64);
      if (rc) return rc;
    }
    L.ws4 = h->use_ws && h->use_ws2 && h->use_ws4 && p.cin == 32 && p.cout == 32;
    if (L.ws4) {
      // even / odd pair planes: elementStrides = 2 on the pair axis, 18 -> 9 pairs per plane row
      cuuint64_t dims[4] = {64, (cuuint64_t)(p.hw / 2), (cuuint64_t)p.hw, (cuuint64_t)ncap};
      cuuint64_t strides[3] = {128, (cuuint64_t)(p.hw / 2) * 128, (cuuint64_t)p.hw * (p.hw / 2) * 128};
      cuuint32_t box[4] = {64, 18, 18, 1};
      cuuint32_t estr[4] = {1, 2, 1, 1};
      CUresult r = g_encode(&L.tmA_ws4, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<bf16*>(conv_input_buffer(h, li, set)), dims, strides,
                            box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) return fail(h, FF_ERR_CUDA, "cuTensorMapEncodeTiled(quad planes, layer %d) failed: %d", li + 1, (int)r);
      rc = tmap_2d(h, &L.tmW_ws4, L.w4, 64, 768, 64, 64);
      if (rc) return rc;
    }
    L.ws2 = h->use_ws && h->use_ws2 && p.cin == 32;
    if (L.ws2) {
      rc = tmap_4d(h, &L.tmA_ws2, conv_input_buffer(h, li, set), 64, p.hw / 2, p.hw, ncap, 64, 10, 18, 1);
      if (rc) return rc;
      rc = tmap_2d(h, &L.tmW_ws2, L.w2, 384, (uint64_t)2 * p.cout, 64, 2 * p.cout);
      if (rc) return rc;
    }
    L.ws = h->use_ws && li <= 5;
    if (L.ws) {
      rc = tmap_4d(h, &L.tmA_ws, conv_input_buffer(h, li, set), p.cin, p.hw, p.hw, ncap, p.cin, 10, 18, 1);
      if (rc) return rc;
      rc = tmap_2d(h, &L.tmW_ws, L.w, (uint64_t)9 * p.cin, p.cout, p.cin, p.cout);
      if (rc) return rc;
    }
  }
  return FF_OK;
}


// ------------------------------------------------------------------------------------------------ GGCA variant weights
// cvit_GGCA_ADD_DEConv_RepBn8.py:361-423: (sequential, conv index, is DEConv, BN index or -1).  Entry 8 is the extra
// Conv2d(128,128) without BN / activation (features1.26); entry 9 the BN-less DEConv(128) + ReLU + pool (features1.27).
struct GgcaPlan { const char* seq; int conv_idx; bool de; int bn_idx; };
const GgcaPlan kGgcaPlan[18] = {
    {"features1", 0, false, 1},  {"features1", 3, true, 4},   {"features1", 6, true, 7},   {"features1", 10, false, 11},
    {"features1", 13, true, 14}, {"features1", 16, true, 17}, {"features1", 20, false, 21}, {"features1", 23, true, 24},
    {"features1", 26, false, -1}, {"features1", 27, true, -1}, {"features1", 30, false, 31}, {"features1", 33, true, 34},
    {"features1", 36, true, 37}, {"features1", 39, true, 40}, {"features2", 0, false, 1},  {"features2", 3, true, 4},
    {"features2", 6, true, 7},   {"features2", 9, true, 10},
};

// [cout][cin][3][3] kernel and bias of a plan entry.  A DEConv (:329-351) is folded exactly as its forward does:
// central difference (centre tap minus the tap sum, :218-235), horizontal / vertical difference built from Conv1d
// weights (:290-326), angular difference w - w[perm] (:238-255, theta = 1) and a plain 3x3 kernel; biases add.
int ggca_conv_weights(ff_cvit* h, const GgcaPlan& gp, int cin, int cout, std::vector<float>* w_out, std::vector<float>* b_out) {
  const std::string p = std::string(gp.seq) + "." + std::to_string(gp.conv_idx);
  auto bad = [&]() { return h->err.find("shape") != std::string::npos ? FF_ERR_SHAPE : FF_ERR_STATE; };
  if (!gp.de) {
    const auto* w = get_w(h, p + ".weight", {cout, cin, 3, 3});
    const auto* b = get_w(h, p + ".bias", {cout});
    if (!w || !b) return bad();
    *w_out = *w;
    *b_out = *b;
    return FF_OK;
  }
  const auto* w1 = get_w(h, p + ".conv1_1.conv.weight", {cout, cin, 3, 3});
  const auto* w2 = get_w(h, p + ".conv1_2.conv.weight", {cout, cin, 3});
  const auto* w3 = get_w(h, p + ".conv1_3.conv.weight", {cout, cin, 3});
  const auto* w4 = get_w(h, p + ".conv1_4.conv.weight", {cout, cin, 3, 3});
  const auto* w5 = get_w(h, p + ".conv1_5.weight", {cout, cin, 3, 3});
  const auto* b1 = get_w(h, p + ".conv1_1.conv.bias", {cout});
  const auto* b2 = get_w(h, p + ".conv1_2.conv.bias", {cout});
  const auto* b3 = get_w(h, p + ".conv1_3.conv.bias", {cout});
  const auto* b4 = get_w(h, p + ".conv1_4.conv.bias", {cout});
  const auto* b5 = get_w(h, p + ".conv1_5.bias", {cout});
  if (!w1 || !w2 || !w3 || !w4 || !w5 || !b1 || !b2 || !b3 || !b4 || !b5) return bad();
  static const int perm[9] = {3, 0, 1, 6, 4, 2, 7, 8, 5};
  w_out->assign((size_t)cout * cin * 9, 0.0f);
  b_out->resize(cout);
  for (size_t oi = 0; oi < (size_t)cout * cin; ++oi) {
    const float* a1 = &(*w1)[oi * 9];
    const float* a2 = &(*w2)[oi * 3];
    const float* a3 = &(*w3)[oi * 3];
    const float* a4 = &(*w4)[oi * 9];
    const float* a5 = &(*w5)[oi * 9];
    float sum1 = 0.0f;
    for (int t = 0; t < 9; ++t) sum1 += a1[t];
    float f[9];
    for (int t = 0; t < 9; ++t) {
      float cd = a1[t];
      if (t == 4) cd = a1[4] - sum1;
      float hd = 0.0f, vd = 0.0f;
      if (t % 3 == 0) hd = a2[t / 3];                 // taps 0,3,6 = +w, taps 2,5,8 = -w
      else if (t % 3 == 2) hd = -a2[t / 3];
      if (t < 3) vd = a3[t];                          // taps 0,1,2 = +w, taps 6,7,8 = -w
      else if (t >= 6) vd = -a3[t - 6];
      const float ad = a4[t] - a4[perm[t]];
      f[t] = (((cd + hd) + vd) + ad) + a5[t];         // the reference's summation order: w1 + w2 + w3 + w4 + w5
    }
    for (int t = 0; t < 9; ++t) (*w_out)[oi * 9 + t] = f[t];
  }
  for (int o = 0; o < cout; ++o) (*b_out)[o] = ((((*b1)[o] + (*b2)[o]) + (*b3)[o]) + (*b4)[o]) + (*b5)[o];
  return FF_OK;
}

// ------------------------------------------------------------------------------------------------ ResVitKan features
// ResNet-50 plan (ResVitKan.py:185-240): the stem is its own kernel (ff_rvk.cuh); every bottleneck convolution is
// one launch of tc_kernel: 1x1 stride-1 convs as GEMMs over pixels, 3x3 / strided convs as implicit GEMMs whose TMA
// descriptor carries the stride.  Eval-mode BN is folded into (scale, shift) of the producing launch.
constexpr int kRvkPlanes[4] = {64, 128, 256, 512}, kRvkBlocks[4] = {3, 4, 6, 3}, kRvkStride[4] = {1, 2, 2, 2};
constexpr size_t kRvkActElems = (size_t)112 * 112 * 64;      // largest activation per crop (stem / layer1 output)

void rvk_tile_geometry(int out_hw, int* bw, int* bh, int* bi) {
  if (out_hw == 56) { *bw = 8; *bh = 8; *bi = 2; }
  else if (out_hw == 28) { *bw = 4; *bh = 4; *bi = 8; }
  else if (out_hw == 14) { *bw = 2; *bh = 2; *bi = 32; }
  else { *bw = 8; *bh = 8; *bi = 2; }                         // 7x7: one masked 8x8 box per image
}

int rvk_fold_bn(ff_cvit* h, const std::string& bn, int c, std::vector<float>* scale, std::vector<float>* shift) {
  const auto* g = get_w(h, bn + ".weight", {c});
  const auto* be = get_w(h, bn + ".bias", {c});
  const auto* mu = get_w(h, bn + ".running_mean", {c});
  const auto* var = get_w(h, bn + ".running_var", {c});
  if (!g || !be || !mu || !var) return h->err.find("shape") != std::string::npos ? FF_ERR_SHAPE : FF_ERR_STATE;
  scale->resize(c);
  shift->resize(c);
  for (int o = 0; o < c; ++o) {
    const float s = (*g)[o] / std::sqrt((*var)[o] + BN_EPS);
    (*scale)[o] = s;
    (*shift)[o] = (*be)[o] - (*mu)[o] * s;                    // the ResNet convolutions have no bias (ResVitKan.py:191,198)
  }
  return FF_OK;
}

int rvk_add_op(ff_cvit* h, const std::string& conv, const std::string& bn, int cin, int cout, int k, int stride, int in_hw,
               int act, int in_buf, int out_buf, int resid, bf16* out_override = nullptr) {
  ff_cvit::RvkOp op;
  op.name = conv;
  op.cin = cin; op.cout = cout; op.taps = k * k; op.stride = stride; op.in_hw = in_hw; op.out_hw = in_hw / stride;
  op.type = (k == 1 && stride == 1) ? 0 : 1;
  op.act = act; op.resid = resid; op.in_buf = in_buf; op.out_buf = out_buf;
  op.bn = std::min(cout, 128);
  const auto* w = get_w(h, conv + ".weight", {cout, cin, k, k});
  if (!w) return h->err.find("shape") != std::string::npos ? FF_ERR_SHAPE : FF_ERR_STATE;
  std::vector<float> scale, shift, wr((size_t)cout * k * k * cin);
  int rc = rvk_fold_bn(h, bn, cout, &scale, &shift);
  if (rc) return rc;
  for (int o = 0; o < cout; ++o)
    for (int ci = 0; ci < cin; ++ci)
      for (int t = 0; t < k * k; ++t) wr[((size_t)o * k * k + t) * cin + ci] = (*w)[((size_t)o * cin + ci) * k * k + t];
  if ((rc = dev_upload(h, &op.w, to_bf16(wr)))) return rc;
  if ((rc = dev_upload(h, &op.scale, scale))) return rc;
  if ((rc = dev_upload(h, &op.shift, shift))) return rc;
  if ((rc = tmap_2d(h, &op.tmB, op.w, (uint64_t)k * k * cin, cout, 64, op.bn))) return rc;
  const bf16* in = h->rvk_buf[in_buf];
  if (op.type == 0) {
    rc = tmap_2d(h, &op.tmA, in, cin, (uint64_t)h->cap * in_hw * in_hw, 64, 128);
    if (!rc) rc = tmap_4d(h, &op.tmA_flat, in, cin, h->cap * in_hw * in_hw, 1, 1, 64, 128, 1, 1);
  } else {
    rvk_tile_geometry(op.out_hw, &op.bw, &op.bh, &op.bi);
    rc = tmap_4d(h, &op.tmA, in, cin, in_hw, in_hw, h->cap, 64, op.bw, op.bh, op.bi, stride);
  }
  if (rc) return rc;
  (void)out_override;
  {
    const bf16* outp = out_buf < 0 ? h->feat : h->rvk_buf[out_buf];
    const int ohw = op.out_hw;
    for (int k = 0; k < 2; ++k) {
      const bf16* base = k == 0 ? outp : (resid >= 0 ? h->rvk_buf[resid] : outp);
      CUtensorMap* m = k == 0 ? &op.tmO : &op.tmR;
      if (op.type == 0) rc = tmap_4d(h, m, base, cout, h->cap * ohw * ohw, 1, 1, 64, 128, 1, 1);
      else rc = tmap_4d(h, m, base, cout, ohw, ohw, h->cap, 64, op.bw, op.bh, op.bi);
      if (rc) return rc;
    }
  }
  h->rvk_ops.push_back(op);
  return FF_OK;
}

int finalize_rvk_features(ff_cvit* h) {
  int rc;
  // ---- stem: [64][3][7][7] -> [kh][cout][8 px][4 ch] with kw = px - 1 (ff_rvk.cuh)
  {
    const auto* w = get_w(h, "features.conv1.weight", {64, 3, 7, 7});
    if (!w) return h->err.find("shape") != std::string::npos ? FF_ERR_SHAPE : FF_ERR_STATE;
    std::vector<float> ws((size_t)7 * 64 * 32, 0.0f), scale, shift;
    for (int kh = 0; kh < 7; ++kh)
      for (int o = 0; o < 64; ++o)
        for (int kw = 0; kw < 7; ++kw)
          for (int c = 0; c < 3; ++c) ws[((size_t)kh * 64 + o) * 32 + (kw + 1) * 4 + c] = (*w)[(((size_t)o * 3 + c) * 7 + kh) * 7 + kw];
    if ((rc = dev_upload(h, &h->rvk_stem_w, to_bf16(ws)))) return rc;
    if ((rc = rvk_fold_bn(h, "features.bn1", 64, &scale, &shift))) return rc;
    for (int o = 0; o < 64; ++o) { h->rvk_stem_scale[o] = scale[o]; h->rvk_stem_shift[o] = shift[o]; }
    cuuint64_t dims[3] = {896, 224, (cuuint64_t)h->cap};
    cuuint64_t strides[2] = {896 * 2, (cuuint64_t)224 * 896 * 2};
    cuuint32_t box[3] = {96, 37, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = g_encode(&h->rvk_tm_x4, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, h->rvk_x4, dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(h, FF_ERR_CUDA, "cuTensorMapEncodeTiled(stem input) failed: %d", (int)r);
  }
  // ---- bottlenecks.  Buffers: X in {0,1} (block input / output, alternating), 2 = conv1 out, 3 = conv2 out, 4 = downsample
  h->rvk_ops.clear();
  int inplanes = 64, hw = 56, xb = 1;           // the max-pool writes buffer 1
  for (int li = 0; li < 4; ++li) {
    const int planes = kRvkPlanes[li];
    for (int b = 0; b < kRvkBlocks[li]; ++b) {
      const std::string p = "features.layer" + std::to_string(li + 1) + "." + std::to_string(b);
      const int stride = b == 0 ? kRvkStride[li] : 1;
      const int yb = xb ^ 1;
      if ((rc = rvk_add_op(h, p + ".conv1", p + ".bn1", inplanes, planes, 1, 1, hw, 1, xb, 2, -1))) return rc;
      if ((rc = rvk_add_op(h, p + ".conv2", p + ".bn2", planes, planes, 3, stride, hw, 1, 2, 3, -1))) return rc;
      int resid = xb;
      if (b == 0) {
        if ((rc = rvk_add_op(h, p + ".downsample.0", p + ".downsample.1", inplanes, planes * 4, 1, stride, hw, 0, xb, 4, -1))) return rc;
        resid = 4;
      }
      hw /= stride;
      // conv3 + bn3 + ReLU, + residual, + ReLU (ResVitKan.py:169-176: both ReLUs are in the reference)
      if ((rc = rvk_add_op(h, p + ".conv3", p + ".bn3", planes, planes * 4, 1, 1, hw, 1, 3, yb, resid))) return rc;
      inplanes = planes * 4;
      xb = yb;
    }
    h->rvk_layer_end[li] = (int)h->rvk_ops.size() - 1;
  }
  // features.channel + bn2 (no activation) writes the [n][49][512] patch vector the embedding GEMM reads
  if ((rc = rvk_add_op(h, "features.channel", "features.bn2", 2048, 512, 1, 1, 7, 0, xb, -1, -1))) return rc;
  // ---- KAN([2048, 64, 2])
  const int kin[2] = {MLP, 64}, kout[2] = {64, 2};
  for (int l = 0; l < 2; ++l) {
    const std::string q = "kan_head.3.layers." + std::to_string(l);
    const auto* bw = get_w(h, q + ".base_weight", {kout[l], kin[l]});
    const auto* sw = get_w(h, q + ".spline_weight", {kout[l], kin[l], 8});
    const auto* gr = get_w(h, q + ".grid", {kin[l], 12});
    if (!bw || !sw || !gr) return h->err.find("shape") != std::string::npos ? FF_ERR_SHAPE : FF_ERR_STATE;
    // enable_standalone_scale_spline=True is the KANLinear default (kan.py:19); accept checkpoints without the scaler
    const std::vector<float>* sc = nullptr;
    if (h->host_w.count(q + ".spline_scaler")) {
      sc = get_w(h, q + ".spline_scaler", {kout[l], kin[l]});
      if (!sc) return FF_ERR_SHAPE;
    }
    std::vector<float> pk((size_t)kin[l] * 9 * kout[l]);
    for (int i = 0; i < kin[l]; ++i)
      for (int o = 0; o < kout[l]; ++o) {
        pk[((size_t)i * 9) * kout[l] + o] = (*bw)[(size_t)o * kin[l] + i];
        const float s = sc ? (*sc)[(size_t)o * kin[l] + i] : 1.0f;
        for (int k = 0; k < 8; ++k) pk[((size_t)i * 9 + 1 + k) * kout[l] + o] = (*sw)[((size_t)o * kin[l] + i) * 8 + k] * s;
      }
    if ((rc = dev_upload(h, l == 0 ? &h->kan_w0 : &h->kan_w1, pk))) return rc;
    if ((rc = dev_upload(h, l == 0 ? &h->kan_g0 : &h->kan_g1, *gr))) return rc;
  }
  return FF_OK;
}

// kind 2: the BN-less, activation-less Conv2d(128,128) of features1.26 (P -> bufR) as one rvk_conv2_kernel op, and the
// GGCA shared convs with their BatchNorm2d(8) folded into the first one.
int finalize_ggca_extras(ff_cvit* h) {
  int rc;
  std::vector<float> wsrc, bias;
  if ((rc = ggca_conv_weights(h, kGgcaPlan[8], 128, 128, &wsrc, &bias))) return rc;
  ff_cvit::RvkOp op;
  op.name = "features1.26";
  op.type = 1; op.cin = 128; op.cout = 128; op.taps = 9; op.stride = 1; op.in_hw = 56; op.out_hw = 56;
  op.act = 0; op.resid = -1; op.bn = 128; op.bw = 8; op.bh = 8; op.bi = 2;
  op.out_ptr = h->bufR;
  std::vector<float> wr((size_t)128 * 9 * 128), ones(128, 1.0f);
  for (int o = 0; o < 128; ++o)
    for (int ci = 0; ci < 128; ++ci)
      for (int t = 0; t < 9; ++t) wr[((size_t)o * 9 + t) * 128 + ci] = wsrc[((size_t)o * 128 + ci) * 9 + t];
  if ((rc = dev_upload(h, &op.w, to_bf16(wr)))) return rc;
  if ((rc = dev_upload(h, &op.scale, ones))) return rc;
  if ((rc = dev_upload(h, &op.shift, bias))) return rc;
  if ((rc = tmap_2d(h, &op.tmB, op.w, 9 * 128, 128, 64, 128))) return rc;
  if ((rc = tmap_4d(h, &op.tmA, conv_output_buffer(h, 7), 128, 56, 56, h->cap, 64, 8, 8, 2))) return rc;
  if ((rc = tmap_4d(h, &op.tmO, h->bufR, 128, 56, 56, h->cap, 64, 8, 8, 2))) return rc;
  op.tmR = op.tmO;
  h->rvk_ops.clear();
  h->rvk_ops.push_back(op);
  // GGCA(512,7,7).shared_conv: Conv2d(128,8,1) + BatchNorm2d(8) + ReLU + Conv2d(8,128,1)  (:159-166)
  const auto* w1 = get_w(h, "ggca.shared_conv.0.weight", {8, 128, 1, 1});
  const auto* b1 = get_w(h, "ggca.shared_conv.0.bias", {8});
  const auto* g = get_w(h, "ggca.shared_conv.1.weight", {8});
  const auto* be = get_w(h, "ggca.shared_conv.1.bias", {8});
  const auto* mu = get_w(h, "ggca.shared_conv.1.running_mean", {8});
  const auto* var = get_w(h, "ggca.shared_conv.1.running_var", {8});
  const auto* w2 = get_w(h, "ggca.shared_conv.3.weight", {128, 8, 1, 1});
  const auto* b2 = get_w(h, "ggca.shared_conv.3.bias", {128});
  if (!w1 || !b1 || !g || !be || !mu || !var || !w2 || !b2) return h->err.find("shape") != std::string::npos ? FF_ERR_SHAPE : FF_ERR_STATE;
  std::vector<float> fw1(8 * 128), fb1(8);
  for (int u = 0; u < 8; ++u) {
    const float s = (*g)[u] / std::sqrt((*var)[u] + BN_EPS);
    for (int k = 0; k < 128; ++k) fw1[u * 128 + k] = (*w1)[u * 128 + k] * s;
    fb1[u] = ((*b1)[u] - (*mu)[u]) * s + (*be)[u];
  }
  if ((rc = dev_upload(h, &h->ggca_w1, fw1))) return rc;
  if ((rc = dev_upload(h, &h->ggca_b1, fb1))) return rc;
  if ((rc = dev_upload(h, &h->ggca_w2, *w2))) return rc;
  if ((rc = dev_upload(h, &h->ggca_b2, *b2))) return rc;
  return FF_OK;
}

int finalize(ff_cvit* h) {
  int rc;
  if (h->kind == 1) {
    if ((rc = finalize_rvk_features(h))) return rc;
  } else
  // ---- conv stack: fold bias + eval BN into (scale, shift); weights -> [cout][kh][kw][cin]
  for (int li = 0; li < 17; ++li) {
    const ConvPlan& p = kConv[li];
    std::vector<float> wsrc, bias_v, scale(p.cout), shift(p.cout), wr((size_t)p.cout * 9 * p.cin);
    std::string bn_key;
    if (h->kind == 2) {
      // plan entry of layer li: the extra BN-less conv (entry 8) sits between layers 8 and 9 (0-based li 7 and 8)
      const GgcaPlan& gp = kGgcaPlan[li < 8 ? li : li + 1];
      if ((rc = ggca_conv_weights(h, gp, p.cin, p.cout, &wsrc, &bias_v))) return rc;
      if (gp.bn_idx >= 0) bn_key = std::string(gp.seq) + "." + std::to_string(gp.bn_idx);
    } else {
      const std::string c = "features." + std::to_string(p.conv_idx);
      bn_key = "features." + std::to_string(p.conv_idx + 1);
      const auto* w = get_w(h, c + ".weight", {p.cout, p.cin, 3, 3});
      const auto* bias = get_w(h, c + ".bias", {p.cout});
      if (!w || !bias) return h->err.find("shape") != std::string::npos ? FF_ERR_SHAPE : FF_ERR_STATE;
      wsrc = *w;
      bias_v = *bias;
    }
    if (!bn_key.empty()) {
      const auto* g = get_w(h, bn_key + ".weight", {p.cout});
      const auto* be = get_w(h, bn_key + ".bias", {p.cout});
      const auto* mu = get_w(h, bn_key + ".running_mean", {p.cout});
      const auto* var = get_w(h, bn_key + ".running_var", {p.cout});
      if (!g || !be || !mu || !var) return h->err.find("shape") != std::string::npos ? FF_ERR_SHAPE : FF_ERR_STATE;
      for (int o = 0; o < p.cout; ++o) {
        const float s = (*g)[o] / std::sqrt((*var)[o] + BN_EPS);
        scale[o] = s;
        shift[o] = (bias_v[o] - (*mu)[o]) * s + (*be)[o];
      }
    } else {
      for (int o = 0; o < p.cout; ++o) { scale[o] = 1.0f; shift[o] = bias_v[o]; }     // BN-less DEConv (features1.27)
    }
    for (int o = 0; o < p.cout; ++o)
      for (int ci = 0; ci < p.cin; ++ci)
        for (int t = 0; t < 9; ++t) wr[((size_t)o * 9 + t) * p.cin + ci] = wsrc[((size_t)o * p.cin + ci) * 9 + t];
    ConvLayerDev& L = h->conv[li];
    if (p.cout <= 64)
      for (int o = 0; o < p.cout; ++o) { L.epi.scale[o] = scale[o]; L.epi.shift[o] = shift[o]; }
    if ((rc = dev_upload(h, &L.scale, scale))) return rc;
    if ((rc = dev_upload(h, &L.shift, shift))) return rc;
    if (li == 0) {
      std::vector<float> w32(32 * 64, 0.0f);     // tensor-core layout: k = kh*16 + kw*4 + cin (ff_c1.cuh)
      for (int o = 0; o < 32; ++o) {
        for (int k = 0; k < 27; ++k) {
          h->conv1.w[o][k] = wr[(size_t)o * 27 + k];
          const int tap = k / 3, c = k % 3, kh = tap / 3, kw = tap % 3;
          w32[o * 64 + kh * 16 + kw * 4 + c] = wr[(size_t)o * 27 + k];
        }
        h->conv1.scale[o] = scale[o];
        h->conv1.shift[o] = shift[o];
      }
      if (h->compute == FF_COMPUTE_BF16) {
        if ((rc = dev_upload(h, &h->c1_w, to_bf16(w32)))) return rc;
        {
          // pair-expanded filter: B[kh][(p,co)][(q,c)] = W[co][kh][q-p][c], zero unless 0 <= q-p <= 2 and c < 3
          std::vector<float> wp(3 * 64 * 16, 0.0f);
          for (int kh = 0; kh < 3; ++kh)
            for (int pp = 0; pp < 2; ++pp)
              for (int o = 0; o < 32; ++o)
                for (int q = 0; q < 4; ++q) {
                  const int kw = q - pp;
                  if (kw < 0 || kw > 2) continue;
                  for (int c = 0; c < 3; ++c) wp[(kh * 64 + pp * 32 + o) * 16 + q * 4 + c] = wr[(size_t)o * 27 + (kh * 3 + kw) * 3 + c];
                }
          if ((rc = dev_upload(h, &h->c1_wp, to_bf16(wp)))) return rc;
        }
        // lut[c][u] = bf16((u/255 - mean_c)/std_c), the fp32 arithmetic of cvit_prediction.py:41-45,214-215
        const float mean[3] = {0.485f, 0.456f, 0.406f}, sd[3] = {0.229f, 0.224f, 0.225f};
        std::vector<float> lut(3 * 256);
        for (int c = 0; c < 3; ++c)
          for (int u = 0; u < 256; ++u) lut[c * 256 + u] = ((float)u / 255.0f - mean[c]) / sd[c];
        if ((rc = dev_upload(h, &h->c1_lut, to_bf16(lut)))) return rc;
        // single-FMA form used by the TMA-fed kernel; keep it only if it reproduces the table bit for bit
        bool fma_ok = true;
        for (int c = 0; c < 3; ++c) {
          h->c1_na[c] = 1.0f / (255.0f * sd[c]);
          h->c1_nb[c] = -mean[c] / sd[c];
          for (int u = 0; u < 256; ++u) {
            const float f = std::fmaf((float)u, h->c1_na[c], h->c1_nb[c]);
            const bf16 x = __float2bfloat16(f), y = __float2bfloat16(lut[c * 256 + u]);
            if (memcmp(&x, &y, sizeof(bf16)) != 0) fma_ok = false;
          }
        }
        if (!fma_ok && h->use_c1_tc >= 2) h->use_c1_tc = 1;
      }
    }
    if (h->compute == FF_COMPUTE_FP32) {
      if ((rc = dev_upload(h, &L.wf, wr))) return rc;
    } else if (li > 0) {
      if ((rc = dev_upload(h, &L.w, to_bf16(wr)))) return rc;
      if (p.cin == 32) {
        // pair-expanded filter: B[(p,co)][(kh,q,ci)] = W[co][kh][q-p][ci], zero unless 0 <= q-p <= 2 (ff_ws.cuh)
        std::vector<float> w2((size_t)2 * p.cout * 384, 0.0f);
        for (int pp = 0; pp < 2; ++pp)
          for (int o = 0; o < p.cout; ++o)
            for (int kh = 0; kh < 3; ++kh)
              for (int q = 0; q < 4; ++q) {
                const int kw = q - pp;
                if (kw < 0 || kw > 2) continue;
                for (int ci = 0; ci < 32; ++ci)
                  w2[((size_t)pp * p.cout + o) * 384 + kh * 128 + q * 32 + ci] = wr[((size_t)o * 9 + kh * 3 + kw) * 32 + ci];
              }
        if ((rc = dev_upload(h, &L.w2, to_bf16(w2)))) return rc;
      }
      if (p.cin == 32 && p.cout == 32) {
        // quad-expanded filter, compact: per kh the 64-element k-blocks b = 0,1,2 (window pixels 2b, 2b+1) keep only the
        // output pixels p that see them: rows (p,co) for p in {0,1} | {0..3} | {2,3};  B[(p,co)][(wl,ci)] = W[co][kh][2b+wl-p][ci]
        std::vector<float> w4((size_t)768 * 64, 0.0f);
        const int p_lo[3] = {0, 0, 2}, p_n[3] = {2, 4, 2}, row_off[3] = {0, 64, 192};
        for (int kh = 0; kh < 3; ++kh)
          for (int b = 0; b < 3; ++b)
            for (int pi = 0; pi < p_n[b]; ++pi)
              for (int o = 0; o < 32; ++o)
                for (int wl = 0; wl < 2; ++wl) {
                  const int kw = 2 * b + wl - (p_lo[b] + pi);
                  if (kw < 0 || kw > 2) continue;
                  const size_t row = (size_t)kh * 256 + row_off[b] + pi * 32 + o;
                  for (int ci = 0; ci < 32; ++ci) w4[row * 64 + wl * 32 + ci] = wr[((size_t)o * 9 + kh * 3 + kw) * 32 + ci];
                }
        if ((rc = dev_upload(h, &L.w4, to_bf16(w4)))) return rc;
      }
      if (p.cin == 64 && p.cout == 64) {
        // CTA-pair variant: B[(pp,co)][(cb,kh,q,ci)] = W[co][kh][q-pp][cb*32+ci]
        std::vector<float> w2((size_t)128 * 768, 0.0f);
        for (int pp = 0; pp < 2; ++pp)
          for (int o = 0; o < 64; ++o)
            for (int cb = 0; cb < 2; ++cb)
              for (int kh = 0; kh < 3; ++kh)
                for (int q = 0; q < 4; ++q) {
                  const int kw = q - pp;
                  if (kw < 0 || kw > 2) continue;
                  for (int ci = 0; ci < 32; ++ci)
                    w2[((size_t)pp * 64 + o) * 768 + cb * 384 + kh * 128 + q * 32 + ci] = wr[((size_t)o * 9 + kh * 3 + kw) * 64 + cb * 32 + ci];
                }
        if ((rc = dev_upload(h, &L.w2x, to_bf16(w2)))) return rc;
      }
    }
  }
  // ---- embedding / tokens
  if ((rc = upload_linear(h, &h->embed, "patch_to_embedding", DIM, PATCH, true, 128))) return rc;
  {
    const auto* pos = get_w(h, "pos_embedding", {SLOTS, 1, DIM});
    const auto* cls = get_w(h, "cls_token", {1, 1, DIM});
    if (!pos || !cls) return FF_ERR_STATE;
    if ((rc = dev_upload(h, &h->pos, *pos))) return rc;
    if ((rc = dev_upload(h, &h->cls, *cls))) return rc;
  }
  // ---- transformer
  for (int l = 0; l < DEPTH; ++l) {
    const std::string p = "transformer.layers." + std::to_string(l);
    XfLayerDev& X = h->xf[l];
    if ((rc = upload_vec(h, &X.ln1_g, p + ".0.fn.norm.weight", DIM))) return rc;
    if ((rc = upload_vec(h, &X.ln1_b, p + ".0.fn.norm.bias", DIM))) return rc;
    // kind 2: the MLP branch is pre-normed by LinearNorm, which in eval() is its norm1 = LayerNorm(eps 1e-6)
    // (cvit_GGCA_ADD_DEConv_RepBn8.py:22-60); its RepBN / schedule buffers are not on the inference path
    const std::string ln2 = h->kind == 2 ? p + ".1.fn.norm.norm1" : p + ".1.fn.norm";
    if ((rc = upload_vec(h, &X.ln2_g, ln2 + ".weight", DIM))) return rc;
    if ((rc = upload_vec(h, &X.ln2_b, ln2 + ".bias", DIM))) return rc;
    if ((rc = upload_linear(h, &X.qkv, p + ".0.fn.fn.to_qkv", 3 * DIM, DIM, false, h->gemm_bn_wide))) return rc;
    if ((rc = upload_linear(h, &X.out, p + ".0.fn.fn.to_out", DIM, DIM, true, 64))) return rc;
    if ((rc = upload_linear(h, &X.ff1, p + ".1.fn.fn.net.0", MLP, DIM, true, h->gemm_bn_wide))) return rc;
    if ((rc = upload_linear(h, &X.ff2, p + ".1.fn.fn.net.2", DIM, MLP, true, 64))) return rc;
  }
  if (h->kind == 1) {     // kan_head = Linear, Dropout, ReLU, KAN (ResVitKan.py:302-307); mlp_head is not on the forward path
    if ((rc = upload_linear(h, &h->head1, "kan_head.0", MLP, DIM, true, 64))) return rc;
  } else {
    if ((rc = upload_linear(h, &h->head1, "mlp_head.0", MLP, DIM, true, 64))) return rc;
    if ((rc = upload_linear(h, &h->head2, "mlp_head.2", 2, MLP, true, 64))) return rc;
  }

  if (h->compute == FF_COMPUTE_BF16) {
    if (h->kind != 1 && (rc = build_conv_maps(h))) return rc;
    if (h->kind == 2 && (rc = finalize_ggca_extras(h))) return rc;
    const int cap128 = (h->cap + 127) / 128 * 128;
    if ((rc = tmap_2d(h, &h->tm_feat, h->feat, PATCH, cap128, 64, 128))) return rc;
    if ((rc = tmap_2d(h, &h->tm_xn, h->xn, DIM, h->rows_cap, 64, 128))) return rc;
    if ((rc = tmap_2d(h, &h->tm_att, h->att, DIM, h->rows_cap, 64, 128))) return rc;
    if ((rc = tmap_2d(h, &h->tm_ffh, h->ffh, MLP, h->rows_cap, 64, 128))) return rc;
    if ((rc = tmap_2d(h, &h->tm_cls, h->clsb, DIM, cap128, 64, 128))) return rc;
    if ((rc = xf_setup(h))) return rc;
  }
  h->host_w.clear();
  h->host_shape.clear();
  h->finalized = true;
  return FF_OK;
}

// ------------------------------------------------------------------------------------------------ encoder kernel
// One cooperative launch for the 6 transformer layers (ff_xf.cuh): groups of 16 CTAs, one group per 128-row token tile
// (groups loop over tiles when there are more tiles than co-resident groups).
int xf_setup(ff_cvit* h) {
  h->xf_ready = false;
  if (!h->use_xf || h->gemm_bn_wide != 64) return FF_OK;   // the kernel loads weights as 64-row boxes
  int coop = 0;
  FF_CUDA(h, cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, h->device));
  if (!coop) return FF_OK;
  std::vector<CUtensorMap> maps(3 + 4 * DEPTH);
  maps[0] = h->tm_xn;
  maps[1] = h->tm_att;
  maps[2] = h->tm_ffh;
  for (int l = 0; l < DEPTH; ++l) {
    maps[3 + 4 * l + 0] = h->xf[l].qkv.tmB;
    maps[3 + 4 * l + 1] = h->xf[l].out.tmB;
    maps[3 + 4 * l + 2] = h->xf[l].ff1.tmB;
    maps[3 + 4 * l + 3] = h->xf[l].ff2.tmB;
  }
  int rc = dev_alloc(h, &h->xf_maps, maps.size());
  if (rc) return rc;
  FF_CUDA(h, cudaMemcpy(h->xf_maps, maps.data(), maps.size() * sizeof(CUtensorMap), cudaMemcpyHostToDevice));
  if ((rc = dev_alloc(h, &h->xf_sync, (size_t)2 * XF_MAX_GROUPS))) return rc;
  FF_CUDA(h, cudaMemset(h->xf_sync, 0, 2 * XF_MAX_GROUPS * sizeof(unsigned int)));   // the kernel re-arms them itself
  FF_CUDA(h, cudaFuncSetAttribute(xf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, XF_SMEM_TOTAL));
  int per_sm = 0;
  FF_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, xf_kernel, XF_THREADS, XF_SMEM_TOTAL));
  h->xf_groups = std::min(per_sm * h->num_sms / XF_CS, (int)XF_MAX_GROUPS);
  if (getenv("FF_VERBOSE")) fprintf(stderr, "ff: encoder kernel: %d co-resident groups of %d CTAs\n", h->xf_groups, XF_CS);
  h->xf_ready = h->xf_groups >= 1;
  return FF_OK;
}

int launch_xf(ff_cvit* h, cudaStream_t st, int n, int depth) {
  XfArgs a;
  memset(&a, 0, sizeof(a));
  a.x = h->x; a.xn = h->xn; a.qkv = h->qkvb; a.att = h->att; a.ffh = h->ffh;
  a.maps = h->xf_maps;
  a.sync = h->xf_sync;
  a.rows = 2 * n; a.n_crops = n; a.depth = depth;
  a.eps1 = 1e-5f; a.eps2 = h->ln2_eps;
  for (int l = 0; l < DEPTH; ++l) {
    const XfLayerDev& X = h->xf[l];
    a.L[l] = XfLayerP{X.ln1_g, X.ln1_b, X.ln2_g, X.ln2_b, X.out.b, X.ff1.b, X.ff2.b};
  }
  const int tiles = (a.rows + 127) / 128;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(XF_CS * std::min(tiles, h->xf_groups));
  cfg.blockDim = dim3(XF_THREADS);
  cfg.dynamicSmemBytes = XF_SMEM_TOTAL;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;    // all CTAs co-resident: the group barriers spin on global counters
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  static long long* trace_buf = nullptr;
  static const bool trace_on = getenv("FF_XF_TRACE") != nullptr;
  if (trace_on && !trace_buf) cudaMalloc(&trace_buf, 64 * sizeof(long long));
  a.trace = trace_on ? trace_buf : nullptr;
  ProfScope ps(h, st, KC_GEMM_XF);
  cudaError_t e = cudaLaunchKernelEx(&cfg, xf_kernel, a);
  if (e == cudaErrorCooperativeLaunchTooLarge || e == cudaErrorNotSupported) {
    // the device cannot hold the groups any more (e.g. SMs reserved by another client): the per-op GPU launches take over
    cudaGetLastError();
    h->xf_ready = false;
    return 1;
  }
  if (e != cudaSuccess) return fail(h, FF_ERR_CUDA, "launch of the encoder kernel failed: %s", cudaGetErrorString(e));
  ++h->launches;
  if (trace_on) {   // developer aid: phase-boundary stamps (cycles between stamps) of CTA 0
    long long t[64];
    cudaStreamSynchronize(st);
    cudaMemcpy(t, trace_buf, sizeof(t), cudaMemcpyDeviceToHost);
    fprintf(stderr, "ff: xf trace (%d rows):", a.rows);
    for (int i = 1; i < (int)t[63] && i < 56; ++i) fprintf(stderr, " %lld", t[i] - t[i - 1]);
    fprintf(stderr, "\nff: xf trace producer: empty-wait %lld cycles over %lld k-blocks | mma thread: full-wait %lld, issue %lld, first-full..last-commit %lld\n",
            t[56], t[57], t[58], t[59], t[60]);
  }
  return FF_OK;
}

// ------------------------------------------------------------------------------------------------ forward
int launch_gemm(ff_cvit* h, cudaStream_t st, const CUtensorMap& tmA, const LinearDev& L, int M, void* out, int ldo, int epi,
                int act, int splits, const char* what) {
  TcArgs a;
  memset(&a, 0, sizeof(a));
  a.M = M;
  a.N = L.out_f;
  a.ldo = ldo;
  a.kb_total = L.in_f / 64;
  a.kb_per_split = (a.kb_total + splits - 1) / splits;
  a.shift = (splits > 1) ? nullptr : L.b;     // split-K: bias is added by the consumer of the partial slabs
  a.split_stride = (splits > 1) ? (long long)h->cap * ldo : 0;
  a.out = out;
  a.epi = epi;
  a.act = act;
  const int zs = (a.kb_total + a.kb_per_split - 1) / a.kb_per_split;
  dim3 grid((M + 127) / 128, L.out_f / L.bn, zs);
  ProfScope ps(h, st, &L == &h->embed ? KC_GEMM_EMBED : (&L == &h->head1 ? KC_GEMM_HEAD : KC_GEMM_XF));
  // (a deeper TMA ring - 8 stages - was measured slower: 192 KB of smem leaves one CTA per SM instead of two)
  cudaError_t e = L.bn == 128 ? launch_tc_t<MODE_GEMM, 128, 128, false, 4>(grid, st, tmA, L.tmB, a)
                              : launch_tc_t<MODE_GEMM, 128, 64, false, 4>(grid, st, tmA, L.tmB, a);
  if (e != cudaSuccess) return fail(h, FF_ERR_CUDA, "launch of gemm %s failed: %s", what, cudaGetErrorString(e));
  ++h->launches;
  return FF_OK;
}

struct DebugTap {
  int blocked_hw = 0;       // != 0: activation is channel-blocked [n][2][hw][hw][32] bf16
  int stop_after = 0;       // 0 = run everything
  const void* ptr = nullptr;
  int64_t elems = 0;
  bool is_bf16 = false;
  bool hit = false;
};

int forward_fp32(ff_cvit* h, const void* x, int layout, const int32_t* slot, int slot_base, int n, float* logits,
                 cudaStream_t st, DebugTap* tap);


// ---- ResVitKan feature extractor: input conversion, stem, max-pool, 16 bottlenecks, channel conv -> h->feat
int rvk_launch_op(ff_cvit* h, const ff_cvit::RvkOp& op, int n, cudaStream_t st, int prof_cls) {
  TcArgs a;
  memset(&a, 0, sizeof(a));
  a.scale = op.scale; a.shift = op.shift;
  a.out = op.out_buf < 0 ? h->feat : h->rvk_buf[op.out_buf];
  a.kb_per_tap = op.cin / 64;
  a.kb_total = op.taps * a.kb_per_tap;
  a.kb_per_split = a.kb_total;
  a.cin = op.cin;
  a.cout = op.cout;
  ProfScope ps(h, st, prof_cls);
  cudaError_t e;
  if (op.type == 0) {
    a.M = n * op.out_hw * op.out_hw;
    a.N = op.cout;
    a.ldo = op.cout;
    a.epi = EPI_BN_BF16;
    a.act = op.act ? ACT_RELU : ACT_NONE;
    a.resid = op.resid >= 0 ? h->rvk_buf[op.resid] : nullptr;
    a.act2 = op.resid >= 0 ? ACT_RELU : ACT_NONE;
    dim3 grid((a.M + 127) / 128, op.cout / op.bn, 1);
    e = op.bn == 128 ? launch_tc_t<MODE_GEMM, 128, 128, false, 4>(grid, st, op.tmA, op.tmB, a)
                     : launch_tc_t<MODE_GEMM, 128, 64, false, 4>(grid, st, op.tmA, op.tmB, a);
  } else {
    a.H = op.out_hw; a.W = op.out_hw;
    a.tiles_w = (op.out_hw + op.bw - 1) / op.bw; a.tiles_h = (op.out_hw + op.bh - 1) / op.bh;
    a.lg_bw = ilog2(op.bw); a.lg_bh = ilog2(op.bh);
    a.n_img = n;
    a.taps = op.taps; a.stride = op.stride;
    a.conv_act = op.act ? 0 : 1;
    dim3 grid(a.tiles_w * a.tiles_h * ((n + op.bi - 1) / op.bi), op.cout / op.bn, 1);
    e = launch_conv(128, op.bn, false, h->variant, grid, st, op.tmA, op.tmB, a);
  }
  if (e != cudaSuccess) return fail(h, FF_ERR_CUDA, "launch of %s failed: %s", op.name.c_str(), cudaGetErrorString(e));
  ++h->launches;
  return FF_OK;
}

template <int BN>
cudaError_t launch_rvk_conv_t(int grid, cudaStream_t st, const CUtensorMap& a, const CUtensorMap& b, const TcArgs& args) {
  using L = RvkSmem<BN, 2, 4>;
  auto k = rvk_conv_kernel<BN, 2, 4>;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL);
    if (e != cudaSuccess) return e;
    attr_done = true;
  }
  return launch_k(k, dim3(grid), dim3(320), L::TOTAL, st, true, a, b, args);
}

template <int BN, int STAGES, bool RESID>
cudaError_t launch_rvk_conv2_t(int grid, cudaStream_t st, const CUtensorMap& a, const CUtensorMap& b, const CUtensorMap& o,
                               const CUtensorMap& r, const TcArgs& args) {
  using L = Rvk2Smem<BN, STAGES, RESID>;
  auto k = rvk_conv2_kernel<BN, STAGES, RESID>;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL);
    if (e != cudaSuccess) return e;
    attr_done = true;
  }
  return launch_k(k, dim3(grid), dim3(320), L::TOTAL, st, true, a, b, o, r, args);
}

int rvk_launch_op_persistent(ff_cvit* h, const ff_cvit::RvkOp& op, int n, cudaStream_t st, int prof_cls) {
  TcArgs a;
  memset(&a, 0, sizeof(a));
  a.scale = op.scale; a.shift = op.shift;
  a.out = op.out_ptr ? op.out_ptr : (op.out_buf < 0 ? h->feat : h->rvk_buf[op.out_buf]);
  a.kb_per_tap = op.cin / 64;
  a.kb_total = op.taps * a.kb_per_tap;
  a.kb_per_split = a.kb_total;
  a.cin = op.cin;
  a.cout = op.cout;
  a.taps = op.taps; a.stride = op.stride;
  a.conv_act = op.act ? 0 : 1;
  a.resid = op.resid >= 0 ? h->rvk_buf[op.resid] : nullptr;
  int bi = 1;
  if (op.type == 0) {           // flat: W = every pixel of the pass
    a.H = 1; a.W = n * op.out_hw * op.out_hw;
    a.tiles_w = (a.W + 127) / 128; a.tiles_h = 1;
    a.lg_bw = 7; a.lg_bh = 0;
    a.n_img = 1;
  } else {
    a.H = op.out_hw; a.W = op.out_hw;
    a.tiles_w = (op.out_hw + op.bw - 1) / op.bw; a.tiles_h = (op.out_hw + op.bh - 1) / op.bh;
    a.lg_bw = ilog2(op.bw); a.lg_bh = ilog2(op.bh);
    a.n_img = n;
    bi = op.bi;
  }
  const int m_tiles = a.tiles_w * a.tiles_h * ((a.n_img + bi - 1) / bi);
  const int tiles = ((m_tiles + 1) / 2) * (op.cout / op.bn);
  const int grid = std::min(tiles, h->num_sms);
  ProfScope ps(h, st, prof_cls);
  const CUtensorMap& tmA = op.type == 0 ? op.tmA_flat : op.tmA;
  cudaError_t e;
  if (h->rvk_persist == 2) {
    if (op.resid >= 0 && op.bn == 128) e = launch_rvk_conv2_t<128, 2, true>(grid, st, tmA, op.tmB, op.tmO, op.tmR, a);
    else if (op.resid >= 0) return fail(h, FF_ERR_STATE, "%s: residual epilogue needs cout >= 128", op.name.c_str());
    else if (op.bn == 128) e = launch_rvk_conv2_t<128, 3, false>(grid, st, tmA, op.tmB, op.tmO, op.tmO, a);
    else e = launch_rvk_conv2_t<64, 4, false>(grid, st, tmA, op.tmB, op.tmO, op.tmO, a);
  } else {
    e = op.bn == 128 ? launch_rvk_conv_t<128>(grid, st, tmA, op.tmB, a) : launch_rvk_conv_t<64>(grid, st, tmA, op.tmB, a);
  }
  if (e != cudaSuccess) return fail(h, FF_ERR_CUDA, "launch of %s failed: %s", op.name.c_str(), cudaGetErrorString(e));
  ++h->launches;
  return FF_OK;
}

int rvk_features(ff_cvit* h, const void* x, int layout, int slot_base, int n, cudaStream_t st, DebugTap* tap) {
  const int stop = tap ? tap->stop_after : 0;
  auto tap_hit = [&](int step, const void* p, int64_t elems) {
    if (stop == step) { tap->ptr = p; tap->elems = elems; tap->is_bf16 = true; tap->hit = true; return true; }
    return false;
  };
  if (h->h2d_chunks_pending > 0) {     // host-buffer entry point: wait for the chunks covering this pass
    const int c1 = std::min(h->h2d_chunks_pending - 1, (slot_base + n - 1) / h->h2d_chunk);
    FF_CUDA(h, cudaStreamWaitEvent(st, h->h2d_ready[c1], 0));
  }
  {
    ProfScope ps(h, st, KC_CONV1);
    const unsigned blocks = (unsigned)(((size_t)n * 224 * 224 + 255) / 256);
    if (layout == FF_X_NHWC_U8) {
      // (u/255 - mean)/std as one fp32 FMA per channel (cvit_prediction.py:41-45 convention)
      const float mean[3] = {0.485f, 0.456f, 0.406f}, sd[3] = {0.229f, 0.224f, 0.225f};
      rvk_convert_kernel<2><<<(blocks + 3) / 4, 256, 0, st>>>(x, h->rvk_x4, n, 1.0f / (255.0f * sd[0]), -mean[0] / sd[0],
                                                    1.0f / (255.0f * sd[1]), -mean[1] / sd[1], 1.0f / (255.0f * sd[2]), -mean[2] / sd[2]);
    } else {
      rvk_convert_kernel<0><<<blocks, 256, 0, st>>>(x, h->rvk_x4, n, 1.f, 0.f, 1.f, 0.f, 1.f, 0.f);
    }
    FF_LAUNCH_CHECK(h, "rvk_convert");
    static bool stem_attr = false;
    if (!stem_attr) {
      FF_CUDA(h, cudaFuncSetAttribute(rvk_stem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, RVK_STEM_SMEM));
      stem_attr = true;
    }
    RvkStemArgs sa;
    sa.out = h->rvk_buf[0]; sa.w = h->rvk_stem_w; sa.n_img = n;
    for (int o = 0; o < 64; ++o) { sa.scale[o] = h->rvk_stem_scale[o]; sa.shift[o] = h->rvk_stem_shift[o]; }
    const int tiles = 14 * 7 * n;
    cudaError_t e = launch_k(rvk_stem_kernel, dim3(std::min(tiles, h->num_sms * 4)), dim3(128), RVK_STEM_SMEM, st, false, h->rvk_tm_x4, sa);
    if (e != cudaSuccess) return fail(h, FF_ERR_CUDA, "launch of the stem failed: %s", cudaGetErrorString(e));
    ++h->launches;
    rvk_maxpool_kernel<<<(unsigned)(((size_t)n * 56 * 56 * 8 + 255) / 256), 256, 0, st>>>(h->rvk_buf[0], h->rvk_buf[1], n);
    FF_LAUNCH_CHECK(h, "rvk_maxpool");
  }
  if (tap_hit(1, h->rvk_buf[1], (int64_t)n * 56 * 56 * 64)) return FF_OK;
  int layer = 0;
  // FF_RVK_PROF_LAYER=L (tools/rvk_profile.py --ops L): profile slots 1..16 = the first 16 ops of layer L, rest -> small
  static const int prof_layer = getenv("FF_RVK_PROF_LAYER") ? atoi(getenv("FF_RVK_PROF_LAYER")) : 0;
  int op_in_layer = 0;
  for (size_t i = 0; i < h->rvk_ops.size(); ++i) {
    const ff_cvit::RvkOp& op = h->rvk_ops[i];
    int cls = KC_TC_CONV + std::min(layer, 4);
    if (prof_layer) cls = (layer + 1 == prof_layer && op_in_layer < 16) ? KC_TC_CONV + op_in_layer : KC_SMALL;
    ++op_in_layer;
    if (layer < 4 && (int)i == h->rvk_layer_end[layer]) op_in_layer = 0;
    int rc = h->rvk_persist ? rvk_launch_op_persistent(h, op, n, st, cls) : rvk_launch_op(h, op, n, st, cls);
    if (rc) return rc;
    if (layer < 4 && (int)i == h->rvk_layer_end[layer]) {
      ++layer;
      if (tap_hit(1 + layer, h->rvk_buf[op.out_buf], (int64_t)n * op.out_hw * op.out_hw * op.cout)) return FF_OK;
    }
  }
  if (tap_hit(6, h->feat, (int64_t)n * PATCH)) return FF_OK;
  return FF_OK;
}

// One pass over n <= cap crops.  x points at the first crop of the pass.
int forward_pass(ff_cvit* h, const void* x, int layout, const int32_t* slot, int slot_base, int n, float* logits,
                 cudaStream_t st, DebugTap* tap) {
  if (h->compute == FF_COMPUTE_FP32) return forward_fp32(h, x, layout, slot, slot_base, n, logits, st, tap);
  const int stop = tap ? tap->stop_after : 0;
  auto tap_hit = [&](int step, const void* p, int64_t elems, bool is_b) {
    if (stop == step) { tap->ptr = p; tap->elems = elems; tap->is_bf16 = is_b; tap->hit = true; return true; }
    return false;
  };
  const size_t crop_in_bytes = layout == FF_X_NHWC_U8 ? (size_t)224 * 224 * 3 : (size_t)224 * 224 * 3 * 4;

  auto run_conv = [&](int li, int n_img, int img_off_out, int set, cudaStream_t st) -> int {
    const ConvPlan& p = kConv[li];
    const ConvLayerDev& L = (set && li <= 5) ? h->conv_alt[li] : h->conv[li];
    TcArgs a;
    memset(&a, 0, sizeof(a));
    a.H = p.hw; a.W = p.hw;
    a.tiles_w = p.hw / L.bw; a.tiles_h = p.hw / L.bh;
    a.lg_bw = ilog2(L.bw); a.lg_bh = ilog2(L.bh);
    a.n_img = n_img;
    a.img_off_out = img_off_out;
    a.cout = p.cout;
    a.cin = p.cin;
    a.kb_per_tap = p.cin / (L.rowb / 2);
    a.kb_total = 9 * a.kb_per_tap;
    a.kb_per_split = a.kb_total;
    a.scale = L.scale; a.shift = L.shift;
    a.out = conv_output_buffer(h, li, set);
    ProfScope ps(h, st, KC_TC_CONV + li - 1);
    if (L.ws2x) {
      a.tiles_w = p.hw / 16; a.tiles_h = p.hw / 16;
      a.out_blocked = (li == 4) ? 1 : 0;          // layer 5 feeds layer 6 (also a pair kernel); layer 6 writes plain NHWC
      const int tiles = a.tiles_w * a.tiles_h * n_img;
      const int g = std::min(2 * ((tiles + 1) / 2), h->num_sms & ~1);
      static long long* ws2x_dbg = nullptr;
      static const bool ws2x_dbg_on = getenv("FF_WS2X_DBG") != nullptr;
      if (ws2x_dbg_on) {
        if (!ws2x_dbg) cudaMalloc(&ws2x_dbg, 4 * sizeof(long long));
        a.resid = ws2x_dbg;
      }
      cudaError_t e = p.pool ? launch_ws2x_t<true>(g, st, L.tmA_ws2x, L.tmW_ws2x, a, L.epi)
                             : launch_ws2x_t<false>(g, st, L.tmA_ws2x, L.tmW_ws2x, a, L.epi);
      if (e != cudaSuccess) return fail(h, FF_ERR_CUDA, "launch of ws2x conv layer %d failed: %s", li + 1, cudaGetErrorString(e));
      ++h->launches;
      if (ws2x_dbg_on) {          // developer aid: cycles of the leader's MMA thread per tile pair (serialises the stream)
        long long hd[4];
        cudaStreamSynchronize(st);
        cudaMemcpy(hd, ws2x_dbg, sizeof(hd), cudaMemcpyDeviceToHost);
        if (hd[3] > 0)
          fprintf(stderr, "[ws2x layer %d] tile pairs %lld | wait TMEM drained %lld | wait patch %lld | issue 48 MMAs %lld cycles/pair\n", li + 1,
                  hd[3], hd[0] / hd[3], hd[1] / hd[3], hd[2] / hd[3]);
      }
      return FF_OK;
    }
    if (L.ws4) {
      a.tiles_w = p.hw / 32; a.tiles_h = p.hw / 16;
      const int tiles = a.tiles_w * a.tiles_h * n_img;
      const int g = std::min(tiles, h->num_sms);
      cudaError_t e = p.pool ? launch_ws4_t<true>(g, st, L.tmA_ws4, L.tmW_ws4, a, L.epi) : launch_ws4_t<false>(g, st, L.tmA_ws4, L.tmW_ws4, a, L.epi);
      if (e != cudaSuccess) return fail(h, FF_ERR_CUDA, "launch of quad conv layer %d failed: %s", li + 1, cudaGetErrorString(e));
      ++h->launches;
      return FF_OK;
    }
    if (L.ws2) {
      a.out_blocked = (li == 3 && h->conv[4].ws2x) ? 1 : 0;   // layer 4 feeds the CTA-pair kernel of layer 5
      a.tiles_w = p.hw / 16; a.tiles_h = p.hw / 16;
      const int tiles = a.tiles_w * a.tiles_h * n_img;
      cudaError_t e;
      if (p.cout == 32) {
        const int g = std::min(tiles, h->num_sms * 2);
        e = p.pool ? launch_ws2_t<64, true, 2>(g, st, L.tmA_ws2, L.tmW_ws2, a, L.epi)
                   : launch_ws2_t<64, false, 2>(g, st, L.tmA_ws2, L.tmW_ws2, a, L.epi);
      } else {
        e = launch_ws2_t<128, false, 4>(std::min(tiles, h->num_sms), st, L.tmA_ws2, L.tmW_ws2, a, L.epi);
      }
      if (e != cudaSuccess) return fail(h, FF_ERR_CUDA, "launch of ws2 conv layer %d failed: %s", li + 1, cudaGetErrorString(e));
      ++h->launches;
      return FF_OK;
    }
    if (L.ws) {
      a.tiles_w = p.hw / 8; a.tiles_h = p.hw / 16;
      a.lg_bw = 3; a.lg_bh = 4;
      const int tiles = a.tiles_w * a.tiles_h * n_img;
      const int g = std::min(tiles, h->num_sms * (p.cin == 32 ? h->ws_ctas_per_sm : 1));
      cudaError_t e = launch_ws(p.cin, p.cout, p.pool, h->ws_ctas_per_sm, g, st, L.tmA_ws, L.tmW_ws, a, L.epi);
      if (e != cudaSuccess) return fail(h, FF_ERR_CUDA, "launch of ws conv layer %d failed: %s", li + 1, cudaGetErrorString(e));
      ++h->launches;
      return FF_OK;
    }
    if (h->use_ptc && L.rowb == 128 && (L.bn == 128 || L.bn == 256)) {
      const int msub = L.bn == 128 ? 2 : 1;
      const int m_tiles = a.tiles_w * a.tiles_h * ((n_img + L.bi - 1) / L.bi);
      const int tiles = ((m_tiles + msub - 1) / msub) * (p.cout / L.bn);
      const int g = std::min(tiles, h->num_sms);
      cudaError_t e;
      if (L.ptc2) {          // one item = two pixel tiles x one 256-channel tile on a CTA pair
        const int items = ((m_tiles + 1) / 2) * (p.cout / 256);
        const int g2 = std::min(2 * items, h->num_sms & ~1);
        e = p.pool ? launch_ptc2_t<true>(g2, st, L.tmA, L.tmB_half, a) : launch_ptc2_t<false>(g2, st, L.tmA, L.tmB_half, a);
      } else if (L.ptc2m) {   // unvalidated opt-in: one item = four pixel tiles x the 128-channel tile on a CTA pair
        const int items = ((m_tiles + 3) / 4) * (p.cout / 128);
        const int g2 = std::min(2 * items, h->num_sms & ~1);
        e = p.pool ? launch_ptc2m_t<true>(g2, st, L.tmA, L.tmB_half, a) : launch_ptc2m_t<false>(g2, st, L.tmA, L.tmB_half, a);
      } else
      if (L.bn == 128) e = p.pool ? launch_ptc_t<128, 2, true, 4>(g, st, L.tmA, L.tmB, a) : launch_ptc_t<128, 2, false, 4>(g, st, L.tmA, L.tmB, a);
      else e = p.pool ? launch_ptc_t<256, 1, true, 4>(g, st, L.tmA, L.tmB, a) : launch_ptc_t<256, 1, false, 4>(g, st, L.tmA, L.tmB, a);
      if (e != cudaSuccess) return fail(h, FF_ERR_CUDA, "launch of persistent conv layer %d failed: %s", li + 1, cudaGetErrorString(e));
      ++h->launches;
      return FF_OK;
    }
    dim3 grid(a.tiles_w * a.tiles_h * ((n_img + L.bi - 1) / L.bi), p.cout / L.bn, 1);
    cudaError_t e = launch_conv(L.rowb, L.bn, p.pool, h->variant, grid, st, L.tmA, L.tmB, a);
    if (e != cudaSuccess) return fail(h, FF_ERR_CUDA, "launch of conv layer %d failed: %s", li + 1, cudaGetErrorString(e));
    ++h->launches;
    return FF_OK;
  };

  if (h->kind == 1) {
    int rc = rvk_features(h, x, layout, slot_base, n, st, tap);
    if (rc || (tap && tap->hit)) return rc;
  } else {
  // ---- stages 1-2 in sub-passes of s12 crops (activations of 3.2 MB/crop stay L2-resident between layers)
  prof_mark(h, st, 0, true, true);
  const int sub = stop ? std::min(n, h->s12_cap) : h->s12;
  if (stop && n > h->s12_cap) return fail(h, FF_ERR_BAD_ARG, "debug tap needs n <= %d", h->s12_cap);
  // consecutive sub-passes alternate between the caller's stream and aux_stream (own ping-pong buffers), so the
  // launch tail / prologue of one chain is filled by the other chain's kernels
  const bool dual = h->use_dual && !stop && (!h->profiling || h->prof_coarse) && n > sub;
  cudaStream_t st_main = st;
  if (dual) {
    FF_CUDA(h, cudaEventRecord(h->ev_fork, st_main));
    FF_CUDA(h, cudaStreamWaitEvent(h->aux_stream, h->ev_fork, 0));
  }
  int sub_idx = 0;
  for (int s0 = 0; s0 < n; s0 += sub, ++sub_idx) {
    const int set = dual ? (sub_idx & 1) : 0;
    cudaStream_t st = set ? h->aux_stream : st_main;
    bf16* bufA = set ? h->bufA2 : h->bufA;
    const int ns = std::min(sub, n - s0);
    const uint8_t* xin = reinterpret_cast<const uint8_t*>(x) + (size_t)s0 * crop_in_bytes;
    if (h->h2d_chunks_pending > 0) {   // input still streaming in: wait for the chunks covering [g0, g0+ns)
      const int g0 = slot_base + s0;   // slot_base == offset of this pass inside the whole batch
      const int c1 = std::min(h->h2d_chunks_pending - 1, (g0 + ns - 1) / h->h2d_chunk);
      FF_CUDA(h, cudaStreamWaitEvent(st, h->h2d_ready[c1], 0));
    }
    dim3 g1(14, 14, ns);
    // layers 1 + 2 in one kernel when the uint8 fast path is active and nobody asks for layer 1's output
    const bool fused12 = h->use_c12 && h->use_c1_tc == 3 && layout == FF_X_NHWC_U8 && h->conv[1].ws2 && !h->conv[1].ws4 && stop != 1;
    if (fused12) {
      ProfScope ps(h, st, KC_TC_CONV);
      CUtensorMap tmX;
      cuuint64_t dims[3] = {672, 224, (cuuint64_t)ns};
      cuuint64_t strides[2] = {672, (cuuint64_t)224 * 672};
      cuuint32_t box[3] = {80, 18, 1};
      cuuint32_t estr[3] = {1, 1, 1};
      CUresult r = g_encode(&tmX, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<uint8_t*>(xin), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) return fail(h, FF_ERR_CUDA, "cuTensorMapEncodeTiled(uint8 crops) failed: %d (is the input 16-byte aligned?)", (int)r);
      C12Args ca;
      ca.out = conv_output_buffer(h, 1, set); ca.w1 = h->c1_wp; ca.n_img = ns;
      for (int c = 0; c < 3; ++c) { ca.na[c] = h->c1_na[c]; ca.nb[c] = h->c1_nb[c]; }
      for (int o = 0; o < 32; ++o) {
        ca.scale1[o] = h->conv1.scale[o]; ca.shift1[o] = h->conv1.shift[o];
        ca.scale2[o] = h->conv[1].epi.scale[o]; ca.shift2[o] = h->conv[1].epi.shift[o];
      }
      ca.dbg = nullptr;
      static long long* c12_dbg = nullptr;
      static const bool c12_dbg_on = getenv("FF_C12_DBG") != nullptr;
      if (c12_dbg_on) {
        if (!c12_dbg) cudaMalloc(&c12_dbg, 8 * sizeof(long long));
        ca.dbg = c12_dbg;
      }
      static bool c12_attr = false;
      if (!c12_attr) {
        FF_CUDA(h, cudaFuncSetAttribute(c12_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, C12Smem::TOTAL));
        c12_attr = true;
      }
      const ConvLayerDev& L2 = set ? h->conv_alt[1] : h->conv[1];
      const int grid = std::min(16 * 14 * ns, h->num_sms * 2);
      cudaError_t e = launch_k(c12_kernel, dim3(grid), dim3(C12_THREADS), C12Smem::TOTAL, st, true, tmX, L2.tmW_ws2, ca);
      if (e != cudaSuccess) return fail(h, FF_ERR_CUDA, "launch of the fused layer-1/2 kernel failed: %s", cudaGetErrorString(e));
      ++h->launches;
      if (c12_dbg_on) {          // developer aid: cycles per pipeline phase of CTA 0 (serialises the stream)
        long long hd[8];
        cudaStreamSynchronize(st);
        cudaMemcpy(hd, c12_dbg, sizeof(hd), cudaMemcpyDeviceToHost);
        if (hd[5] > 0)
          fprintf(stderr, "[c12] tiles %lld | wait conv1 %lld | epi1+sync %lld | issue conv2 %lld | convert+issue conv1 %lld | wait conv2+epi2 %lld cycles/tile\n",
                  hd[5], hd[0] / hd[5], hd[1] / hd[5], hd[2] / hd[5], hd[3] / hd[5], hd[4] / hd[5]);
      }
    } else {
      ProfScope ps(h, st, KC_CONV1);
      if (h->use_c1_tc == 3 && layout == FF_X_NHWC_U8) {
        CUtensorMap tmX;
        cuuint64_t dims[3] = {672, 224, (cuuint64_t)ns};
        cuuint64_t strides[2] = {672, (cuuint64_t)224 * 672};
        cuuint32_t box[3] = {80, 18, 1};
        cuuint32_t estr[3] = {1, 1, 1};
        CUresult r = g_encode(&tmX, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<uint8_t*>(xin), dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail(h, FF_ERR_CUDA, "cuTensorMapEncodeTiled(uint8 crops) failed: %d (is the input 16-byte aligned?)", (int)r);
        C1PairArgs ca;
        ca.out = bufA; ca.w = h->c1_wp; ca.n_img = ns;
        for (int c = 0; c < 3; ++c) { ca.na[c] = h->c1_na[c]; ca.nb[c] = h->c1_nb[c]; }
        for (int o = 0; o < 32; ++o) { ca.scale[o] = h->conv1.scale[o]; ca.shift[o] = h->conv1.shift[o]; }
        const int grid = std::min(196 * ns, h->num_sms * h->c1_ctas_per_sm);
        launch_k(conv1_pair_kernel, dim3(grid), dim3(128), 0, st, true, tmX, ca);
      } else if (h->use_c1_tc >= 2 && layout == FF_X_NHWC_U8) {
        // TMA-fed uint8 path: per-launch 3-D map over the caller's uint8 crops viewed as [ns][224][672]
        CUtensorMap tmX;
        cuuint64_t dims[3] = {672, 224, (cuuint64_t)ns};
        cuuint64_t strides[2] = {672, (cuuint64_t)224 * 672};
        cuuint32_t box[3] = {48, 18, 1};
        cuuint32_t estr[3] = {1, 1, 1};
        CUresult r = g_encode(&tmX, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<uint8_t*>(xin), dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail(h, FF_ERR_CUDA, "cuTensorMapEncodeTiled(uint8 crops) failed: %d (is the input 16-byte aligned?)", (int)r);
        C1TmaArgs ca;
        ca.out = bufA; ca.w = h->c1_w; ca.n_img = ns;
        for (int c = 0; c < 3; ++c) { ca.na[c] = h->c1_na[c]; ca.nb[c] = h->c1_nb[c]; }
        for (int o = 0; o < 32; ++o) { ca.scale[o] = h->conv1.scale[o]; ca.shift[o] = h->conv1.shift[o]; }
        const int grid = std::min(392 * ns, h->num_sms * h->c1_ctas_per_sm);
        launch_k(conv1_tma_kernel, dim3(grid), dim3(128), 0, st, true, tmX, ca);
      } else if (h->use_c1_tc) {
        C1Args ca;
        ca.x = xin; ca.out = bufA; ca.w = h->c1_w; ca.lut = h->c1_lut;
        ca.n_img = ns;
        for (int o = 0; o < 32; ++o) { ca.scale[o] = h->conv1.scale[o]; ca.shift[o] = h->conv1.shift[o]; }
        const int grid = std::min(392 * ns, h->num_sms * h->c1_ctas_per_sm);
        if (layout == FF_X_NHWC_U8) launch_k(conv1_tc_kernel<2>, dim3(grid), dim3(128), 0, st, true, ca);
        else launch_k(conv1_tc_kernel<0>, dim3(grid), dim3(128), 0, st, true, ca);
      } else if (layout == FF_X_NHWC_U8) conv1_kernel<2><<<g1, 256, 0, st>>>(xin, bufA, ns, h->conv1);
      else conv1_kernel<0><<<g1, 256, 0, st>>>(xin, bufA, ns, h->conv1);
    }
    if (!fused12) {
      FF_LAUNCH_CHECK(h, "conv1");
      if (tap_hit(1, bufA, (int64_t)ns * 224 * 224 * 32, true)) return FF_OK;
    }
    for (int li = 1; li <= 5; ++li) {
      int rc = (fused12 && li == 1) ? FF_OK : run_conv(li, ns, li == 5 ? s0 : 0, set, st);
      if (rc) return rc;
      const ConvPlan& p = kConv[li];
      const int ohw = p.pool ? p.hw / 2 : p.hw;
      if (tap_hit(li + 1, conv_output_buffer(h, li, set), (int64_t)ns * ohw * ohw * p.cout, true)) {
        if ((li == 3 || li == 4) && h->conv[4].ws2x) tap->blocked_hw = ohw;
        return FF_OK;
      }
    }
  }
  if (dual) {
    FF_CUDA(h, cudaEventRecord(h->ev_join, h->aux_stream));
    FF_CUDA(h, cudaStreamWaitEvent(st_main, h->ev_join, 0));
  }
  prof_mark(h, st, 0, false, true);
  prof_mark(h, st, 1, true, true);
  // ---- stages 3-5 on the whole pass
  for (int li = 6; li < 17; ++li) {
    int rc = run_conv(li, n, 0, 0, st);
    if (rc) return rc;
    const ConvPlan& p = kConv[li];
    const int ohw = p.pool ? p.hw / 2 : p.hw;
    if (tap_hit(li + 1, conv_output_buffer(h, li), (int64_t)n * ohw * ohw * p.cout, true)) return FF_OK;
    if (h->kind == 2 && li == 7) {   // features1.26: Conv2d(128,128), no BN, no activation -> bufR (input of layer 9)
      if ((rc = rvk_launch_op_persistent(h, h->rvk_ops[0], n, st, KC_SMALL))) return rc;
      if (tap_hit(26, h->bufR, (int64_t)n * 56 * 56 * 128, true)) return FF_OK;
    }
  }
  if (h->kind == 2) {                // x = x * ggca(x)  (cvit_GGCA_ADD_DEConv_RepBn8.py:447-448), in place on feat
    ProfScope ps(h, st, KC_SMALL);
    ggca_gate_kernel<<<n, 512, 0, st>>>(h->feat, h->ggca_w1, h->ggca_b1, h->ggca_w2, h->ggca_b2, n);
    FF_LAUNCH_CHECK(h, "ggca_gate");
    if (tap_hit(27, h->feat, (int64_t)n * PATCH, true)) return FF_OK;
  }
  prof_mark(h, st, 1, false, true);
  }   // kind == 0
  prof_mark(h, st, 2, true, true);
  // ---- patch embedding (split-K, fp32 atomics) + token assembly
  int rc = launch_gemm(h, st, h->tm_feat, h->embed, n, h->emb, DIM, EPI_STORE_F32, ACT_NONE, EMBED_SPLITS, "patch_to_embedding");
  if (rc) return rc;
  { ProfScope ps(h, st, KC_SMALL); launch_k(tokens_kernel, dim3(n), dim3(256), 0, st, true, (const float*)h->emb, (int)EMBED_SPLITS, (long long)h->cap * DIM, (const float*)h->embed.b, (const float*)h->cls, (const float*)h->pos, slot, slot_base, h->x, n); }
  FF_LAUNCH_CHECK(h, "tokens");
  const int rows = 2 * n;
  if (tap_hit(18, h->x, (int64_t)rows * DIM, false)) return FF_OK;
  // ---- transformer: one cooperative launch for all layers (a debug tap inside the encoder shortens the depth)
  bool encoder_done = false;
  if (h->xf_ready) {
    const int depth = (stop >= 19 && stop < 19 + DEPTH) ? stop - 18 : DEPTH;
    rc = launch_xf(h, st, n, depth);
    if (rc < 0) return rc;
    if (rc == 0) {
      if (tap_hit(18 + depth, h->x, (int64_t)rows * DIM, false)) return FF_OK;
      encoder_done = true;
    }
  }
  for (int l = 0; l < DEPTH && !encoder_done; ++l) {
    const XfLayerDev& X = h->xf[l];
    { ProfScope ps(h, st, KC_SMALL); launch_k(layernorm_kernel, dim3((rows + 7) / 8), dim3(256), 0, st, true, (const float*)h->x, (const float*)X.ln1_g, (const float*)X.ln1_b, h->xn, rows, 1e-5f); }
    FF_LAUNCH_CHECK(h, "layernorm1");
    if ((rc = launch_gemm(h, st, h->tm_xn, X.qkv, rows, h->qkvb, 3 * DIM, EPI_STORE_BF16, ACT_NONE, 1, "to_qkv"))) return rc;
    { ProfScope ps(h, st, KC_SMALL); launch_k(attention2_kernel, dim3((n * 8 + 7) / 8), dim3(256), 0, st, true, (const bf16*)h->qkvb, h->att, n); }
    FF_LAUNCH_CHECK(h, "attention2");
    if ((rc = launch_gemm(h, st, h->tm_att, X.out, rows, h->x, DIM, EPI_RESID_F32, ACT_NONE, 1, "to_out"))) return rc;
    { ProfScope ps(h, st, KC_SMALL); launch_k(layernorm_kernel, dim3((rows + 7) / 8), dim3(256), 0, st, true, (const float*)h->x, (const float*)X.ln2_g, (const float*)X.ln2_b, h->xn, rows, h->ln2_eps); }
    FF_LAUNCH_CHECK(h, "layernorm2");
    if ((rc = launch_gemm(h, st, h->tm_xn, X.ff1, rows, h->ffh, MLP, EPI_STORE_BF16, ACT_GELU, 1, "ff1"))) return rc;
    if ((rc = launch_gemm(h, st, h->tm_ffh, X.ff2, rows, h->x, DIM, EPI_RESID_F32, ACT_NONE, 1, "ff2"))) return rc;
    if (tap_hit(19 + l, h->x, (int64_t)rows * DIM, false)) return FF_OK;
  }
  // ---- head
  { ProfScope ps(h, st, KC_SMALL); launch_k(cls_gather_kernel, dim3(n), dim3(256), 0, st, true, (const float*)h->x, h->clsb, n); }
  FF_LAUNCH_CHECK(h, "cls_gather");
  if ((rc = launch_gemm(h, st, h->tm_cls, h->head1, n, h->hid, MLP, EPI_STORE_F32, ACT_RELU, 1, "head.0"))) return rc;
  if (h->kind == 1) {
    ProfScope ps(h, st, KC_SMALL);
    kan_l0_kernel<<<dim3(KAN_CHUNKS, (n + KAN_SG - 1) / KAN_SG), 256, 0, st>>>(h->hid, h->kan_w0, h->kan_g0, h->kan_part, n, h->cap);
    FF_LAUNCH_CHECK(h, "kan_l0");
    kan_l1_kernel<<<(n + 7) / 8, 256, 0, st>>>(h->kan_part, h->kan_w1, h->kan_g1, logits, n, h->cap);
    FF_LAUNCH_CHECK(h, "kan_l1");
  } else {
    { ProfScope ps(h, st, KC_SMALL); launch_k(head2_kernel, dim3((n + 7) / 8), dim3(256), 0, st, true, (const float*)h->hid, (const float*)h->head2.wf, (const float*)h->head2.b, logits, n); }
    FF_LAUNCH_CHECK(h, "head2");
  }
  prof_mark(h, st, 2, false, true);
  if (tap_hit(25, logits, (int64_t)n * 2, false)) return FF_OK;
  return FF_OK;
}

// ---- fp32 CUDA-core path (parity to 1e-4; reuses LN/attention-like kernels in fp32)
int forward_fp32(ff_cvit* h, const void* x, int layout, const int32_t* slot, int slot_base, int n, float* logits,
                 cudaStream_t st, DebugTap* tap) {
  const int stop = tap ? tap->stop_after : 0;
  auto tap_hit = [&](int step, const void* p, int64_t elems) {
    if (stop == step) { tap->ptr = p; tap->elems = elems; tap->is_bf16 = false; tap->hit = true; return true; }
    return false;
  };
  if (h->h2d_chunks_pending > 0) FF_CUDA(h, cudaStreamWaitEvent(st, h->h2d_ready[h->h2d_chunks_pending - 1], 0));
  // conv stack, NHWC fp32, ping-pong fA/fB; processed `chunk` crops at a time to bound the workspace
  const int chunk = h->s12_cap;
  const size_t crop_in_bytes = layout == FF_X_NHWC_U8 ? (size_t)224 * 224 * 3 : (size_t)224 * 224 * 3 * 4;
  float* featf = h->fB + (size_t)chunk * 224 * 224 * 32;    // tail of fB is reserved for [cap][25088]
  if (stop && n > chunk) return fail(h, FF_ERR_BAD_ARG, "debug tap needs n <= %d", chunk);
  for (int s0 = 0; s0 < n; s0 += chunk) {
    const int ns = std::min(chunk, n - s0);
    const uint8_t* xin = reinterpret_cast<const uint8_t*>(x) + (size_t)s0 * crop_in_bytes;
    float* cur = h->fA;
    float* nxt = h->fB;
    for (int li = 0; li < 17; ++li) {
      const ConvPlan& p = kConv[li];
      const ConvLayerDev& L = h->conv[li];
      float* dst = (li == 16) ? featf + (size_t)s0 * PATCH : nxt;
      const int ohw = p.pool ? p.hw / 2 : p.hw;
      const size_t total = (size_t)ns * ohw * ohw * p.cout;
      const int blocks = (int)std::min<size_t>((total + 255) / 256, 1u << 30);
      if (li == 0)
        conv3x3_fp32_kernel<<<blocks, 256, 0, st>>>(xin, layout == FF_X_NHWC_U8 ? 2 : 1, nullptr, L.wf, L.scale, L.shift, dst,
                                                   ns, p.hw, p.cin, p.cout, p.pool ? 1 : 0);
      else
        conv3x3_fp32_kernel<<<blocks, 256, 0, st>>>(nullptr, 0, cur, L.wf, L.scale, L.shift, dst, ns, p.hw, p.cin, p.cout,
                                                   p.pool ? 1 : 0);
      FF_LAUNCH_CHECK(h, "conv3x3_fp32");
      if (tap_hit(li + 1, dst, (int64_t)total)) return FF_OK;
      std::swap(cur, nxt);
    }
  }
  const int rows = 2 * n;
  float* xn = reinterpret_cast<float*>(h->fA);                  // [rows][1024]
  float* att = xn + (size_t)h->rows_cap * DIM;                  // [rows][1024]
  float* ffh = att + (size_t)h->rows_cap * DIM;                 // [rows][2048]
  auto lin = [&](const float* A, const LinearDev& L, int M, float* out, int act, int resid, const char* what) -> int {
    dim3 grid((L.out_f + 63) / 64, (M + 63) / 64);
    linear_fp32_kernel<<<grid, 256, 0, st>>>(A, L.wf, L.b, out, M, L.out_f, L.in_f, act, resid);
    FF_LAUNCH_CHECK(h, what);
    return FF_OK;
  };
  int rc;
  if ((rc = lin(featf, h->embed, n, h->emb, 0, 0, "embed_fp32"))) return rc;
  tokens_kernel<<<n, 256, 0, st>>>(h->emb, 1, 0, nullptr, h->cls, h->pos, slot, slot_base, h->x, n);
  FF_LAUNCH_CHECK(h, "tokens");
  if (tap_hit(18, h->x, (int64_t)rows * DIM)) return FF_OK;
  for (int l = 0; l < DEPTH; ++l) {
    const XfLayerDev& X = h->xf[l];
    layernorm_f32_kernel<<<(rows + 7) / 8, 256, 0, st>>>(h->x, X.ln1_g, X.ln1_b, xn, rows);
    FF_LAUNCH_CHECK(h, "layernorm_f32");
    if ((rc = lin(xn, X.qkv, rows, h->qkv, 0, 0, "qkv_fp32"))) return rc;
    attention2_f32_kernel<<<(n * 8 + 7) / 8, 256, 0, st>>>(h->qkv, att, n);
    FF_LAUNCH_CHECK(h, "attention2_f32");
    if ((rc = lin(att, X.out, rows, h->x, 0, 1, "out_fp32"))) return rc;
    layernorm_f32_kernel<<<(rows + 7) / 8, 256, 0, st>>>(h->x, X.ln2_g, X.ln2_b, xn, rows);
    FF_LAUNCH_CHECK(h, "layernorm_f32");
    if ((rc = lin(xn, X.ff1, rows, ffh, 2, 0, "ff1_fp32"))) return rc;
    if ((rc = lin(ffh, X.ff2, rows, h->x, 0, 1, "ff2_fp32"))) return rc;
    if (tap_hit(19 + l, h->x, (int64_t)rows * DIM)) return FF_OK;
  }
  cls_gather_f32_kernel<<<n, 256, 0, st>>>(h->x, xn, n);
  FF_LAUNCH_CHECK(h, "cls_gather_f32");
  if ((rc = lin(xn, h->head1, n, h->hid, 1, 0, "head1_fp32"))) return rc;
  head2_kernel<<<(n + 7) / 8, 256, 0, st>>>(h->hid, h->head2.wf, h->head2.b, logits, n);
  FF_LAUNCH_CHECK(h, "head2");
  if (tap_hit(25, logits, (int64_t)n * 2)) return FF_OK;
  return FF_OK;
}

int forward_all(ff_cvit* h, const void* x, int layout, const int32_t* slot, int n, float* logits, cudaStream_t st,
                DebugTap* tap) {
  if (!h->finalized) return fail(h, FF_ERR_STATE, "weights not finalized");
  if (n < 0 || (n > 0 && (!x || !logits))) return fail(h, FF_ERR_BAD_ARG, "bad forward arguments");
  if (layout != FF_X_NCHW_F32 && layout != FF_X_NHWC_U8) return fail(h, FF_ERR_BAD_ARG, "unknown x_layout %d", layout);
  int dev = -1;
  FF_CUDA(h, cudaGetDevice(&dev));
  if (dev != h->device) FF_CUDA(h, cudaSetDevice(h->device));
  // the workspace is shared by all calls on this handle: order this call after the previous one (any stream)
  FF_CUDA(h, cudaStreamWaitEvent(st, h->done_ev, 0));
  const size_t crop_in_bytes = layout == FF_X_NHWC_U8 ? (size_t)224 * 224 * 3 : (size_t)224 * 224 * 3 * 4;
  int rc = FF_OK;
  for (int p0 = 0; p0 < n && rc == FF_OK; p0 += h->cap) {
    const int np = std::min(h->cap, n - p0);
    rc = forward_pass(h, reinterpret_cast<const uint8_t*>(x) + (size_t)p0 * crop_in_bytes, layout,
                      slot ? slot + p0 : nullptr, p0, np, logits + (size_t)2 * p0, st, tap);
    if (tap && tap->hit) break;
  }
  cudaEventRecord(h->done_ev, st);
  if (dev != h->device && dev >= 0) cudaSetDevice(dev);
  return rc;
}

__global__ void slots_from_offsets_kernel(const int* __restrict__ off, int n_videos, int* __restrict__ slot) {
  const int v = blockIdx.x;
  if (v >= n_videos) return;
  const int a = off[v], e = off[v + 1];
  for (int i = a + threadIdx.x; i < e; i += blockDim.x) slot[i] = (i - a) & 31;
}
// channel-blocked [n][2][hw][hw][32] bf16 -> NHWC fp32 [n][hw][hw][64] (debug tap of layers 4, 5)
__global__ void unblock_bf16_to_f32_kernel(const __nv_bfloat16* __restrict__ in, float* __restrict__ out, int n, int hw) {
  const size_t total = (size_t)n * hw * hw * 64;
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % 64);
  const size_t pix = i / 64;
  const size_t per = (size_t)hw * hw;
  const size_t img = pix / per, rem = pix % per;
  out[i] = __bfloat162float(in[((img * 2 + c / 32) * per + rem) * 32 + (c % 32)]);
}
__global__ void bf16_to_f32_kernel(const __nv_bfloat16* __restrict__ in, float* __restrict__ out, size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i < n) out[i] = __bfloat162float(in[i]);
}

}  // namespace

// ================================================================================================ C-ABI
extern "C" {

const char* ff_last_error(const ff_cvit_t* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

namespace {
int create_impl(ff_cvit_t** out, int device, int max_crops, int compute_dtype, int kind) {
  if (!out || max_crops <= 0) return fail(nullptr, FF_ERR_BAD_ARG, "ff_cvit_create: bad arguments");
  if (compute_dtype != FF_COMPUTE_BF16 && compute_dtype != FF_COMPUTE_FP32)
    return fail(nullptr, FF_ERR_BAD_ARG, "ff_cvit_create: unknown compute_dtype %d", compute_dtype);
  if (kind != 0 && compute_dtype != FF_COMPUTE_BF16)
    return fail(nullptr, FF_ERR_BAD_ARG, "only FF_COMPUTE_BF16 is implemented for the ResVitKan / GGCA variants");
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev <= 0)
    return fail(nullptr, FF_ERR_CUDA, "no CUDA device (%s): libfacfake has no CPU fallback", cudaGetErrorString(e));
  if (device < 0 || device >= ndev) return fail(nullptr, FF_ERR_BAD_ARG, "device %d out of range (%d devices)", device, ndev);
  cudaDeviceProp prop;
  if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess)
    return fail(nullptr, FF_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
  if (prop.major != 10)
    return fail(nullptr, FF_ERR_CUDA, "device %d is sm_%d%d; libfacfake is built for sm_100a (B200) only", device, prop.major, prop.minor);
  if ((e = cudaSetDevice(device)) != cudaSuccess) return fail(nullptr, FF_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
  if (!g_encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || !fn) return fail(nullptr, FF_ERR_CUDA, "cuTensorMapEncodeTiled not available: %s", cudaGetErrorString(e));
    g_encode = reinterpret_cast<EncodeTiledFn>(fn);
  }
  ff_cvit* h = new ff_cvit();
  h->device = device;
  h->compute = compute_dtype;
  h->kind = kind;
  h->ln2_eps = kind == 2 ? 1e-6f : 1e-5f;
  h->cap = (max_crops + 31) / 32 * 32;
  h->rows_cap = (2 * h->cap + 127) / 128 * 128;
  h->s12_cap = 256;
  h->s12 = std::min(64, h->cap);
  if (const char* v = getenv("FF_TC_VARIANT")) h->variant = atoi(v);
  if (const char* v = getenv("FF_WS")) h->use_ws = atoi(v);
  if (const char* v = getenv("FF_RVK_PERSIST")) h->rvk_persist = atoi(v);
  if (const char* v = getenv("FF_PDL")) g_use_pdl = atoi(v) != 0;
  if (const char* v = getenv("FF_DUAL")) h->use_dual = atoi(v);
  if (const char* v = getenv("FF_GEMM_BN")) h->gemm_bn_wide = atoi(v) == 128 ? 128 : 64;
  if (const char* v = getenv("FF_XF")) h->use_xf = atoi(v);
  if (const char* v = getenv("FF_PTC2")) h->use_ptc2 = atoi(v);
  if (const char* v = getenv("FF_PTC2_128")) h->use_ptc2m = atoi(v);
  if (const char* v = getenv("FF_PTC")) h->use_ptc = atoi(v);
  if (const char* v = getenv("FF_WS2")) h->use_ws2 = atoi(v);
  if (const char* v = getenv("FF_WS4")) h->use_ws4 = atoi(v);
  if (const char* v = getenv("FF_C12")) h->use_c12 = atoi(v);
  if (const char* v = getenv("FF_WS2X")) h->use_ws2x = atoi(v);
  if (const char* v = getenv("FF_C1_TC")) h->use_c1_tc = atoi(v);
  if (const char* v = getenv("FF_C1_CPS")) h->c1_ctas_per_sm = std::max(1, atoi(v));
  if (const char* v = getenv("FF_WS_CPS")) h->ws_ctas_per_sm = std::max(1, atoi(v));
  cudaDeviceGetAttribute(&h->num_sms, cudaDevAttrMultiProcessorCount, device);
  if (const char* v = getenv("FF_S12")) h->s12 = std::max(1, std::min(h->s12_cap, atoi(v)));
  int rc = FF_OK;
  const int cap128 = (h->cap + 127) / 128 * 128;
  do {
    if (cudaEventCreateWithFlags(&h->done_ev, cudaEventDisableTiming) != cudaSuccess) { rc = fail(h, FF_ERR_CUDA, "event create failed"); break; }
    if ((rc = dev_alloc(h, &h->emb, (size_t)EMBED_SPLITS * h->cap * DIM))) break;
    if ((rc = dev_alloc(h, &h->x, (size_t)h->rows_cap * DIM))) break;
    if ((rc = dev_alloc(h, &h->qkv, (size_t)h->rows_cap * 3 * DIM))) break;
    if ((rc = dev_alloc(h, &h->hid, (size_t)cap128 * MLP))) break;
    if (kind == 1) {
      for (int i = 0; i < 5 && rc == FF_OK; ++i) rc = dev_alloc(h, &h->rvk_buf[i], (size_t)h->cap * kRvkActElems);
      if (rc) break;
      if ((rc = dev_alloc(h, &h->rvk_x4, (size_t)h->cap * 224 * 224 * 4))) break;
      if ((rc = dev_alloc(h, &h->kan_part, (size_t)KAN_CHUNKS * h->cap * 64))) break;
    }
    if (kind == 2 && (rc = dev_alloc(h, &h->bufR, (size_t)h->cap * 56 * 56 * 128))) break;
    if (compute_dtype == FF_COMPUTE_BF16 && kind != 1) {
      if ((rc = dev_alloc(h, &h->bufA, (size_t)h->s12_cap * 224 * 224 * 32))) break;
      if ((rc = dev_alloc(h, &h->bufB, (size_t)h->s12_cap * 224 * 224 * 32))) break;
      if ((rc = dev_alloc(h, &h->bufA2, (size_t)h->s12_cap * 224 * 224 * 32))) break;
      if ((rc = dev_alloc(h, &h->bufB2, (size_t)h->s12_cap * 224 * 224 * 32))) break;
      if (cudaStreamCreateWithFlags(&h->aux_stream, cudaStreamNonBlocking) != cudaSuccess ||
          cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
          cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming) != cudaSuccess) { rc = fail(h, FF_ERR_CUDA, "aux stream/event create failed"); break; }
      if ((rc = dev_alloc(h, &h->P, (size_t)h->cap * 56 * 56 * 128))) break;
      if ((rc = dev_alloc(h, &h->Q, (size_t)h->cap * 56 * 56 * 128))) break;
    }
    if (compute_dtype == FF_COMPUTE_BF16) {
      if ((rc = dev_alloc(h, &h->feat, (size_t)cap128 * PATCH))) break;
      if ((rc = dev_alloc(h, &h->xn, (size_t)h->rows_cap * DIM))) break;
      if ((rc = dev_alloc(h, &h->att, (size_t)h->rows_cap * DIM))) break;
      if ((rc = dev_alloc(h, &h->ffh, (size_t)h->rows_cap * MLP))) break;
      if ((rc = dev_alloc(h, &h->qkvb, (size_t)h->rows_cap * 3 * DIM))) break;
      if ((rc = dev_alloc(h, &h->clsb, (size_t)cap128 * DIM))) break;
      // rows beyond the valid ones are read by TMA (results masked): keep them finite
      cudaMemset(h->feat, 0, (size_t)cap128 * PATCH * 2);
      cudaMemset(h->xn, 0, (size_t)h->rows_cap * DIM * 2);
      cudaMemset(h->att, 0, (size_t)h->rows_cap * DIM * 2);
      cudaMemset(h->ffh, 0, (size_t)h->rows_cap * MLP * 2);
      cudaMemset(h->clsb, 0, (size_t)cap128 * DIM * 2);
    } else {
      const size_t act = (size_t)h->s12_cap * 224 * 224 * 32;
      const size_t tail = std::max((size_t)h->rows_cap * DIM * 4, (size_t)1);
      if ((rc = dev_alloc(h, &h->fA, std::max(act, tail)))) break;
      if ((rc = dev_alloc(h, &h->fB, act + (size_t)h->cap * PATCH))) break;
    }
  } while (0);
  if (rc != FF_OK) {
    g_create_error = h->err;
    ff_cvit_destroy(h);
    return rc;
  }
  *out = h;
  return FF_OK;
}
}  // namespace

int ff_cvit_create(ff_cvit_t** out, int device, int max_crops, int compute_dtype) {
  return create_impl(out, device, max_crops, compute_dtype, 0);
}
int ff_resvitkan_create(ff_cvit_t** out, int device, int max_crops) {
  return create_impl(out, device, max_crops, FF_COMPUTE_BF16, 1);
}
int ff_cvit_ggca_create(ff_cvit_t** out, int device, int max_crops) {
  return create_impl(out, device, max_crops, FF_COMPUTE_BF16, 2);
}

void ff_cvit_destroy(ff_cvit_t* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  cudaDeviceSynchronize();
  for (void* p : h->allocs) cudaFree(p);
  if (h->slot_buf) cudaFree(h->slot_buf);
  if (h->xin_buf) cudaFree(h->xin_buf);
  if (h->logit_buf) cudaFree(h->logit_buf);
  if (h->off_buf) cudaFree(h->off_buf);
  if (h->score_buf) cudaFree(h->score_buf);
  if (h->crop_desc) cudaFree(h->crop_desc);
  if (h->done_ev) cudaEventDestroy(h->done_ev);
  for (cudaEvent_t e : h->ev_pool) cudaEventDestroy(e);
  for (cudaEvent_t e : h->h2d_ready) cudaEventDestroy(e);
  if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
  if (h->aux_stream) cudaStreamDestroy(h->aux_stream);
  if (h->ev_fork) cudaEventDestroy(h->ev_fork);
  if (h->ev_join) cudaEventDestroy(h->ev_join);
  delete h;
}

int ff_cvit_load_weight(ff_cvit_t* h, const char* key, const float* host_fp32, const int64_t* shape, int ndim) {
  if (!h) return FF_ERR_BAD_ARG;
  std::lock_guard<std::mutex> lk(h->mu);
  if (!key || (!host_fp32 && ndim > 0) || ndim < 0 || ndim > 4) return fail(h, FF_ERR_BAD_ARG, "ff_cvit_load_weight: bad arguments");
  if (h->finalized) return fail(h, FF_ERR_STATE, "weights already finalized");
  const std::string k(key);
  if (k.size() > 19 && k.compare(k.size() - 19, 19, "num_batches_tracked") == 0) return FF_OK;
  int64_t cnt = 1;
  std::vector<int64_t> shp;
  for (int i = 0; i < ndim; ++i) {
    if (shape[i] <= 0) return fail(h, FF_ERR_SHAPE, "weight '%s': non-positive dimension", key);
    cnt *= shape[i];
    shp.push_back(shape[i]);
  }
  if (cnt > (int64_t)1024 * 25088) return fail(h, FF_ERR_SHAPE, "weight '%s' too large for CViT", key);
  h->host_w[k].assign(host_fp32, host_fp32 + cnt);
  h->host_shape[k] = shp;
  return FF_OK;
}

int ff_cvit_finalize_weights(ff_cvit_t* h) {
  if (!h) return FF_ERR_BAD_ARG;
  std::lock_guard<std::mutex> lk(h->mu);
  if (h->finalized) return fail(h, FF_ERR_STATE, "weights already finalized");
  cudaSetDevice(h->device);
  return finalize(h);
}

int ff_cvit_forward(ff_cvit_t* h, const void* x, int x_layout, const int32_t* slot, int n, float* logits, void* stream) {
  if (!h) return FF_ERR_BAD_ARG;
  std::lock_guard<std::mutex> lk(h->mu);
  return forward_all(h, x, x_layout, slot, n, logits, reinterpret_cast<cudaStream_t>(stream), nullptr);
}

int ff_video_scores(ff_cvit_t* h, const float* logits, const int32_t* video_offsets, int n_videos, int mode, float* scores,
                    void* stream) {
  if (!h) return FF_ERR_BAD_ARG;
  std::lock_guard<std::mutex> lk(h->mu);
  if (n_videos < 0 || (n_videos > 0 && (!video_offsets || !scores))) return fail(h, FF_ERR_BAD_ARG, "ff_video_scores: bad arguments");
  if (mode != FF_REDUCE_REFERENCE && mode != FF_REDUCE_SOFTMAX_MEAN) return fail(h, FF_ERR_BAD_ARG, "unknown reduction mode %d", mode);
  if (n_videos == 0) return FF_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  video_reduce_kernel<<<(n_videos + 7) / 8, 256, 0, st>>>(logits, video_offsets, n_videos, mode, scores);
  FF_LAUNCH_CHECK(h, "video_reduce");
  return FF_OK;
}

int ff_cvit_predict(ff_cvit_t* h, const void* x, int x_layout, const int32_t* off_host, const int32_t* off_dev, int n_videos,
                    int mode, float* logits_out, float* scores, void* stream) {
  if (!h) return FF_ERR_BAD_ARG;
  std::lock_guard<std::mutex> lk(h->mu);
  if (n_videos < 0 || (n_videos > 0 && (!off_host || !off_dev || !scores))) return fail(h, FF_ERR_BAD_ARG, "ff_cvit_predict: bad arguments");
  if (mode != FF_REDUCE_REFERENCE && mode != FF_REDUCE_SOFTMAX_MEAN) return fail(h, FF_ERR_BAD_ARG, "unknown reduction mode %d", mode);
  if (n_videos == 0) return FF_OK;
  for (int v = 0; v < n_videos; ++v)
    if (off_host[v + 1] < off_host[v]) return fail(h, FF_ERR_BAD_ARG, "video_offsets must be non-decreasing");
  if (off_host[0] != 0) return fail(h, FF_ERR_BAD_ARG, "video_offsets[0] must be 0");
  const int n = off_host[n_videos];
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  cudaSetDevice(h->device);
  int rc;
  if ((rc = grow(h, &h->slot_buf, &h->slot_cap, (size_t)std::max(n, 1)))) return rc;
  float* lg = logits_out;
  if (!lg) {
    if ((rc = grow(h, &h->logit_buf, &h->logit_cap, (size_t)std::max(n, 1) * 2))) return rc;
    lg = h->logit_buf;
  }
  if (n > 0) {
    slots_from_offsets_kernel<<<n_videos, 64, 0, st>>>(off_dev, n_videos, h->slot_buf);
    FF_LAUNCH_CHECK(h, "slots_from_offsets");
    if ((rc = forward_all(h, x, x_layout, h->slot_buf, n, lg, st, nullptr))) return rc;
  }
  video_reduce_kernel<<<(n_videos + 7) / 8, 256, 0, st>>>(lg, off_dev, n_videos, mode, scores);
  FF_LAUNCH_CHECK(h, "video_reduce");
  return FF_OK;
}

int ff_cvit_predict_host(ff_cvit_t* h, const uint8_t* x_host, const int32_t* off_host, int n_videos, int mode, float* scores_host,
                         void* stream) {
  if (!h) return FF_ERR_BAD_ARG;
  if (n_videos < 0 || (n_videos > 0 && (!off_host || !scores_host))) return fail(h, FF_ERR_BAD_ARG, "ff_cvit_predict_host: bad arguments");
  if (n_videos == 0) return FF_OK;
  const int n = off_host[n_videos];
  if (n < 0 || (n > 0 && !x_host)) return fail(h, FF_ERR_BAD_ARG, "ff_cvit_predict_host: bad arguments");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  int rc;
  {
    std::lock_guard<std::mutex> lk(h->mu);
    cudaSetDevice(h->device);
    if ((rc = grow(h, &h->xin_buf, &h->xin_cap, (size_t)std::max(n, 1) * 224 * 224 * 3))) return rc;
    if ((rc = grow(h, &h->off_buf, &h->off_cap, (size_t)n_videos + 1))) return rc;
    if ((rc = grow(h, &h->score_buf, &h->score_cap, (size_t)n_videos))) return rc;
    FF_CUDA(h, cudaStreamWaitEvent(st, h->done_ev, 0));
    FF_CUDA(h, cudaMemcpyAsync(h->off_buf, off_host, ((size_t)n_videos + 1) * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    // crops stream in on a second stream, one stage-1/2 sub-pass at a time, overlapping the forward of earlier chunks
    if (!h->copy_stream) FF_CUDA(h, cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    h->h2d_chunk = std::min(h->s12, 32);
    const int chunks = (n + h->h2d_chunk - 1) / h->h2d_chunk;
    while ((int)h->h2d_ready.size() < chunks) {
      cudaEvent_t e;
      FF_CUDA(h, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      h->h2d_ready.push_back(e);
    }
    FF_CUDA(h, cudaStreamWaitEvent(h->copy_stream, h->done_ev, 0));   // previous call has finished reading xin_buf
    const size_t crop_bytes = (size_t)224 * 224 * 3;
    for (int c = 0; c < chunks; ++c) {
      const int c0 = c * h->h2d_chunk, cn = std::min(h->h2d_chunk, n - c0);
      FF_CUDA(h, cudaMemcpyAsync(h->xin_buf + c0 * crop_bytes, x_host + c0 * crop_bytes, cn * crop_bytes, cudaMemcpyHostToDevice,
                                 h->copy_stream));
      FF_CUDA(h, cudaEventRecord(h->h2d_ready[c], h->copy_stream));
    }
    h->h2d_chunks_pending = chunks;
  }
  rc = ff_cvit_predict(h, h->xin_buf, FF_X_NHWC_U8, off_host, h->off_buf, n_videos, mode, nullptr, h->score_buf, stream);
  {
    std::lock_guard<std::mutex> lk(h->mu);
    h->h2d_chunks_pending = 0;
  }
  if (rc) return rc;
  FF_CUDA(h, cudaMemcpyAsync(scores_host, h->score_buf, (size_t)n_videos * sizeof(float), cudaMemcpyDeviceToHost, st));
  FF_CUDA(h, cudaStreamSynchronize(st));
  return FF_OK;
}

int ff_preprocess_crops(ff_cvit_t* h, const uint8_t* const* crop_ptrs, const int32_t* hw, const int32_t* pitch, int n, int swap_rb,
                        uint8_t* out_u8, float* out_norm_nchw, void* stream) {
  if (!h) return FF_ERR_BAD_ARG;
  std::lock_guard<std::mutex> lk(h->mu);
  if (n < 0 || (n > 0 && (!crop_ptrs || !hw || !pitch || (!out_u8 && !out_norm_nchw)))) return fail(h, FF_ERR_BAD_ARG, "ff_preprocess_crops: bad arguments");
  if (n == 0) return FF_OK;
  std::vector<CropDesc> d(n);
  for (int i = 0; i < n; ++i) {
    const int ch = hw[2 * i], cw = hw[2 * i + 1];
    if (!crop_ptrs[i] || ch <= 0 || cw <= 0 || pitch[i] < cw * 3) return fail(h, FF_ERR_SHAPE, "crop %d: bad pointer/size/pitch", i);
    d[i].ptr = crop_ptrs[i]; d[i].h = ch; d[i].w = cw; d[i].pitch = pitch[i];
    const double sx = (double)cw / 224.0, sy = (double)ch / 224.0;
    if (sx >= 1.0 && sy >= 1.0) {
      const int isx = (int)std::lrint(sx), isy = (int)std::lrint(sy);
      const bool fast = std::fabs(sx - isx) < 2.220446049250313e-16 && std::fabs(sy - isy) < 2.220446049250313e-16;
      d[i].mode = fast ? PRE_FAST : PRE_FRAC; d[i].isx = isx; d[i].isy = isy;
    } else {
      d[i].mode = PRE_LINEAR; d[i].isx = d[i].isy = 1;
    }
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  cudaSetDevice(h->device);
  int rc;
  if ((rc = grow(h, &h->crop_desc, &h->crop_desc_cap, (size_t)n))) return rc;
  // pageable-host source: cudaMemcpyAsync stages the descriptors before returning, so `d` may die with this call;
  // the descriptor buffer is reused by the next call only after the stream-ordered kernel below has been enqueued
  FF_CUDA(h, cudaMemcpyAsync(h->crop_desc, d.data(), sizeof(CropDesc) * n, cudaMemcpyHostToDevice, st));
  {
    ProfScope ps(h, st, KC_SMALL);
    preprocess_kernel<<<dim3(224 / 4, n), 256, 0, st>>>(h->crop_desc, n, swap_rb, out_u8, out_norm_nchw);
  }
  FF_LAUNCH_CHECK(h, "preprocess");
  return FF_OK;
}

int64_t ff_cvit_launch_count(const ff_cvit_t* h) { return h ? h->launches : 0; }

int64_t ff_cvit_debug_activation(ff_cvit_t* h, const void* x, int x_layout, const int32_t* slot, int n, int stop_after,
                                 float* out_host, int64_t out_elems, void* stream) {
  if (!h) return FF_ERR_BAD_ARG;
  std::lock_guard<std::mutex> lk(h->mu);
  if (stop_after < 1 || stop_after > 27 || !out_host || n <= 0 || n > h->cap) return fail(h, FF_ERR_BAD_ARG, "ff_cvit_debug_activation: bad arguments");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  DebugTap tap;
  tap.stop_after = stop_after;
  float* lg = nullptr;
  if (cudaMalloc(&lg, (size_t)n * 2 * sizeof(float)) != cudaSuccess) return fail(h, FF_ERR_CUDA, "debug alloc failed");
  int rc = forward_all(h, x, x_layout, slot, n, lg, st, &tap);
  int64_t ret = rc;
  if (rc == FF_OK) {
    if (!tap.hit) ret = fail(h, FF_ERR_STATE, "debug tap %d not reached", stop_after);
    else if (tap.elems > out_elems) ret = fail(h, FF_ERR_BAD_ARG, "debug buffer too small: need %lld floats", (long long)tap.elems);
    else {
      cudaError_t e;
      if (tap.is_bf16) {
        float* tmp = nullptr;
        e = cudaMalloc(&tmp, (size_t)tap.elems * sizeof(float));
        if (e == cudaSuccess) {
          if (tap.blocked_hw)
            unblock_bf16_to_f32_kernel<<<(unsigned)((tap.elems + 255) / 256), 256, 0, st>>>(reinterpret_cast<const bf16*>(tap.ptr), tmp, n, tap.blocked_hw);
          else
            bf16_to_f32_kernel<<<(unsigned)((tap.elems + 255) / 256), 256, 0, st>>>(reinterpret_cast<const bf16*>(tap.ptr), tmp, (size_t)tap.elems);
          e = cudaMemcpyAsync(out_host, tmp, (size_t)tap.elems * sizeof(float), cudaMemcpyDeviceToHost, st);
          if (e == cudaSuccess) e = cudaStreamSynchronize(st);
          cudaFree(tmp);
        }
      } else {
        e = cudaMemcpyAsync(out_host, tap.ptr, (size_t)tap.elems * sizeof(float), cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
      }
      ret = (e == cudaSuccess) ? tap.elems : fail(h, FF_ERR_CUDA, "debug copy failed: %s", cudaGetErrorString(e));
    }
  }
  cudaStreamSynchronize(st);
  cudaFree(lg);
  return ret;
}

int ff_cvit_set_profiling(ff_cvit_t* h, int enable) {
  if (!h) return FF_ERR_BAD_ARG;
  std::lock_guard<std::mutex> lk(h->mu);
  h->profiling = enable != 0;
  h->prof_coarse = enable == 2;
  h->ev_used = 0;
  h->ev_class.clear();
  for (int i = 0; i < KC_COUNT; ++i) { h->prof_ms[i] = 0; h->prof_launches[i] = 0; }
  return FF_OK;
}

int ff_cvit_get_profile(ff_cvit_t* h, double* ms_by_class, int64_t* launches_by_class) {
  if (!h || !ms_by_class || !launches_by_class) return FF_ERR_BAD_ARG;
  std::lock_guard<std::mutex> lk(h->mu);
  cudaSetDevice(h->device);
  const size_t pairs = std::min(h->ev_class.size(), h->ev_used / 2);
  if (pairs > 0) FF_CUDA(h, cudaEventSynchronize(h->ev_pool[2 * pairs - 1]));
  for (size_t i = 0; i < pairs; ++i) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, h->ev_pool[2 * i], h->ev_pool[2 * i + 1]) == cudaSuccess) {
      h->prof_ms[h->ev_class[i]] += ms;
      h->prof_launches[h->ev_class[i]] += 1;
    }
  }
  h->ev_used = 0;
  h->ev_class.clear();
  for (int i = 0; i < KC_COUNT; ++i) { ms_by_class[i] = h->prof_ms[i]; launches_by_class[i] = h->prof_launches[i]; }
  return FF_OK;
}

int ff_cvit_set_tuning(ff_cvit_t* h, int stage12_sub_batch, int use_cuda_graph) {
  if (!h) return FF_ERR_BAD_ARG;
  std::lock_guard<std::mutex> lk(h->mu);
  (void)use_cuda_graph;
  if (stage12_sub_batch < 0 || stage12_sub_batch > h->s12_cap) return fail(h, FF_ERR_BAD_ARG, "stage12_sub_batch must be in [1,%d]", h->s12_cap);
  if (stage12_sub_batch > 0) h->s12 = std::min(stage12_sub_batch, h->cap);
  return FF_OK;
}

}  // extern "C"

// ================================================================================================ S3D
#include "ff_s3d.cuh"
