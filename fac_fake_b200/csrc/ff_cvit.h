// ff_cvit.h — the engine handle behind `ff_cvit_t` and the host helpers its translation units share:
//   ff_cvit.cu       CViT weights / launch schedule / C-ABI   (reference: model/cvit.py:80-179, cvit_prediction.py:209-281)
//   ff_resvitkan.cu  ResNet-50 trunk + KAN head of the ResVitKan variant
//   ff_ggca.cu       DEConv folding, the BN-less conv and the GGCA gate of the GGCA variant
// Host-only declarations: kernels live in the .cuh files and are instantiated by the unit that launches them.
#pragma once
#include <cstdarg>
#include <cstdint>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/facfake.h"
#include "ff_host.h"

namespace ffe {

using ffh::bf16;

// ------------------------------------------------------------------------------------------------ model plan
struct ConvPlan { int cin, cout, hw; bool pool; int conv_idx; };
extern const ConvPlan kConv[17];
constexpr int DIM = 1024, DEPTH = 6, MLP = 2048, PATCH = 25088, SLOTS = 32;
constexpr float BN_EPS = 1e-5f;
constexpr int EMBED_SPLITS = 8;   // split-K of the 25088-deep patch embedding (392 k-blocks = 8 x 49)
constexpr int KAN_PART_CHUNKS = 32;   // split-K slabs of KAN layer 0 (== KAN_CHUNKS of ff_rvk.cuh)
constexpr int MAX_FRAMES = 90;    // cvit_prediction.py:235-238: chunks [0:32],[32:64],[64:90]; later frames are dropped

// folded BN scale / shift passed by value to the stage-1/2 kernels (constant-bank operands of the epilogue FMAs)
struct EpiParams {
  float scale[64];
  float shift[64];
};

struct ConvLayerDev {
  bf16* w = nullptr;        // [cout][3][3][cin] bf16
  float* wf = nullptr;      // fp32 path: [cout][3][3][cin]
  float* scale = nullptr;
  float* shift = nullptr;
  CUtensorMap tmA, tmB;     // per-tap implicit GEMM (layers 7..17): A boxes of 128 pixels x 64 ch, B boxes {64, bn}
  int bn = 128;
  int bw = 16, bh = 8, bi = 1;
  bool pair2 = false;       // CTA-pair kernel (ff_ptc2.cuh): half filter tile per CTA
  CUtensorMap tmA_row;      // Cout = 128 (ff_ptcw.cuh): boxes {64 ch, bw + 2, bh, bi} = the three taps of one filter row
  CUtensorMap tmB_half;     // box {64, 128}
  EpiParams epi;            // host copy of (scale, shift), cout <= 64
  bool ws2 = false;         // pixel-pair formulation (Cin = 32): N = 2*Cout
  bf16* w2 = nullptr;       // pair-expanded filter [2*Cout][384]
  CUtensorMap tmA_ws2, tmW_ws2, tmO_ws2;
  bool ws2x = false;        // pixel-pair formulation on a CTA pair (Cin = 64): N = 128, cta_group::2
  bf16* w2x = nullptr;      // pair-expanded filter [128][768]
  CUtensorMap tmA_ws2x, tmW_ws2x, tmO_ws2x;
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
  // encoder kernel (ff_xf.cuh): LayerNorm folded into the two linears that follow it.  W'[n][k] = W[n][k] * gamma[k] (bf16),
  // c1[n] = sum_k bf16(W'[n][k]), c2[n] = sum_k beta[k] * W[n][k] (+ bias[n])
  LinearDev qkv_f, ff1_f;
  float *c1q = nullptr, *c2q = nullptr, *c1f = nullptr, *c2f = nullptr;
};

// profile slots: 0 = conv1, 1..16 = tcgen05 conv layer (index li), 17 = embed GEMM, 18 = transformer GEMMs,
// 19 = head GEMM, 20 = small kernels
enum { KC_CONV1 = 0, KC_TC_CONV = 1, KC_GEMM_EMBED = 17, KC_GEMM_XF = 18, KC_GEMM_HEAD = 19, KC_SMALL = 20, KC_COUNT = 21 };

}  // namespace ffe

struct ff_cvit {
  int device = 0;
  int kind = 0;            // 0 = CViT, 1 = ResVitKan, 2 = GGCA/DEConv/RepBN variant
  int cap = 0;             // crops per pass (multiple of 32)
  int rows_cap = 0;        // token rows capacity (multiple of 128)
  int s12 = 64;            // crops per stage-1/2 sub-pass
  int s12_cap = 256;
  int compute = FF_COMPUTE_BF16;
  bool act_f16 = false;    // conv-stack activations and filters in fp16 instead of bf16 (kind 2: the DEConv difference
                           // filters pass rounding noise at full gain; 3 more mantissa bits restore the 2e-2 gate)
  int num_sms = 148;
  bool finalized = false;
  std::mutex mu;           // serialises every call on this handle (the workspace is shared)
  mutable std::string err;
  int64_t launches = 0;
  cudaEvent_t done_ev = nullptr;

  std::map<std::string, std::vector<float>> host_w;
  std::map<std::string, std::vector<int64_t>> host_shape;
  std::map<std::string, bool> host_used;    // keys consumed by finalize (strict loading reports the rest)
  std::string unused_keys;                  // filled by finalize: comma-separated keys nobody asked for

  // ---- CViT conv stack
  ffh::bf16* c1_w = nullptr;     // [32][64] bf16, k = kh*16 + kw*4 + cin (fp32-input conv1 kernel)
  ffh::bf16* c1_wp = nullptr;    // [3][64][16] pair-expanded conv1 filter (uint8 kernels)
  float c1_na[3] = {0, 0, 0}, c1_nb[3] = {0, 0, 0};   // normalisation as one FMA per channel (verified against the table)
  float c1_scale[32], c1_shift[32];
  ffe::ConvLayerDev conv[17];
  ffe::ConvLayerDev conv_alt[6];   // layers 1..6 with TMA descriptors over the second ping-pong set
  ffe::LinearDev embed, head1, head2;
  ffe::XfLayerDev xf[ffe::DEPTH];
  float *pos = nullptr, *cls = nullptr;

  // ---- workspace
  ffh::bf16 *bufA = nullptr, *bufB = nullptr;     // stage 1/2 ping-pong, s12_cap crops
  ffh::bf16 *bufA2 = nullptr, *bufB2 = nullptr;   // second ping-pong set: odd sub-passes run on aux_stream
  cudaStream_t aux_stream = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  ffh::bf16 *P = nullptr, *Q = nullptr;           // stage 3..5 ping-pong, cap crops
  ffh::bf16* feat = nullptr;                      // [cap_rows128][25088]
  float* emb = nullptr;                           // [EMBED_SPLITS][cap][1024]
  float* x = nullptr;                             // [rows_cap][1024] residual stream
  ffh::bf16* xn = nullptr;                        // [rows_cap][1024]
  float* qkv = nullptr;                           // [rows_cap][3072] (fp32 path)
  ffh::bf16* qkvb = nullptr;                      // [rows_cap][3072] (bf16 path)
  ffh::bf16* att = nullptr;                       // [rows_cap][1024]
  ffh::bf16* ffh_buf = nullptr;                   // [rows_cap][2048]
  ffh::bf16* clsb = nullptr;                      // [cap128][1024]
  float* hid = nullptr;                           // [cap128][2048]
  CUtensorMap tm_feat, tm_xn, tm_att, tm_ffh, tm_cls;
  // whole-encoder cooperative kernel (ff_xf.cuh): device copy of the tensor maps it indexes, readiness
  std::vector<CUtensorMap> xf_maps;  // host copy; passed by value in the kernel parameters
  unsigned int* xf_sync = nullptr;   // group-barrier counters
  float* xf_stats = nullptr;         // [2][rows_cap][32] (sum, centred sum of squares) of 32-column segments of the residual stream
  int xf_groups = 0;                 // co-resident groups of 16 CTAs
  bool xf_ready = false;             // false: the per-op GPU launches run the encoder (cooperative launch unavailable)
  float* fA = nullptr;               // fp32-path workspace
  float* fB = nullptr;
  float* featf = nullptr;            // [cap][25088] fp32 patch vectors (tail of fB)

  // ---- GGCA / DEConv / RepBN variant (kind == 2): the CViT plan + one BN-less linear conv + the gate
  ffh::bf16* bufR = nullptr;         // output of the extra Conv2d(128,128) (features1.26), read by layer 9
  float *ggca_w1 = nullptr, *ggca_b1 = nullptr, *ggca_w2 = nullptr, *ggca_b2 = nullptr;
  float ln2_eps = 1e-5f;             // eps of the MLP-branch LayerNorm (1e-6 for LinearNorm.norm1)

  // ---- ResVitKan (kind == 1): ResNet-50 features + the same ViT + KAN head
  struct RvkOp {
    int type = 0;            // 0 = 1x1 stride-1 conv run flat, 1 = implicit-GEMM conv (3x3 any stride, or 1x1 stride 2)
    int cin = 0, cout = 0, taps = 9, stride = 1, in_hw = 0, out_hw = 0;
    int act = 0;             // 1 = ReLU after BN
    int resid = -1;          // buffer index of the residual (conv3) or -1
    int in_buf = 0, out_buf = 0;
    int bn = 64, bw = 8, bh = 8, bi = 2;
    ffh::bf16* w = nullptr;
    float* wf = nullptr;            // FF_COMPUTE_FP32: [cout][taps][cin] fp32
    ffh::bf16* out_ptr = nullptr;   // explicit output (kind 2); otherwise rvk_buf[out_buf] / feat
    float *scale = nullptr, *shift = nullptr;
    CUtensorMap tmA, tmB;
    CUtensorMap tmO, tmR;    // TMA-store epilogue: output / residual tiles in the output geometry
    std::string name;
  };
  std::vector<RvkOp> rvk_ops;            // kind 2 keeps its single extra conv (features1.26) here
  int rvk_layer_end[4] = {0, 0, 0, 0};   // index of the last op of layer1..4 (debug taps)
  ffh::bf16* rvk_buf[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  ffh::bf16* rvk_x4 = nullptr;           // normalised bf16 NHWC4 input
  ffh::bf16* rvk_stem_w = nullptr;
  float rvk_stem_scale[64], rvk_stem_shift[64];
  CUtensorMap rvk_tm_x4;
  float *kan_w0 = nullptr, *kan_w1 = nullptr, *kan_g0 = nullptr, *kan_g1 = nullptr;
  float* kan_part = nullptr;             // layer-0 split-K slabs [KAN_CHUNKS][cap][64]
  // FF_COMPUTE_FP32 trunk: same buffer roles as rvk_buf, `rvk_f32_chunk` crops at a time
  float* rvk_f32[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  float* rvk_x4f = nullptr;              // normalised fp32 NHWC4 input
  float *rvk_stem_wf = nullptr, *rvk_stem_scale_d = nullptr, *rvk_stem_shift_d = nullptr;
  int rvk_f32_chunk = 0;

  // ---- grow-only scratch for predict()
  int32_t* slot_buf = nullptr; size_t slot_cap = 0;
  uint8_t* xin_buf = nullptr; size_t xin_cap = 0;
  float* logit_buf = nullptr; size_t logit_cap = 0;
  int32_t* off_buf = nullptr; size_t off_cap = 0;
  float* score_buf = nullptr; size_t score_cap = 0;
  void* crop_desc = nullptr; size_t crop_desc_cap = 0;
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
  double prof_ms[ffe::KC_COUNT] = {0};
  int64_t prof_launches[ffe::KC_COUNT] = {0};
};

namespace ffe {

int fail(const ff_cvit* h, int code, const char* fmt, ...);
void prof_mark(ff_cvit* h, cudaStream_t st, int cls, bool begin, bool coarse = false);
struct ProfScope {
  ff_cvit* h; cudaStream_t st; int cls;
  ProfScope(ff_cvit* h_, cudaStream_t st_, int cls_) : h(h_), st(st_), cls(cls_) { prof_mark(h, st, cls, true); }
  ~ProfScope() { prof_mark(h, st, cls, false); }
};

#define FF_CUDA(h, call)                                                                             \
  do {                                                                                               \
    cudaError_t e_ = (call);                                                                         \
    if (e_ != cudaSuccess) return ffe::fail(h, FF_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

#define FF_LAUNCH_CHECK(h, what)                                                                     \
  do {                                                                                               \
    cudaError_t e_ = cudaGetLastError();                                                             \
    if (e_ != cudaSuccess) return ffe::fail(h, FF_ERR_CUDA, "launch of %s failed: %s", what, cudaGetErrorString(e_)); \
    ++(h)->launches;                                                                                 \
  } while (0)

template <typename T>
int dev_alloc(ff_cvit* h, T** p, size_t count) {
  void* q = nullptr;
  FF_CUDA(h, cudaMalloc(&q, (count > 0 ? count : 1) * sizeof(T)));
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
// 16-bit activations / filters of this handle: bf16, or fp16 bit patterns when h->act_f16
std::vector<bf16> to_act16(const ff_cvit* h, const std::vector<float>& v);

// weight lookup by state_dict key with shape check; records the error (missing -> FF_ERR_STATE, shape -> FF_ERR_SHAPE)
const std::vector<float>* get_w(ff_cvit* h, const std::string& key, std::initializer_list<int64_t> shape);
int weight_rc(const ff_cvit* h);   // the code that goes with the last get_w failure
int upload_linear(ff_cvit* h, LinearDev* L, const std::string& name, int out_f, int in_f, bool bias, int bn);
int upload_vec(ff_cvit* h, float** p, const std::string& key, int64_t n);

int tmap_2d(ff_cvit* h, CUtensorMap* m, const void* base, uint64_t inner, uint64_t rows, uint32_t box_inner, uint32_t box_rows);
int tmap_4d(ff_cvit* h, CUtensorMap* m, const void* base, int C, int W, int H, int N, int boxC, int bw, int bh, int bi,
            int estride = 1);

const bf16* conv_input_buffer(const ff_cvit* h, int li, int set = 0);
bf16* conv_output_buffer(const ff_cvit* h, int li, int set = 0);

struct DebugTap {
  int blocked_hw = 0;       // != 0: activation is channel-blocked [n][2][hw][hw][32]
  int stop_after = 0;       // 0 = run everything
  const void* ptr = nullptr;
  int64_t elems = 0;
  bool is_16 = false;       // 16-bit activation (bf16, or fp16 when h->act_f16)
  bool hit = false;
};

// ---- ff_resvitkan.cu
int finalize_rvk_features(ff_cvit* h);
int rvk_features(ff_cvit* h, const void* x, int layout, int slot_base, int n, cudaStream_t st, DebugTap* tap);
int rvk_features_fp32(ff_cvit* h, const void* x, int layout, int n, float* featf, cudaStream_t st);
int rvk_launch_op(ff_cvit* h, const ff_cvit::RvkOp& op, int n, cudaStream_t st, int prof_cls);
int kan_head(ff_cvit* h, int n, float* logits, cudaStream_t st);
// ---- ff_ggca.cu
struct GgcaPlan { const char* seq; int conv_idx; bool de; int bn_idx; };
extern const GgcaPlan kGgcaPlan[18];
int ggca_conv_weights(ff_cvit* h, const GgcaPlan& gp, int cin, int cout, std::vector<float>* w_out, std::vector<float>* b_out);
int finalize_ggca_extras(ff_cvit* h);
int ggca_gate(ff_cvit* h, int n, cudaStream_t st);
int ggca_gate_fp32(ff_cvit* h, float* featf, int n, cudaStream_t st);

}  // namespace ffe
