/*
 * facfake.h — C-ABI of libfacfake.so: the B200 (sm_100a) engine for FAC_fake's CViT hot path.
 *
 * The reference (xiaomo9/FAC_fake) has no FFI of its own; its seam for this path is Python
 * (SURVEY.md §8b).  Each entry point below names the reference interface it replaces
 * (paths relative to /root/reference/CViT-main/).  Plain pointers and sizes only — no torch
 * types.  Unless stated otherwise every data pointer is a DEVICE pointer on the engine's
 * GPU and `stream` is a `cudaStream_t` passed as `void*` (NULL = legacy default stream).
 *
 * Every function returns FF_OK (0) or a negative FF_ERR_* code and never throws;
 * `ff_last_error()` gives the message of the most recent failure on that handle.
 * There is no CPU fallback: without a usable sm_100 device `ff_cvit_create` fails.
 */
#ifndef FACFAKE_H_
#define FACFAKE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ff_cvit ff_cvit_t;

enum {
  FF_OK = 0,
  FF_ERR_BAD_ARG = -1, /* NULL pointer, negative size, unknown enum value                     */
  FF_ERR_SHAPE = -2,   /* weight / input shape does not match the CViT configuration           */
  FF_ERR_CUDA = -3,    /* a CUDA runtime / driver call failed (message has the CUDA error)    */
  FF_ERR_STATE = -4    /* call order violated (e.g. forward before finalize, missing weights) */
};

/* input layouts accepted by ff_cvit_forward / ff_cvit_predict */
enum {
  FF_X_NCHW_F32 = 0, /* [n,3,224,224] fp32, already normalised — what `model(x)` receives at
                        cvit_prediction.py:229                                                  */
  FF_X_NHWC_U8 = 2   /* [n,224,224,3] uint8 crops exactly as stored by cvit_prediction.py:108-117;
                        the (x/255-mean)/std of cvit_prediction.py:41-45,214-215 is fused in    */
};

/* compute types */
enum {
  FF_COMPUTE_BF16 = 0, /* bf16 operands, fp32 accumulate (tcgen05), fp32 epilogues/residual     */
  FF_COMPUTE_FP32 = 1  /* fp32 CUDA-core path (parity to 1e-4; slow)                            */
};

/* per-video reduction modes for ff_video_scores / ff_cvit_predict */
enum {
  FF_REDUCE_REFERENCE = 0,    /* sigmoid per logit, mean, decision rule: cvit_prediction.py:258-281     */
  FF_REDUCE_SOFTMAX_MEAN = 1, /* mean over frames of softmax(logits)[fake] (extra, not the oracle)      */
  FF_REDUCE_REFERENCE_PROBS = 2 /* ff_video_scores only: the rows already went through pred_sig
                                   (cvit_prediction.py:258-259) — pre_process_prediction (:266-281) alone */
};

/* ---- lifetime ----------------------------------------------------------------------------
 * Replaces `CViT(image_size=224, patch_size=7, num_classes=2, channels=512, dim=1024, depth=6,
 * heads=8, mlp_dim=2048)` + `.to(device)` (cvit_prediction.py:62-64; model/cvit.py:80-165).
 * The engine is specialised for exactly that configuration.  `max_crops` is the number of
 * crops one internal pass holds (workspace is allocated here; none on the forward path);
 * larger batches are processed in several passes.                                          */
int ff_cvit_create(ff_cvit_t** out, int device, int max_crops, int compute_dtype);
/* Same handle type for the ResVitKan variant: replaces `CViT(...)` of
 * ResVitKan/ResVitKan.py:284-329 — ResNet-50 `features` (:185-240), the same patch embedding and
 * 6-layer ViT, and `kan_head` = Linear, Dropout, ReLU, KAN([2048,64,2]) (kan.py:90-206).  Weight keys are
 * that module's state_dict names (`features.layer{L}.{b}.conv{1,2,3}.weight`, `kan_head.3.layers.{0,1}.*`, …;
 * `mlp_head.*` and `num_batches_tracked` are accepted and ignored).  compute_dtype as for ff_cvit_create
 * (FF_COMPUTE_FP32 = CUDA-core path, logits within 1e-4, debug tap 25 only).  Every
 * other entry point below works unchanged on such a handle; ff_cvit_debug_activation steps are
 * 1 = stem + max-pool, 2..5 = layer1..layer4, 6 = channel conv + bn2, 18..25 as for CViT.            */
int ff_resvitkan_create(ff_cvit_t** out, int device, int max_crops, int compute_dtype);
/* The `cvit_GGCA_ADD_DEConv_RepBn8` variant: replaces `CViT(...)` of
 * model/cvit_GGCA_ADD_DEConv_RepBn8.py:353-455.  Weight keys are that module's state_dict names
 * (`features1.N.*`, `features2.N.*`, DEConv branches `….conv1_{1..4}.conv.*` / `….conv1_5.*`,
 * `ggca.shared_conv.*`, `transformer.layers.L.1.fn.norm.norm1.*`; RepBN / schedule buffers and the unused
 * `Deconv.*` are accepted and ignored).  Every DEConv is folded into one 3x3 kernel at finalize (:337-351),
 * LinearNorm is its eval() form LayerNorm(eps 1e-6) (:22-47), the gate is x * GGCA(x) (:143-213,447-448).
 * FF_COMPUTE_BF16: the conv stack runs on fp16 activations and filters (the difference filters need the extra
 * mantissa bits to hold the 2e-2 logit gate), the shared embedding / encoder / head on bf16.  FF_COMPUTE_FP32: CUDA-core
 * path, logits within 1e-4, debug tap 25 only.  Debug steps: 1..17 = the 17 pooled-plan conv layers (step 9 = features1.27), 26 = the extra
 * BN-less Conv2d(128,128) features1.26 (between steps 8 and 9), 27 = gated feature map, 18..25 as for CViT. */
int ff_cvit_ggca_create(ff_cvit_t** out, int device, int max_crops, int compute_dtype);
void ff_cvit_destroy(ff_cvit_t* h);
const char* ff_last_error(const ff_cvit_t* h); /* h may be NULL: last create() error */

/* ---- weights ------------------------------------------------------------------------------
 * Replaces `model.load_state_dict(checkpoint)` + `model.eval()` (cvit_prediction.py:66-70).
 * One call per state_dict entry, with the reference's key names, HOST fp32
 * data, C-contiguous.  `num_batches_tracked` entries are accepted and ignored.
 * `ff_cvit_finalize_weights` folds conv bias + eval BatchNorm (eps 1e-5) into per-channel
 * (scale, shift), converts to the kernels' layouts ([Cout][kh][kw][Cin] bf16, linears as
 * stored) and builds the TMA descriptors.  A missing key fails with FF_ERR_STATE, a key of the
 * wrong shape with FF_ERR_SHAPE.  `ff_cvit_unused_keys` returns, after finalize, the comma-
 * separated keys that were loaded but are not part of the model (what `load_state_dict(strict=
 * True)` reports as "unexpected keys"); the string lives as long as the handle.              */
int ff_cvit_load_weight(ff_cvit_t* h, const char* state_dict_key, const float* host_fp32,
                        const int64_t* shape, int ndim);
int ff_cvit_finalize_weights(ff_cvit_t* h);
const char* ff_cvit_unused_keys(const ff_cvit_t* h);

/* ---- crop preprocessing (K0) ----------------------------------------------------------------
 * Replaces `cv2.resize(face, (224,224), interpolation=cv2.INTER_AREA)` + `cv2.cvtColor(RGB2BGR)`
 * (cvit_prediction.py:114-115; same at :96-97, :141-142) for n variable-size crops.
 * crop_ptrs[i] (HOST array of DEVICE pointers) -> uint8 HWC crop i with hw[2i], hw[2i+1] =
 * (height, width) and pitch[i] bytes per row (HOST arrays).  out_u8 = [n,224,224,3] uint8.
 * If out_norm_nchw != NULL it additionally receives the normalised fp32 [n,3,224,224] tensor
 * of cvit_prediction.py:209-215.                                                               */
int ff_preprocess_crops(ff_cvit_t* h, const uint8_t* const* crop_ptrs, const int32_t* hw,
                        const int32_t* pitch, int n, int swap_rb, uint8_t* out_u8,
                        float* out_norm_nchw, void* stream);

/* ---- forward (K1..K7) ------------------------------------------------------------------------
 * Replaces `model(dfdc_tensor[a:b])` (cvit_prediction.py:229,234,238; model/cvit.py:167-179).
 * slot[i] in [0,32) is the index crop i would have inside the <=32 batch the reference passes
 * to forward() — it selects pos_embedding[slot] (model/cvit.py:175).  slot may be NULL
 * (= i % 32).  `slot` is a DEVICE pointer.  logits = [n,2] fp32 (fake, real).  Any n >= 0.    */
int ff_cvit_forward(ff_cvit_t* h, const void* x, int x_layout, const int32_t* slot, int n,
                    float* logits, void* stream);

/* ---- per-video reduction (K8) -----------------------------------------------------------------
 * Replaces `pre_process_prediction(pred_sig(y))` (cvit_prediction.py:240,258-281) for many videos:
 * video v owns logits rows [video_offsets[v], video_offsets[v+1]).  <= 2 frames -> 0.5.
 * Only the first 90 frames of a video count: the reference evaluates the chunks [0:32], [32:64],
 * [64:90] and drops the rest (cvit_prediction.py:224-238, `upper_bound=90`).
 * video_offsets is a DEVICE int32 [n_videos+1]; scores a DEVICE fp32 [n_videos].               */
int ff_video_scores(ff_cvit_t* h, const float* logits, const int32_t* video_offsets, int n_videos,
                    int mode, float* scores, void* stream);

/* ---- fused predict: forward + reduction --------------------------------------------------------
 * The model half of `predict()` (cvit_prediction.py:209-242) for n_videos at once.  Crops of
 * video v are rows [off[v], off[v+1]) of x; slot = (frame index within the video) % 32, i.e.
 * the reference's [0:32],[32:64],[64:90] chunking; frames >= 90 of a video do not enter its score
 * (cvit_prediction.py:235-238) — their rows of `logits_out` are still written.  HOST copy of the
 * offsets is required to size the pass; `video_offsets_host` and `video_offsets_dev` hold the same
 * n_videos+1 values.  Calls on one handle are serialised by a mutex (the reference drives
 * predict() from a ThreadPoolExecutor, cvit_prediction.py:73-83); every entry point runs on the
 * handle's device and restores the caller's current device before returning.                   */
int ff_cvit_predict(ff_cvit_t* h, const void* x, int x_layout, const int32_t* video_offsets_host,
                    const int32_t* video_offsets_dev, int n_videos, int mode, float* logits_out,
                    float* scores, void* stream);

/* ---- host-buffer convenience (what bench.py's e2e leg and a cgo/JNI caller would use) -----------
 * x_host: pinned or pageable HOST uint8 [n,224,224,3]; scores_host: HOST fp32 [n_videos].
 * Copies in, runs ff_cvit_predict, copies the scores back and synchronises the stream.          */
int ff_cvit_predict_host(ff_cvit_t* h, const uint8_t* x_host, const int32_t* video_offsets_host,
                         int n_videos, int mode, float* scores_host, void* stream);

/* ---- introspection -------------------------------------------------------------------------------
 * Kernel launches issued by this handle since creation (bench.py's `gpu_launches`).              */
int64_t ff_cvit_launch_count(const ff_cvit_t* h);
/* Debug/test hook: run the forward on n <= max_crops crops and stop after step `stop_after`
 * (1..17 = conv layer k incl. its pool; 18 = tokens; 19..24 = transformer layer; 25 = logits),
 * copying that step's activation to out (HOST, fp32, NHWC for conv layers, row-major otherwise).
 * out_elems is the capacity of `out` in floats; returns the element count written or <0.        */
int64_t ff_cvit_debug_activation(ff_cvit_t* h, const void* x, int x_layout, const int32_t* slot,
                                 int n, int stop_after, float* out_host, int64_t out_elems,
                                 void* stream);
/* Per-launch timing for bench.py's roofline: when enabled every kernel launch is bracketed by a CUDA event
 * pair on the launching stream.  ff_cvit_get_profile synchronises and returns accumulated milliseconds and
 * launch counts in 21 slots: 0 = stand-alone conv1 (fp32 input / debug tap; on the uint8 path layers 1+2 are one kernel
 * counted in slot 1), 1..16 = tcgen05 conv of feature layer 2..17,
 * 17 = patch-embedding GEMM, 18 = transformer GEMMs, 19 = head GEMM, 20 = small kernels.
 * enable == 2 selects a coarse mode that only times three phases per pass (slot 0 = feature layers 1-6,
 * slot 1 = feature layers 7-17, slot 2 = embedding + transformer + head) and leaves the launches PDL-chained.
 * ff_cvit_set_profiling resets the accumulators.                                                            */
int ff_cvit_set_profiling(ff_cvit_t* h, int enable);
int ff_cvit_get_profile(ff_cvit_t* h, double* ms_by_slot /*[21]*/, int64_t* launches_by_slot /*[21]*/);
/* Tunable (0 keeps the current value): crops per stage-1/2 sub-pass, 1..256.  There are no environment switches:
 * each layer has exactly one kernel.                                                           */
int ff_cvit_set_tuning(ff_cvit_t* h, int stage12_sub_batch);

/* ---- BlazeFace face detector -------------------------------------------------------------
 * Replaces the `facedet` object the reference builds in cvit_prediction.py:27-33 (`BlazeFace().to(device)`,
 * `load_weights`, `load_anchors`; helpers/blazeface.py) for its network + box decoding:
 * `BlazeFace.predict_on_batch(x, apply_nms=False)` (:182-223) = `_preprocess` (:162), `forward` (:109-148),
 * `_tensors_to_detections` up to the score mask (:236-303).  The data-dependent part — the `>= 0.75` mask and the
 * blending NMS (:225-234,305-358) — is run by the caller on the dense result, as the reference runs it on the host.
 *   load_weight   one call per entry of blazeface.pth's state_dict (reference key names, HOST fp32) plus the
 *                 key "anchors" with the [896,4] array of anchors.npy
 *   predict       tiles: DEVICE uint8 NHWC [n,128,128,3];  detections: DEVICE fp32 [n,896,17] =
 *                 (ymin, xmin, ymax, xmax, 6 x (kx, ky), sigmoid(clamp(score, +-100))) for every anchor;
 *                 raw_boxes [n,896,16] / raw_scores [n,896] (DEVICE, optional, may be NULL) are the network outputs.
 * fp32 CUDA-core path (the detector is 30 MFLOP per tile and its outputs are thresholded).                    */
typedef struct ff_blazeface ff_blazeface_t;
int ff_blazeface_create(ff_blazeface_t** out, int device, int max_tiles);
void ff_blazeface_destroy(ff_blazeface_t* h);
const char* ff_blazeface_last_error(const ff_blazeface_t* h); /* h may be NULL: last create() error */
int ff_blazeface_load_weight(ff_blazeface_t* h, const char* key, const float* host_fp32, const int64_t* shape, int ndim);
int ff_blazeface_finalize(ff_blazeface_t* h);
int ff_blazeface_predict(ff_blazeface_t* h, const uint8_t* tiles, int n, float* detections, float* raw_boxes,
                         float* raw_scores, void* stream);
/* Blending NMS of blazeface.py:305-358 on the device for `predict_on_batch(apply_nms=True)` (:223): per tile, the
 * detections with score >= min_score are merged greedily (IoU > iou_threshold, score-weighted mean, mean score).
 * detections: DEVICE [n,896,17] from ff_blazeface_predict;  faces: DEVICE [n,16,17];  counts: DEVICE int32 [n] —
 * number of faces of the tile, or -1 when a tile has more than 64 candidates / 16 faces (the caller then runs the
 * reference's host loop on that tile).                                                                      */
int ff_blazeface_nms(ff_blazeface_t* h, const float* detections, int n, float min_score, float iou_threshold, float* faces,
                     int32_t* counts, void* stream);
/* `BlazeFace.nms(detections)` (blazeface.py:225-234): the same blending NMS over n caller-supplied detection lists.
 * detections: DEVICE [offsets[n],17] (the lists back to back);  offsets: DEVICE int32 [n+1];  faces / counts as above
 * (count -1: more than 64 detections in the list or more than 16 faces; run the host loop for that list).        */
int ff_blazeface_nms_lists(ff_blazeface_t* h, const float* detections, const int32_t* offsets, int n, float iou_threshold,
                           float* faces, int32_t* counts, void* stream);
/* The tiling / untiling / cropping around the detector that the reference's FaceExtractor does on the host
 * (helpers/helpers_face_extract_1.py), on the device:
 *   tile_frames   `_tile_frames` (:139-205): frames = DEVICE uint8 [n_frames, H, W, 3]; every frame is cut into square
 *                 windows of side min(H, W) — three, (W - side) / 2 apart, when W > H, else one — and each is resized to
 *                 128 x 128 with cv2.INTER_AREA semantics (bit-exact).  tiles = DEVICE uint8 [n_frames * T, 128, 128, 3].
 *   frame_faces   `_resize_detections` + `_untile_detections` + `facedet.nms` + `_add_margin_to_detections` + the integer
 *                 rectangle of `_crop_faces` (:207-312) for every frame: detections = DEVICE [n_frames * T, 896, 17] from
 *                 ff_blazeface_predict on those tiles; faces = DEVICE fp32 [n_frames, 16, 17] in frame coordinates;
 *                 boxes = DEVICE int32 [n_frames, 16, 4] = (ymin, xmin, ymax, xmax) of the crop incl. the margin;
 *                 counts = DEVICE int32 [n_frames], -1 when a frame has more than 64 candidates / 16 faces.
 * The crops themselves are then views of the frame on the device (`frame[ymin:ymax, xmin:xmax]`, pitch = W * 3) handed to
 * ff_preprocess_crops: pixels never return to the host between the decoder's upload and the CViT scores.          */
int ff_blazeface_tile_frames(ff_blazeface_t* h, const uint8_t* frames, int n_frames, int frame_h, int frame_w, uint8_t* tiles,
                             void* stream);
int ff_blazeface_frame_faces(ff_blazeface_t* h, const float* detections, int n_frames, int frame_h, int frame_w, float min_score,
                             float iou_threshold, float margin, float* faces, int32_t* boxes, int32_t* counts, void* stream);
int64_t ff_blazeface_launch_count(const ff_blazeface_t* h);

/* ---- S3D clip classifier ------------------------------------------------------------------
 * Replaces `S3D(num_class, 'no')` + `load_state_dict` + `model(video_faces)` of
 * sx_exp_deepfakedetect-master/S3D/S3D-test.py:210-212,199-205,267-272 (S3D/model.py:6-342).  Weight keys are that
 * module's state_dict names (`base.N...`, `fc.0.*`, `SRM.hpf.weight`; `num_batches_tracked` is accepted and ignored).
 * `srm_net` = 1 builds the SRM front-end of `S3D(num_class, 'yes')` (model.py:11-16,38-39; SRM/HPF.py:11-37): the 30
 * high-pass residual filters `SRM.hpf.weight` [30,3,1,5,5] run before `base`, whose first convolution then takes 30
 * channels (`base.0.conv_s.weight` [64,30,1,7,7]); with `srm_net` = 0 `SRM.hpf.weight` is accepted and ignored, as the
 * reference module itself ignores it.  Clips are 224x224, `frames_per_clip` in 16..71.
 *   forward   x: DEVICE, x_layout FF_X_NCHW_F32 = fp32 [n,3,T,224,224] (the module's own input: raw 0..255 BGR,
 *             S3D-test.py:94-96) or FF_X_NHWC_U8 = uint8 [n,T,224,224,3] (frames as decoded);
 *             logits: DEVICE fp32 [n,num_class] (temporal mean of the per-window fc outputs, model.py:40-46).
 *   debug     activation after `base[base_index]` (0..15, model.py:17-34) as fp32 [n,T',H',W',C] on the HOST.
 * bf16 tensor-core path only.                                                                                   */
typedef struct ff_s3d ff_s3d_t;
int ff_s3d_create(ff_s3d_t** out, int device, int max_clips, int frames_per_clip, int num_class, int srm_net);
void ff_s3d_destroy(ff_s3d_t* h);
const char* ff_s3d_last_error(const ff_s3d_t* h); /* h may be NULL: last create() error */
int ff_s3d_load_weight(ff_s3d_t* h, const char* key, const float* host_fp32, const int64_t* shape, int ndim);
int ff_s3d_finalize(ff_s3d_t* h);
int ff_s3d_forward(ff_s3d_t* h, const void* x, int x_layout, int n, float* logits, void* stream);
int64_t ff_s3d_debug_activation(ff_s3d_t* h, const void* x, int x_layout, int n, int base_index, float* out_host,
                                int64_t out_elems, void* stream);
int64_t ff_s3d_launch_count(const ff_s3d_t* h);

#ifdef __cplusplus
}
#endif
#endif /* FACFAKE_H_ */
