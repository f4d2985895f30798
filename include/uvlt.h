/*
 * uvlt.h -- C ABI of libuvlt_sm100.so, the B200 (sm_100a) implementation of UVLTrack's per-frame forward hot path.
 *
 * The reference (OpenSpaceAI/UVLTrack) is pure Python/PyTorch: it has no FFI of its own.  The entry points below
 * are what a reference-side binding for this path binds (ctypes stub shown in INTEGRATION.md); each one cites the
 * reference function it replaces (paths relative to the reference checkout).
 *
 * Conventions
 *   - plain C, no torch types: raw device pointers, sizes, an opaque handle and a cudaStream_t passed as void*.
 *   - every function returns 0 on success, non-zero on failure; uvlt_last_error() returns a thread-local message.
 *     Nothing throws across the ABI.  There is no CPU fallback: without a CUDA device every compute call fails.
 *   - all work is enqueued on the caller's stream; no hidden synchronisation unless documented.
 *   - caller owns every tensor it passes in; the library owns only its packed weights and workspace arenas.
 *   - a handle is bound to one device and is not thread-safe (one process per GPU, as the reference's
 *     lib/test/evaluation/running.py:97-100 does).
 */
#ifndef UVLT_H_
#define UVLT_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define UVLT_ABI_VERSION 1

#if defined(__GNUC__)
#define UVLT_API __attribute__((visibility("default")))
#else
#define UVLT_API
#endif

/* activation codes for the GEMM epilogue */
#define UVLT_ACT_NONE 0
#define UVLT_ACT_GELU 1 /* exact erf GELU: lib/models/backbones/utils.py:50, bert_backbone.py:118-124 */
#define UVLT_ACT_RELU 2 /* lib/models/heads/utils.py:126-130 */

UVLT_API int uvlt_abi_version(void);
UVLT_API const char* uvlt_last_error(void);

/* ------------------------------------------------------------------------------------------------------------
 * Engine: the whole per-frame path behind UVLTrack.forward_test (lib/models/uvltrack/uvltrack.py:41-45)
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct uvlt_engine* uvlt_handle;

typedef struct uvlt_config {
  int32_t embed_dim;       /* cfg.MODEL.HIDDEN_DIM: 768 (B) / 1024 (L) */
  int32_t num_heads;       /* 12 / 16 (head dim is 64 for both, mae_vit.py:218-229) */
  int32_t depth;           /* 12 / 24 transformer blocks */
  int32_t mlp_hidden;      /* 4 * embed_dim */
  int32_t template_size;   /* cfg.DATA.TEMPLATE.SIZE (pixels, multiple of 16) */
  int32_t search_size;     /* cfg.DATA.SEARCH.SIZE */
  int32_t text_len;        /* cfg.MODEL.BACKBONE.LANGUAGE.BERT.MAX_QUERY_LEN (40) */
  int32_t fusion_start;    /* min(cfg.MODEL.BACKBONE.FUSION_LAYER); layers >= this run jointly over image+text */
  int32_t head_channels;   /* cfg.MODEL.HEAD.HEAD_DIM (256) */
  int32_t vocab_size;      /* BERT vocabulary (30522) */
  int32_t max_position;    /* BERT max_position_embeddings (512) */
  int32_t max_batch;       /* largest number of sequences per call */
  int32_t softmax_one;     /* cfg.MODEL.HEAD.SOFTMAX_ONE */
  int32_t offset_sigmoid;  /* cfg.MODEL.HEAD.OFFSET_SIGMOID */
  int32_t txt_token_mean;  /* cfg.MODEL.BACKBONE.TXT_TOKEN_MODE == 'mean' (0 = 'cls') */
  int32_t num_cont_layers; /* len(cfg.MODEL.BACKBONE.CONT_LOSS_LAYER) */
  int32_t cont_layers[32]; /* cfg.MODEL.BACKBONE.CONT_LOSS_LAYER */
} uvlt_config;

/* Replaces build_model(cfg) (lib/models/uvltrack/uvltrack.py:47-57): allocates weight + workspace arenas on the
 * current CUDA device. */
UVLT_API int uvlt_create(const uvlt_config* cfg, uvlt_handle* out);
UVLT_API void uvlt_destroy(uvlt_handle h);

/* Replaces nn.Module.load_state_dict(strict=False) (lib/test/tracker/uvltrack.py:24).  `key` is the reference
 * state_dict key (e.g. "backbone.vit.blocks.3.attn.qkv.weight"); `data` is a HOST pointer to contiguous fp32.
 * Returns 0 when the key is consumed, 2 when it is not part of the hot path (ignored), 1 on a shape error. */
UVLT_API int uvlt_set_weight(uvlt_handle h, const char* key, const float* data, const int64_t* shape, int32_t ndim);
/* Repack: fp32 -> bf16 GEMM operands, fused BERT q/k/v, conv weights to (ky,kx,c) order with BatchNorm folded
 * (lib/models/heads/utils.py:126-130).  Fails if a required tensor was never set.  Synchronises the device. */
UVLT_API int uvlt_finalize_weights(uvlt_handle h);

/* options: "graph" (1: replay the layer chain as a CUDA graph, default 1), "bn" (force GEMM tile width, 0 = auto),
 * "pdl" (1: programmatic dependent launch between the kernels of the chain, default 1; process-wide),
 * "splitk" (1: split-K fc2 at small batch with the partials summed by the next LayerNorm, default 1) */
UVLT_API int uvlt_set_option(uvlt_handle h, const char* name, int32_t value);

typedef struct uvlt_outputs {
  /* all device pointers into the engine's arena, valid until the next forward on this handle; fp32 */
  const float* tokens;        /* [B, 1+Nz+Nx+T, D] final token stream: rows [cls | template | search | text] */
  const float* cls_score;     /* [B, S, S]   'cls_score' == 'cls_score_test' (JOINT_CLS false) */
  const float* bbox_map;      /* [B, S*S, 4] (cx, cy, w, h) relative to the search crop */
  const float* pred_boxes;    /* [B, 4]      bbox_map at argmax(cls * softmax(cont)[..., 0]) */
  const float* cont_score;    /* [B, S*S, cont_cols] */
  const float* cont_prob;     /* [B, S*S]    softmax(cont_score)[..., 0] */
  const float* logits;        /* [B, num_cont_layers, S, S] backbone contrastive logits (NULL when not requested) */
  const float* prompts;       /* [B, 3, D] the prompt the head used */
  int32_t batch, n_tokens, embed_dim, feat_size, cont_cols, reserved;
} uvlt_outputs;

/* flags for the forward calls */
#define UVLT_WANT_LOGITS 1 /* also evaluate the per-layer contrastive logits (training-only consumer) */
#define UVLT_SKIP_TEXT 2   /* every flag in the batch is 0 (BBOX): text keys are masked in every fusion layer
                              (modality_unified_feature_extractor.py:47), so the BERT branch and the text rows are
                              not evaluated; `tokens` text rows are then undefined.  Image-side results are identical. */

#define UVLT_FRAME_SLOT1 8 /* uvlt_track_frame_image_host: crop from frame-staging slot 1 (filled with
                              uvlt_upload_frames_slot) instead of slot 0 */

#define UVLT_NO_SYNC 16    /* uvlt_track_frame_image_host: enqueue the step (incl. the copy of the result rows to out_host,
                              which must then be page-locked) and return without synchronising the stream; the caller
                              waits with uvlt_step_wait (or uvlt_stream_sync) before it reads out_host.  Lets a caller that knows its next
                              frames launch step t+1 before it post-processes the rows of step t. */

#define UVLT_TEXT_CACHED 4 /* the text rows entering the first fusion layer were computed by uvlt_text_encode for these
                              sequences (ids / text_mask are constant per sequence): the BERT embedding and the
                              BERT-only layers are not re-run, their rows are restored from the cache.  Bit-identical. */

/* BertModel.embedding + the first min(FUSION_LAYER) BertLayers (bert_backbone.py:740-750, :383-394) for `batch`
 * sequences, once per sequence (Tracker.initialize); fills the engine's text cache used by UVLT_TEXT_CACHED
 * (SURVEY 8f row n4).  ids int64 [B,T], text_mask fp32 [B,T], flag int64 [B]: device pointers. */
UVLT_API int uvlt_text_encode(uvlt_handle h, const int64_t* ids, const float* text_mask, const int64_t* flag,
                              int32_t batch, void* stream);

/* UVLTrack.forward_test (lib/models/uvltrack/uvltrack.py:41-45).
 *   template [B,3,Hz,Hz] fp32, search [B,3,Hx,Hx] fp32, ids int64 [B,T], text_mask fp32 [B,T] (1 = real token),
 *   prompt fp32 [B,3,D], flag int64 [B] (0 BBOX, 1 NL, 2 NL+BBOX) -- all DEVICE pointers. */
UVLT_API int uvlt_forward_test(uvlt_handle h, const float* tmpl, const float* search, const int64_t* ids,
                               const float* text_mask, const float* prompt, const int64_t* flag, int32_t batch,
                               int32_t flags, uvlt_outputs* out, void* stream);

/* UVLTrack.forward (lib/models/uvltrack/uvltrack.py:18-24), the entry point Tracker.grounding() uses
 * (lib/test/tracker/uvltrack.py:45-62): backbone -> prompter on the batch's own features with the context taken
 * from sequence (b + B/2) % B (modality_adaptive_box_head.py:130-131) -> head with the two-column cont_score.
 *   template_mask uint8 [B,Nz], context_mask uint8 [B,Nx] (1 = inside the target box). */
UVLT_API int uvlt_forward_train(uvlt_handle h, const float* tmpl, const float* search, const int64_t* ids,
                                const float* text_mask, const int64_t* flag, const uint8_t* template_mask,
                                const uint8_t* context_mask, int32_t batch, int32_t flags, uvlt_outputs* out,
                                void* stream);

/* Backbone only (first half of forward_prompt_init, lib/models/uvltrack/uvltrack.py:26-27): fills out->tokens. */
UVLT_API int uvlt_backbone(uvlt_handle h, const float* tmpl, const float* search, const int64_t* ids,
                           const float* text_mask, const int64_t* flag, int32_t batch, int32_t flags,
                           uvlt_outputs* out, void* stream);

/* box_head.forward_prompt (modality_adaptive_box_head.py:96-106, heads/utils.py:78-99) on a token stream
 * [B, 1+Nz+Nx+T, D] (fp32 device; NULL = the stream of the last forward on this handle).
 * prompt_out fp32 [B,3,D] device. */
UVLT_API int uvlt_forward_prompt(uvlt_handle h, const float* tokens, const int64_t* flag, const float* text_mask,
                                 const uint8_t* template_mask, const uint8_t* context_mask, int32_t batch,
                                 float* prompt_out, void* stream);

/* ModalityAdaptiveBoxHead.forward (lib/models/heads/modality_adaptive_box_head.py:62-94), test branch (the prompt is given,
 * :140-148): conv towers, contrastive score against the prompt, convert2bbox.
 *   search_tokens fp32 [B, Nx, D] = backbone_info['search'] (NULL: the search rows of the token stream the last
 *   uvlt_backbone / forward on this handle left in the engine), prompt fp32 [B,3,D], flag int64 [B]: device pointers.
 * Fills out->cls_score / bbox_map / pred_boxes / cont_score / cont_prob / prompts. */
UVLT_API int uvlt_head(uvlt_handle h, const float* search_tokens, const float* prompt, const int64_t* flag, int32_t batch,
                       uvlt_outputs* out, void* stream);

/* Post-processing of Tracker.track (lib/test/tracker/uvltrack.py:116-121,127-130) on the maps of the last forward:
 * merge = cls * window * softmax(cont)[...,0] in float64, argmax, gather.
 *   window: device float64 [S*S] (np.outer(np.hanning(S), np.hanning(S)), tracker :64-68)
 *   out: device fp32 [B,6] = cx, cy, w, h, score (= cls*cont at the argmax), argmax index.
 *   max_score (device fp32 [B], in/out, may be NULL) and snapshot (device fp32 [B, N, D], may be NULL): when
 *   has_cont and score > max_score[b] the token stream of sequence b is copied to snapshot[b] and max_score[b] is
 *   raised -- the `self.out_dict = out_dict` bookkeeping of tracker :127-130 kept on the device. */
UVLT_API int uvlt_track_decode(uvlt_handle h, const double* window, int32_t has_cont, float* max_score,
                               float* snapshot, float* out, void* stream);

/* One tracker step end to end from HOST memory (pinned recommended): copies the raw uint8 RGB search crops
 * [B,Hx,Hx,3] to the device, fuses Preprocessor_wo_mask (tracker_utils.py:25-29) into the patch embedding, runs
 * forward_test + uvlt_track_decode, copies the [B,6] result rows back to out_host and synchronises the stream.
 * tmpl (fp32 [B,3,Hz,Hz]), ids, text_mask, prompt, flag, window, max_score, snapshot are DEVICE pointers. */
UVLT_API int uvlt_track_frame_host(uvlt_handle h, const uint8_t* search_u8_host, const float* tmpl, const int64_t* ids,
                                   const float* text_mask, const float* prompt, const int64_t* flag,
                                   const double* window, int32_t batch, int32_t flags, int32_t has_cont,
                                   float* max_score, float* snapshot, float* out_host, void* stream);

/* One tracker step with NOTHING but the frame upload on the host: raw uint8 RGB frames
 * [B, frame_h, frame_w, 3] (HOST, pinned recommended; one size per call) are copied to the device, and there
 *   sample_target (lib/train/data/processing_utils.py:159-243): crop of side ceil(sqrt(w*h)*search_factor) around the
 *     box in `state`, zero padded, cv2.resize INTER_LINEAR fixed-point arithmetic reproduced bit for bit,
 *   Preprocessor_wo_mask + forward_test + the window merge as in uvlt_track_frame_host,
 *   pred_box * search_size / resize_factor, map_box_back, clip_box(margin 10) (lib/test/tracker/uvltrack.py:123-125,
 *     :167-173, lib/utils/box_ops.py:117-126) in the reference's precisions (fp32 tensor ops, then fp64)
 * run back to back.  `state` is DEVICE fp64 [B,4] (x, y, w, h in frame pixels), updated in place.
 * out_host: HOST fp64 [B,10] = new state (4), network box cx cy w h (4), score, argmax index (-1 when the crop side
 * is < 1 pixel: the reference raises "Too small bounding box.", the state is then left unchanged).
 * Synchronises the stream before returning (unless UVLT_NO_SYNC). */
UVLT_API int uvlt_track_frame_image_host(uvlt_handle h, const uint8_t* frames_host, int32_t frame_h, int32_t frame_w,
                                         double* state, double search_factor, const float* tmpl, const int64_t* ids,
                                         const float* text_mask, const float* prompt, const int64_t* flag,
                                         const double* window, int32_t batch, int32_t flags, int32_t has_cont,
                                         float* max_score, float* snapshot, double* out_host, void* stream);

/* cudaStreamSynchronize(stream): the wait that goes with UVLT_NO_SYNC. */
UVLT_API int uvlt_stream_sync(void* stream);

/* Waits for the most recent uvlt_track_frame_image_host step that used frame slot `slot` (0, or 1 with UVLT_FRAME_SLOT1) --
 * its result rows are then in out_host -- WITHOUT waiting for work enqueued after it: a caller that has committed to its
 * next frames enqueues step t+1 (other slot, other out_host) first and then waits for step t only. */
UVLT_API int uvlt_step_wait(uvlt_handle h, int32_t slot);

/* Asynchronous piecewise upload of the raw frames of the next uvlt_track_frame_image_host call (which is then given
 * frames_host == NULL): `nbytes` from `host` (pinned recommended) to byte offset `dst_offset` of the engine's
 * [B, frame_h, frame_w, 3] staging buffer of `total_bytes`.  Lets the caller overlap its host-side staging copies with
 * the DMA, piece by piece.  Enqueued on `stream`; no synchronisation (except when the buffer has to grow). */
UVLT_API int uvlt_upload_frames(uvlt_handle h, const uint8_t* host, int64_t dst_offset, int64_t nbytes,
                                int64_t total_bytes, void* stream);

/* Double-buffered variant: the engine keeps TWO frame-staging buffers.  While a step crops from one slot the caller
 * uploads the next step's frames into the other one on a SECOND stream (and makes the step's stream wait on an event
 * recorded after the upload), then passes UVLT_FRAME_SLOT1 / no flag to uvlt_track_frame_image_host accordingly: the H2D
 * copy of step t+1 overlaps the forward of step t.  The three upload entry points only enqueue copies (and grow the
 * buffer under a mutex): they are the one part of the ABI that may be called from other host threads than the handle's
 * owner. */
UVLT_API int uvlt_upload_frames_slot(uvlt_handle h, const uint8_t* host, int64_t dst_offset, int64_t nbytes,
                                     int64_t total_bytes, int32_t slot, void* stream);

/* Same, for a sub-rectangle: `rows` rows of `width_bytes` from `host` (row pitch src_pitch) to byte offset dst_offset of
 * the staging buffer with row pitch dst_pitch (= frame_w * 3).  sample_target only reads the search window
 * (processing_utils.py:183-199: x1..x2, y1..y2 clipped to the frame), so a caller that knows the box state uploads just
 * that window of every frame; the rest of the staging buffer is never read by that step. */
UVLT_API int uvlt_upload_frames_2d(uvlt_handle h, const uint8_t* host, int64_t src_pitch, int64_t dst_offset,
                                   int64_t dst_pitch, int64_t width_bytes, int64_t rows, int64_t total_bytes, void* stream);

/* sample_target alone (device pointers): frames uint8 [B,H,W,3], state fp64 [B,4] -> crops uint8 [B,S,S,3] and
 * resize_factor fp64 [B] (0 when the crop side is < 1). */
UVLT_API int uvlt_op_crop_resize(const uint8_t* frames, int32_t frame_h, int32_t frame_w, const double* state,
                                 double factor, int32_t out_size, uint8_t* crops, double* resize_factor, int32_t batch,
                                 void* stream);

/* The state update of Tracker.track alone (lib/test/tracker/uvltrack.py:123-125: pred_box * search_size / resize_factor in
 * fp32, map_box_back :167-173 and clip_box(margin 10) lib/utils/box_ops.py:117-126 in fp64), device pointers:
 * net_boxes fp32 [B,4] (cx, cy, w, h of the network), resize_factor fp64 [B], state fp64 [B,4] updated in place. */
UVLT_API int uvlt_op_box_update(const float* net_boxes, const double* resize_factor, int32_t search_size, int32_t frame_h,
                                int32_t frame_w, double* state, int32_t batch, void* stream);

/* Tracker.anno2mask (lib/test/tracker/uvltrack.py:183-194): boxes fp32 [B,4] normalised (x, y, w, h) inside the crop ->
 * mask uint8 [B, size*size] (1 = cell centre inside the box, plus the cell under the box centre).  Device pointers. */
UVLT_API int uvlt_op_anno2mask(const float* boxes, int32_t size, uint8_t* mask, int32_t batch, void* stream);

/* grounding_resize (lib/train/data/processing_utils.py:60-141, image part; NL-mode first frame, lib/test/tracker/
 * uvltrack.py:45-62): frames uint8 [B, frame_h, frame_w, 3] -> out uint8 [B, out_size, out_size, 3]: the whole frame
 * resized with its aspect ratio kept (longer side = out_size, cv2.resize INTER_LINEAR 8-bit arithmetic, bit-exact) and
 * centred in a canvas of zeros.  Device pointers. */
UVLT_API int uvlt_op_grounding_resize(const uint8_t* frames, int32_t frame_h, int32_t frame_w, int32_t out_size,
                                      uint8_t* out, int32_t batch, void* stream);

/* Preprocessor_wo_mask.process (lib/test/tracker/tracker_utils.py:25-29): crops uint8 [B,S,S,3] -> out fp32 [B,3,S,S],
 * ((x / 255) - mean) / std.  Device pointers.  (Search crops never take this path: their normalisation is fused into the
 * patch embedding; this is for the template / context crops of Tracker.initialize.) */
UVLT_API int uvlt_op_normalize_u8(const uint8_t* crops, float* out, int32_t size, int32_t batch, void* stream);

/* number of kernels the last forward/track call launched (for bench.py's gpu_launches) */
UVLT_API int uvlt_last_launch_count(uvlt_handle h);

/* ------------------------------------------------------------------------------------------------------------
 * Operator-level entry points (device pointers; used by the per-operator parity tests)
 * ---------------------------------------------------------------------------------------------------------- */

/* Host-only query (no GPU touched): which kernel configuration the library picks for a [groups][M,K] x [N,K]^T GEMM.
 * split_k != 0: the caller can consume split-K partials (fp32 output summed by the next LayerNorm, i.e. fc2).
 * plan[0] = 1 when the persistent CTA-pair kernel (tcgen05 cta_group::2) runs it, plan[1] = tile width BN,
 * plan[2] = split-K factor, plan[3] = split-K factor if this were the box head's first conv GEMM.  The rules encode
 * B200 measurements (profiles/r01_gemm_2sm.md, tools/kernel_sweep.py); tests/test_host_logic.py pins them. */
UVLT_API int uvlt_gemm_plan(int M, int N, int K, int groups, int out_f32, int act, int split_k, int32_t* plan);

/* Host-only query: the process-wide kernel switches as the library resolved them from the environment (UVLT_PDL,
 * UVLT_MULTICAST, UVLT_GEMM_2SM, UVLT_ATTN_V, UVLT_ATTN_SPLIT, UVLT_ATTN_POLY).  out[0..5] = programmatic dependent
 * launch on, TMA-multicast GEMM variant on, CTA-pair GEMM mode (0 never / 1 by rule / 2 forced), attention generation
 * (0 = automatic), key-split attention on, softmax variant of the third-generation kernel.  The defaults are the
 * measured-best configuration; tests/test_host_logic.py pins them (a default that flips silently costs 6 % at batch 1). */
UVLT_API int uvlt_runtime_switches(int32_t* out6);

/* out[M,N] = act(A[M,K] @ W[N,K]^T + bias) + resid;  A, W bf16 row-major; bias fp32 [N] or NULL; resid fp32 [M,N]
 * or NULL (may alias out when out_f32); out bf16 or fp32.  nn.Linear semantics (block.py:49,59; utils.py:63-69).
 * bn: tile width 32/64/128/256, 0 = auto, 512 = the CTA-pair kernel (tcgen05 cta_group::2, 256 x 256 tile per pair of
 * CTAs; needs N % 256 == 0).  Requires K % 64 == 0, N % 32 == 0. */
UVLT_API int uvlt_op_gemm(const void* A, const void* W, const float* bias, const float* resid, void* out, int M, int N, int K,
                 int act, int out_f32, int bn, void* stream);

/* The engine's small-batch fc2 configuration: out[M,N] (fp32, may alias resid) = A @ W^T + bias + resid with K cut in
 * `splits` ranges (0 = the engine's own choice for this shape); split s > 0 writes its raw partial product to
 * partials + (s-1)*M*N and the consumer adds them (the engine's next LayerNorm does).  *splits_used (may be NULL)
 * receives the number of splits. */
UVLT_API int uvlt_op_gemm_splitk(const void* A, const void* W, const float* bias, const float* resid, float* out,
                                 float* partials, int M, int N, int K, int splits, int* splits_used, void* stream);

/* groups independent GEMMs (the four conv towers): A [G][M,K], W [G][N,K], bias [G][N],
 * out bf16 at out + g*out_gstride + row*out_ld. */
UVLT_API int uvlt_op_gemm_grouped(const void* A, const void* W, const float* bias, void* out, int groups, int M, int N, int K,
                         int act, long long out_ld, long long out_gstride, int bn, void* stream);

/* Fused MHA on packed qkv bf16 [B, n, 3*H*64] -> out bf16 [B, n, H*64]; key_bias fp32 [B, n] or NULL.
 * (block.py:47-58, bert_backbone.py:299-325).  v_t / n_pad are reserved (pass NULL / 0). */
UVLT_API int uvlt_op_attention(const void* qkv, const float* key_bias, void* out, int B, int n, int H, const void* v_t,
                      int n_pad, void* stream);

/* LayerNorm over `rows` rows per sequence of an fp32 token stream (see csrc/rowwise.cuh LnParams). */
UVLT_API int uvlt_op_layernorm(float* x, long long x_bstride, int x_row_off, int rows, const float* add0,
                      const float* add1, int split, int dst_mode, void* dst_bf16, const float* gamma,
                      const float* beta, float eps, int B, int D, void* stream);

/* mae_vit.py:92,99,203-214: patchify as im2col (bf16 [B*(Nz+Nx), 768]) + cls rows of the token stream. */
UVLT_API int uvlt_op_patch_im2col(const float* tmpl, const float* srch, const uint8_t* tmpl_u8, const uint8_t* srch_u8,
                         int B, int Hz, int Hx, void* out, const float* cls, float* x_stream, long long x_bstride,
                         int D, void* stream);

/* 3x3/pad-1 im2col on an SxS token grid, G channel groups of C channels -> bf16 [G][B*S*S][9*C] (ky,kx,c). */
UVLT_API int uvlt_op_im2col3x3(const void* src, int src_f32, long long src_bstride, long long src_row_off, long long src_ld,
                      int G, int C, int S, int B, void* dst, void* stream);

/* bert_backbone.py:260-274 */
UVLT_API int uvlt_op_bert_embed(const long long* ids, const float* word, const float* pos, const float* type0,
                       const float* gamma, const float* beta, float* dst_f32, long long dst_bstride, int dst_row_off,
                       void* dst_bf16, int B, int T, int D, int vocab, void* stream);

/* modality_unified_feature_extractor.py:43-50 + bert_backbone.py:746-748 as additive key biases. */
UVLT_API int uvlt_op_build_bias(const long long* flag, const float* text_mask, int B, int Nz, int Nx, int T, float* bias_vis,
                       float* bias_joint, float* bias_bert, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* UVLT_H_ */
