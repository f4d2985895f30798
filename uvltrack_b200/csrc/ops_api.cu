// Operator-level C-ABI entry points (declared in include/uvlt.h).  They exist so that every kernel of the hot path
// can be parity-tested in isolation against the oracle through the same boundary the engine uses.
#include <cstdlib>
#include "../../include/uvlt.h"
#include "host_utils.h"
#include "rowwise.cuh"
#include "head.cuh"

using namespace uvlt;

extern "C" {

const char* uvlt_last_error(void) { return get_error(); }

int uvlt_abi_version(void) { return UVLT_ABI_VERSION; }



int uvlt_op_gemm(const void* A, const void* W, const float* bias, const float* resid, void* out, int M, int N, int K,
                 int act, int out_f32, int bn, void* stream) {
  if (init_kernel_attributes()) return 1;
  GemmEpilogue ep{};
  ep.bias = bias;
  ep.resid = resid;
  ep.resid_ld = N;
  ep.act = act;
  ep.out = out;
  ep.out_f32 = out_f32;
  ep.out_ld = N;
  GemmLaunch g;
  if (gemm_prepare(&g, A, K, 0, W, K, 0, M, N, K, 1, bn, ep)) return 1;
  return gemm_launch(g, static_cast<cudaStream_t>(stream));
}

int uvlt_op_gemm_splitk(const void* A, const void* W, const float* bias, const float* resid, float* out, float* partials,
                        int M, int N, int K, int splits, int* splits_used, void* stream) {
  if (init_kernel_attributes()) return 1;
  if (splits == 0) splits = pick_splits(M, N, K);
  GemmEpilogue ep{};
  ep.bias = bias;
  ep.resid = resid;
  ep.resid_ld = N;
  ep.act = ACT_NONE;
  ep.out = out;
  ep.out_f32 = 1;
  ep.out_ld = N;
  ep.split_out = partials;
  ep.split_stride = static_cast<long long>(M) * N;
  GemmLaunch g;
  int bn = splits > 1 ? 64 : 0;
  if (const char* v = std::getenv("UVLT_SPLITK_BN")) bn = std::atoi(v);  // tools/kernel_sweep.py only
  if (gemm_prepare(&g, A, K, 0, W, K, 0, M, N, K, 1, bn, ep, splits)) return 1;
  if (gemm_launch(g, static_cast<cudaStream_t>(stream))) return 1;
  if (splits_used) *splits_used = splits;
  return 0;
}

int uvlt_gemm_plan(int M, int N, int K, int groups, int out_f32, int act, int split_k, int32_t* plan) {
  if (!plan || M <= 0 || N <= 0 || K <= 0 || groups <= 0) return 1;
  const int splits = split_k ? pick_splits(M, N, K) : 1;
  const bool two = use_2sm(M, N, K, groups, splits);
  plan[0] = two ? 1 : 0;
  plan[1] = two ? 256 : (splits > 1 ? 64 : pick_bn(M, N, groups, out_f32 != 0, act));
  plan[2] = splits;
  plan[3] = pick_head_splits(M, N, K);
  return 0;
}

int uvlt_runtime_switches(int32_t* out6) {
  if (!out6) return 1;
  runtime_switches(out6);
  return 0;
}

int uvlt_op_gemm_grouped(const void* A, const void* W, const float* bias, void* out, int groups, int M, int N, int K,
                         int act, long long out_ld, long long out_gstride, int bn, void* stream) {
  if (init_kernel_attributes()) return 1;
  GemmEpilogue ep{};
  ep.bias = bias;
  ep.bias_gstride = N;
  ep.act = act;
  ep.out = out;
  ep.out_f32 = 0;
  ep.out_ld = out_ld;
  ep.out_gstride = out_gstride;
  GemmLaunch g;
  if (gemm_prepare(&g, A, K, static_cast<long long>(M) * K, W, K, static_cast<long long>(N) * K, M, N, K, groups, bn,
                   ep))
    return 1;
  return gemm_launch(g, static_cast<cudaStream_t>(stream));
}

int uvlt_op_attention(const void* qkv, const float* key_bias, void* out, int B, int n, int H, const void* v_t,
                      int n_pad, void* stream) {
  if (init_kernel_attributes()) return 1;
  AttnLaunch a;
  if (attn_prepare(&a, qkv, B, n, H, key_bias, out, v_t, n_pad)) return 1;
  return attn_launch(a, static_cast<cudaStream_t>(stream));
}

int uvlt_op_layernorm(float* x, long long x_bstride, int x_row_off, int rows, const float* add0, const float* add1,
                      int split, int dst_mode, void* dst_bf16, const float* gamma, const float* beta, float eps, int B,
                      int D, void* stream) {
  LnParams p{};
  p.x = x; p.x_bstride = x_bstride; p.x_row_off = x_row_off; p.rows = rows;
  p.add0 = add0; p.add1 = add1; p.split = split; p.dst_mode = dst_mode;
  p.dst_bf16 = reinterpret_cast<__nv_bfloat16*>(dst_bf16);
  p.gamma = gamma; p.beta = beta; p.eps = eps;
  p.total_rows = B * rows;
  return launch_layernorm(p, D, static_cast<cudaStream_t>(stream));
}

int uvlt_op_patch_im2col(const float* tmpl, const float* srch, const uint8_t* tmpl_u8, const uint8_t* srch_u8, int B,
                         int Hz, int Hx, void* out, const float* cls, float* x_stream, long long x_bstride, int D,
                         void* stream) {
  PatchParams p{tmpl, tmpl_u8, srch, srch_u8, B, Hz, Hx, reinterpret_cast<__nv_bfloat16*>(out), cls, x_stream,
                x_bstride, D};
  return launch_patch_im2col(p, static_cast<cudaStream_t>(stream));
}

int uvlt_op_im2col3x3(const void* src, int src_f32, long long src_bstride, long long src_row_off, long long src_ld,
                      int G, int C, int S, int B, void* dst, void* stream) {
  Im2col3Params p{src, src_f32, src_bstride, src_row_off, src_ld, G, C, S, B, reinterpret_cast<__nv_bfloat16*>(dst)};
  if (launch_im2col3x3(p, static_cast<cudaStream_t>(stream))) { set_error("im2col3x3 launch failed"); return 1; }
  return 0;
}

int uvlt_op_bert_embed(const long long* ids, const float* word, const float* pos, const float* type0,
                       const float* gamma, const float* beta, float* dst_f32, long long dst_bstride, int dst_row_off,
                       void* dst_bf16, int B, int T, int D, int vocab, void* stream) {
  BertEmbedParams p{ids, word, pos, type0, gamma, beta, dst_f32, dst_bstride, dst_row_off,
                    reinterpret_cast<__nv_bfloat16*>(dst_bf16), T, B * T, vocab};
  return launch_bert_embed(p, D, static_cast<cudaStream_t>(stream));
}

int uvlt_op_build_bias(const long long* flag, const float* text_mask, int B, int Nz, int Nx, int T, float* bias_vis,
                       float* bias_joint, float* bias_bert, void* stream) {
  BiasParams p{flag, text_mask, B, Nz, Nx, T, bias_vis, bias_joint, bias_bert, nullptr, nullptr, nullptr, nullptr, 0};
  if (launch_build_bias(p, static_cast<cudaStream_t>(stream))) { set_error("build_bias launch failed"); return 1; }
  return 0;
}

int uvlt_op_box_update(const float* net_boxes, const double* resize_factor, int32_t search_size, int32_t frame_h,
                       int32_t frame_w, double* state, int32_t B, void* stream) {
  if (!net_boxes || !resize_factor || !state || B < 1 || search_size < 1) { set_error("uvlt_op_box_update: bad argument"); return 1; }
  UVLT_LAUNCH(box_update_kernel, dim3((B + 63) / 64), dim3(64), 0, static_cast<cudaStream_t>(stream), net_boxes,
              resize_factor, search_size, frame_h, frame_w, state, B);
  if (cudaGetLastError() != cudaSuccess) { set_error("box_update launch failed"); return 1; }
  return 0;
}

int uvlt_op_anno2mask(const float* boxes, int32_t size, uint8_t* mask, int32_t B, void* stream) {
  if (!boxes || !mask || B < 1 || size < 1) { set_error("uvlt_op_anno2mask: bad argument"); return 1; }
  UVLT_LAUNCH(anno2mask_kernel, dim3((size * size + 127) / 128, B), dim3(128), 0, static_cast<cudaStream_t>(stream), boxes,
              size, mask, B);
  if (cudaGetLastError() != cudaSuccess) { set_error("anno2mask launch failed"); return 1; }
  return 0;
}

int uvlt_op_grounding_resize(const uint8_t* frames, int32_t frame_h, int32_t frame_w, int32_t out_size, uint8_t* out,
                             int32_t B, void* stream) {
  if (!frames || !out || B < 1 || frame_h < 2 || frame_w < 2 || out_size < 1 || out_size > 4096) {
    set_error("uvlt_op_grounding_resize: bad argument");
    return 1;
  }
  GroundParams gp{frames, frame_h, frame_w, out_size, out};
  UVLT_LAUNCH(grounding_resize_kernel, dim3((out_size * out_size + 255) / 256, B), dim3(256), 0,
              static_cast<cudaStream_t>(stream), gp);
  if (cudaGetLastError() != cudaSuccess) { set_error("grounding_resize launch failed"); return 1; }
  return 0;
}

int uvlt_op_normalize_u8(const uint8_t* crops, float* out, int32_t size, int32_t B, void* stream) {
  if (!crops || !out || B < 1 || size < 1) { set_error("uvlt_op_normalize_u8: bad argument"); return 1; }
  UVLT_LAUNCH(normalize_u8_kernel, dim3((size * size + 255) / 256, B), dim3(256), 0, static_cast<cudaStream_t>(stream),
              crops, out, size, B);
  if (cudaGetLastError() != cudaSuccess) { set_error("normalize_u8 launch failed"); return 1; }
  return 0;
}

}  // extern "C"
