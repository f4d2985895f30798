// Box-head tail kernels: final 1x1 convs + sigmoid, contrastive-alignment similarity,
// softmax-one, score map, box decode, and the Hanning-windowed argmax of Tracker.track().
// All HBM-bound / latency-bound: one warp per search token, warp-shuffle reductions.
#pragma once
#include "common.cuh"
#include "track.cuh"

namespace uvlt {

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

// ----------------------------------------------------------------------------------------------
// modality_adaptive_box_head.py:73-82 (last Conv2d(32->{1,2,2,2}, k=1) + sigmoid, size tower chosen by flag),
// :140-148 (cont_score = e^{logit_scale} * <x^, p^_k>, softmax-one), :108-119 (ctr = (grid + offset) / S, bbox_map)
// ----------------------------------------------------------------------------------------------
struct HeadFinalParams {
  const __nv_bfloat16* y4;  // [B*SS, 4*C4]  tower order: cls | offset | size(track) | size(grounding)
  int C4;                   // channels of the last 3x3 layer (HEAD_DIM / 8 = 32)
  const float* w5;          // [7, C4]  rows: cls, off_x, off_y, w_tr, h_tr, w_gr, h_gr
  const float* b5;          // [7]
  const float* x_stream;    // fp32 token stream [B, n_tok, D]
  long long x_bstride;      // n_tok * D
  int x_row_off;            // first search token row (1 + Nz)
  int D;
  const float* prompt;      // [B, 3, D]
  const long long* flag;    // [B]
  float logit_scale_exp;
  int softmax_one;
  int train_branch;         // 1: two-column cont_score of the training branch (modality_adaptive_box_head.py:132-137)
  int offset_sigmoid;
  int S, B;
  float* cls_map;           // [B, SS]
  float* bbox_map;          // [B, SS, 4]  (cx, cy, w, h) relative to the search crop
  float* cont_score;        // [B, SS, 3] (softmax_one) or [B, SS, 2]
  float* cont_prob;         // [B, SS]  softmax(cont_score)[..., 0]
};

static __global__ void __launch_bounds__(256) head_final_kernel(const HeadFinalParams p) {
  pdl_wait();
  pdl_trigger();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int SS = p.S * p.S;
  const int r = blockIdx.x * (blockDim.x >> 5) + warp;
  if (r >= p.B * SS) return;
  const int b = r / SS, t = r - b * SS;
  const long long f = p.flag[b];

  // ---- last 1x1 convs: 7 dot products of length C4 over the four tower outputs ----
  float acc[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const __nv_bfloat16* y = p.y4 + static_cast<long long>(r) * 4 * p.C4;
  for (int c = lane; c < p.C4; c += 32) {
    const float y_cls = __bfloat162float(y[c]);
    const float y_off = __bfloat162float(y[p.C4 + c]);
    const float y_tr = __bfloat162float(y[2 * p.C4 + c]);
    const float y_gr = __bfloat162float(y[3 * p.C4 + c]);
    acc[0] += y_cls * p.w5[c];
    acc[1] += y_off * p.w5[p.C4 + c];
    acc[2] += y_off * p.w5[2 * p.C4 + c];
    acc[3] += y_tr * p.w5[3 * p.C4 + c];
    acc[4] += y_tr * p.w5[4 * p.C4 + c];
    acc[5] += y_gr * p.w5[5 * p.C4 + c];
    acc[6] += y_gr * p.w5[6 * p.C4 + c];
  }
#pragma unroll
  for (int i = 0; i < 7; ++i) acc[i] = warp_sum(acc[i]) + p.b5[i];
  const float cls = sigmoidf_(acc[0]);
  const float off_x = p.offset_sigmoid ? sigmoidf_(acc[1]) : acc[1];
  const float off_y = p.offset_sigmoid ? sigmoidf_(acc[2]) : acc[2];
  const bool gr = (f == 1);  // size_map_group = [track, grounding, track][flag]
  const float bw = sigmoidf_(gr ? acc[5] : acc[3]);
  const float bh = sigmoidf_(gr ? acc[6] : acc[4]);

  // ---- contrastive alignment: L2-normalised search token vs the three prompts ----
  const float* x = p.x_stream + static_cast<long long>(b) * p.x_bstride + static_cast<long long>(p.x_row_off + t) * p.D;
  const float* pr = p.prompt + static_cast<long long>(b) * 3 * p.D;
  float xx = 0.f, d0 = 0.f, d1 = 0.f, d2 = 0.f, n0 = 0.f, n1 = 0.f, n2 = 0.f;
  for (int i = lane * 4; i < p.D; i += 128) {
    const float4 xv = *reinterpret_cast<const float4*>(x + i);
    const float4 a = __ldg(reinterpret_cast<const float4*>(pr + i));
    const float4 c = __ldg(reinterpret_cast<const float4*>(pr + p.D + i));
    const float4 e = __ldg(reinterpret_cast<const float4*>(pr + 2 * p.D + i));
    xx += xv.x * xv.x + xv.y * xv.y + xv.z * xv.z + xv.w * xv.w;
    d0 += xv.x * a.x + xv.y * a.y + xv.z * a.z + xv.w * a.w;
    d1 += xv.x * c.x + xv.y * c.y + xv.z * c.z + xv.w * c.w;
    d2 += xv.x * e.x + xv.y * e.y + xv.z * e.z + xv.w * e.w;
    n0 += a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w;
    n1 += c.x * c.x + c.y * c.y + c.z * c.z + c.w * c.w;
    n2 += e.x * e.x + e.y * e.y + e.z * e.z + e.w * e.w;
  }
  xx = warp_sum(xx); d0 = warp_sum(d0); d1 = warp_sum(d1); d2 = warp_sum(d2);
  n0 = warp_sum(n0); n1 = warp_sum(n1); n2 = warp_sum(n2);
  const float inv_x = 1.0f / fmaxf(sqrtf(xx), 1e-12f);  // F.normalize: v / max(||v||, eps)
  const float c0 = p.logit_scale_exp * d0 * inv_x / fmaxf(sqrtf(n0), 1e-12f);
  const float c1 = p.logit_scale_exp * d1 * inv_x / fmaxf(sqrtf(n1), 1e-12f);
  const float c2 = p.logit_scale_exp * d2 * inv_x / fmaxf(sqrtf(n2), 1e-12f);
  float s1, prob0;
  const bool three = p.softmax_one && !p.train_branch;
  if (three) {
    s1 = fmaxf(fmaxf(c1, c2), 0.0f);
    const float m = fmaxf(fmaxf(c0, s1), 0.0f);
    const float e0 = expf(c0 - m), e1 = expf(s1 - m), e2 = expf(0.0f - m);
    prob0 = e0 / (e0 + e1 + e2);
  } else {
    s1 = p.softmax_one ? fmaxf(fmaxf(c1, c2), 0.0f) : fmaxf(c1, c2);
    const float m = fmaxf(c0, s1);
    const float e0 = expf(c0 - m), e1 = expf(s1 - m);
    prob0 = e0 / (e0 + e1);
  }
  if (lane == 0) {
    p.cls_map[r] = cls;
    const int col = t % p.S, row = t / p.S;  // coodinate channel 0 = column index (cx), channel 1 = row index (cy)
    const float add = p.offset_sigmoid ? 0.0f : 0.5f;
    float4 bb;
    bb.x = (static_cast<float>(col) + add + off_x) / static_cast<float>(p.S);
    bb.y = (static_cast<float>(row) + add + off_y) / static_cast<float>(p.S);
    bb.z = bw;
    bb.w = bh;
    *reinterpret_cast<float4*>(p.bbox_map + static_cast<long long>(r) * 4) = bb;
    if (three) {
      float* cs = p.cont_score + static_cast<long long>(r) * 3;
      cs[0] = c0; cs[1] = s1; cs[2] = 0.0f;
    } else {
      float* cs = p.cont_score + static_cast<long long>(r) * 2;
      cs[0] = c0; cs[1] = s1;
    }
    p.cont_prob[r] = prob0;
  }
}

// ----------------------------------------------------------------------------------------------
// Per-sequence argmax + gather.  One CTA per sequence, first-max tie rule (torch.argmax).
//   mode 0 (convert2bbox, modality_adaptive_box_head.py:110-118): score = cls * prob0 in fp32 -> pred_boxes [B,4]
//   mode 1 (Tracker.track, lib/test/tracker/uvltrack.py:116-121): merge = cls * window * prob0 evaluated in float64
//          exactly as the reference does on the host (float32 maps promoted by the float64 numpy window);
//          result row = [cx, cy, w, h, score = cls*prob0 (fp32), argmax index]
// ----------------------------------------------------------------------------------------------
struct DecodeParams {
  const float* cls_map;    // [B, SS]
  const float* cont_prob;  // [B, SS] or nullptr (has_cont == false -> 1)
  const float* bbox_map;   // [B, SS, 4]
  const double* window;    // [SS] (mode 1)
  int SS, mode;
  float* out;              // mode 0: [B, 4]; mode 1: [B, 6]
  float* max_score;        // mode 1, optional [B]: raised when the new score beats it (tracker :127-130)
  int* snap_flag;          // mode 1, optional [B]: 1 when the token stream of this sequence must be snapshotted
  // mode 1, optional device-resident box state (track.cuh): when `state` is set the new box is computed here
  double* state;           // [B, 4] x, y, w, h in frame pixels (in/out)
  const double* rf;        // [B] resize factor of this frame's crop
  int frame_h, frame_w, search_size;
  double* out10;           // [B, 10] = new state (4), network box cx cy w h (4), score, argmax index
};

static __global__ void __launch_bounds__(256) decode_kernel(const DecodeParams p) {
  pdl_wait();
  pdl_trigger();
  __shared__ double s_val[8];
  __shared__ int s_idx[8];
  const int b = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double best = -1.0e300;
  int best_i = 0x7fffffff;
  for (int t = threadIdx.x; t < p.SS; t += blockDim.x) {
    const float c = p.cls_map[b * p.SS + t];
    const float q = p.cont_prob ? p.cont_prob[b * p.SS + t] : 1.0f;
    double v;
    if (p.mode == 0) v = static_cast<double>(c * q);
    else v = (static_cast<double>(c) * p.window[t]) * static_cast<double>(q);
    if (v > best) { best = v; best_i = t; }  // strided scan keeps the smallest index per thread on ties
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
    if (ov > best || (ov == best && oi < best_i)) { best = ov; best_i = oi; }
  }
  if (lane == 0) { s_val[warp] = best; s_idx[warp] = best_i; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (blockDim.x >> 5); ++w)
      if (s_val[w] > best || (s_val[w] == best && s_idx[w] < best_i)) { best = s_val[w]; best_i = s_idx[w]; }
    if (best_i == 0x7fffffff) best_i = 0;  // all-NaN guard
    const float4 bb = *reinterpret_cast<const float4*>(p.bbox_map + (static_cast<long long>(b) * p.SS + best_i) * 4);
    if (p.mode == 0) {
      *reinterpret_cast<float4*>(p.out + b * 4) = bb;
    } else {
      float* o = p.out + b * 6;
      o[0] = bb.x; o[1] = bb.y; o[2] = bb.z; o[3] = bb.w;
      const float q = p.cont_prob ? p.cont_prob[b * p.SS + best_i] : 1.0f;
      const float score = p.cls_map[b * p.SS + best_i] * q;
      o[4] = score;
      o[5] = static_cast<float>(best_i);
      if (p.snap_flag) {
        const bool better = p.max_score && score > p.max_score[b];
        p.snap_flag[b] = better ? 1 : 0;
        if (better) p.max_score[b] = score;
      }
      if (p.state) {
        double* st = p.state + 4 * b;
        double* o10 = p.out10 + 10 * b;
        const double rf = p.rf[b];
        if (rf > 0.0) box_update(bb, rf, p.search_size, p.frame_h, p.frame_w, st);
        o10[0] = st[0]; o10[1] = st[1]; o10[2] = st[2]; o10[3] = st[3];
        o10[4] = bb.x; o10[5] = bb.y; o10[6] = bb.z; o10[7] = bb.w;
        o10[8] = score;
        o10[9] = rf > 0.0 ? static_cast<double>(best_i) : -1.0;  // -1: "Too small bounding box." (crop side < 1)
      }
    }
  }
}

// ----------------------------------------------------------------------------------------------
// Backbone contrastive logits (modality_unified_feature_extractor.py:85-93), one cont-loss layer per launch:
//   logit[b, t] = e^{ls} * <x^_t, token^>  with token^ = vis / txt, and their average of the two logits for flag 2.
// ----------------------------------------------------------------------------------------------
struct BackboneLogitParams {
  const float* img;        // token stream holding [cls | z | x] rows
  long long img_bstride;   // elements between batch elements
  const float* txt;        // text rows (txt token = first text row, TXT_TOKEN_MODE 'cls')
  long long txt_bstride;
  const float* text_mask;  // [B, T] (only for TXT_TOKEN_MODE 'mean')
  int txt_mean, T;
  int Nz, Nx, D, B;
  const long long* flag;
  float logit_scale_exp;
  float* out;              // [B, n_layers, Nx]; this launch writes slice `layer_slot`
  int n_layers, layer_slot;
};

static __global__ void __launch_bounds__(256) backbone_logits_kernel(const BackboneLogitParams p) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ float s_tok[];  // [2][D] : vis token, txt token of this batch element
  const int b = blockIdx.y;
  const float* vis = p.img + static_cast<long long>(b) * p.img_bstride;
  const float* txt = p.txt + static_cast<long long>(b) * p.txt_bstride;
  for (int i = threadIdx.x; i < p.D; i += blockDim.x) {
    s_tok[i] = vis[i];
    float tv;
    if (p.txt_mean) {
      float num = 0.f, den = 0.f;
      for (int j = 0; j < p.T; ++j) {
        const float m = p.text_mask[b * p.T + j];
        num += txt[static_cast<long long>(j) * p.D + i] * m;
        den += m;
      }
      tv = num / den;
    } else {
      tv = txt[i];
    }
    s_tok[p.D + i] = tv;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t = blockIdx.x * (blockDim.x >> 5) + warp;
  if (t >= p.Nx) return;
  const float* x = vis + static_cast<long long>(1 + p.Nz + t) * p.D;
  float xx = 0.f, dv = 0.f, dt = 0.f, nv = 0.f, nt = 0.f;
  for (int i = lane * 4; i < p.D; i += 128) {
    const float4 xv = *reinterpret_cast<const float4*>(x + i);
    const float4 a = *reinterpret_cast<const float4*>(s_tok + i);
    const float4 c = *reinterpret_cast<const float4*>(s_tok + p.D + i);
    xx += xv.x * xv.x + xv.y * xv.y + xv.z * xv.z + xv.w * xv.w;
    dv += xv.x * a.x + xv.y * a.y + xv.z * a.z + xv.w * a.w;
    dt += xv.x * c.x + xv.y * c.y + xv.z * c.z + xv.w * c.w;
    nv += a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w;
    nt += c.x * c.x + c.y * c.y + c.z * c.z + c.w * c.w;
  }
  xx = warp_sum(xx); dv = warp_sum(dv); dt = warp_sum(dt); nv = warp_sum(nv); nt = warp_sum(nt);
  if (lane == 0) {
    const float inv_x = 1.0f / fmaxf(sqrtf(xx), 1e-12f);
    const float lv = p.logit_scale_exp * dv * inv_x / fmaxf(sqrtf(nv), 1e-12f);
    const float lt = p.logit_scale_exp * dt * inv_x / fmaxf(sqrtf(nt), 1e-12f);
    const long long f = p.flag[b];
    const float l = (f == 0) ? lv : (f == 1 ? lt : (lv + lt) * 0.5f);
    p.out[(static_cast<long long>(b) * p.n_layers + p.layer_slot) * p.Nx + t] = l;
  }
}

// ----------------------------------------------------------------------------------------------
// Conditional snapshot of the token stream (the `self.out_dict = out_dict` of lib/test/tracker/uvltrack.py:127-130)
// ----------------------------------------------------------------------------------------------
static __global__ void __launch_bounds__(256) snapshot_kernel(const float* __restrict__ x, float* __restrict__ snap,
                                                       const int* __restrict__ snap_flag, long long per_seq4) {
  pdl_wait();
  pdl_trigger();
  const int b = blockIdx.y;
  if (!snap_flag[b]) return;
  const float4* src = reinterpret_cast<const float4*>(x) + b * per_seq4;
  float4* dst = reinterpret_cast<float4*>(snap) + b * per_seq4;
  for (long long i = blockIdx.x * blockDim.x + threadIdx.x; i < per_seq4; i += static_cast<long long>(gridDim.x) * blockDim.x)
    dst[i] = src[i];
}

// ----------------------------------------------------------------------------------------------
// Prompter, pooling half (heads/utils.py:45-99, DistributionBasedCrossAttention.forward without its MLP):
// one CTA per sequence.
//   token     = [vis, txt, (vis+txt)/2][flag]                      (modality_adaptive_box_head.py:97-102)
//   tgt       = [template rows | context rows]; sim = e^{ls} <token^, tgt^_j>
//   tgt_score = softmax(sim | tgt_mask), bgd_score = softmax(sim | ~tgt_mask)
//   threshold = first ascending-sorted bgd_score whose running sum reaches 0.25 (1.0 if none) -> distractor mask
//   src       = [tgt_token, dis_token, bgd_token] + src_,  src_ = query_embed (+ token on slot 0)
// ----------------------------------------------------------------------------------------------
struct PrompterParams {
  const float* tokens;          // [B, N, D]
  long long bstride;
  int Nz, Nx, Nv, T, D, B;
  int ctx_rot;                  // context rows come from sequence (b + ctx_rot) % B
  const uint8_t* template_mask; // [B, Nz]
  const uint8_t* context_mask;  // [B, Nx]
  const long long* flag;        // [B]
  int txt_mean;
  const float* text_mask;       // [B, T] (TXT_TOKEN_MODE 'mean')
  float logit_scale_exp;
  const float* query_embed;     // [3, D]
  float* src;                   // [B, 3, D]
  float* src0;                  // [B, 3, D]
  __nv_bfloat16* src_bf16;      // [3B, D]
};

__device__ __forceinline__ float block_reduce(float v, bool is_max, float* red) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  v = is_max ? warp_max(v) : warp_sum(v);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float r = red[0];
  for (int w = 1; w < (blockDim.x >> 5); ++w) r = is_max ? fmaxf(r, red[w]) : r + red[w];
  return r;
}

// masked softmax over sim[0..n): logits = keep[j] ? sim[j] : -1e20, written to out[0..n)
__device__ __forceinline__ void masked_softmax(const float* sim, int n, float* out, float* red, int mode,
                                               const uint8_t* m0, const uint8_t* m1) {
  // mode 0: keep = m0 (target)         mode 1: keep = !m0 (background)
  // mode 2: keep = !m0 && !m1 (pure background)    mode 3: keep = !m0-filled logit && m1 ... see below
  float mx = -INFINITY;
  for (int j = threadIdx.x; j < n; j += blockDim.x) {
    float l = sim[j];
    if (mode == 0) l = m0[j] ? l : -1e20f;
    else {
      l = m0[j] ? -1e20f : l;                 // bgd_logit
      if (mode == 2) l = m1[j] ? -1e20f : l;  // masked_fill(dis_mask)
      if (mode == 3) l = m1[j] ? l : -1e20f;  // masked_fill(~dis_mask)
    }
    out[j] = l;
    mx = fmaxf(mx, l);
  }
  mx = block_reduce(mx, true, red);
  float sum = 0.f;
  for (int j = threadIdx.x; j < n; j += blockDim.x) {
    const float e = expf(out[j] - mx);
    out[j] = e;
    sum += e;
  }
  sum = block_reduce(sum, false, red);
  for (int j = threadIdx.x; j < n; j += blockDim.x) out[j] = out[j] / sum;
  __syncthreads();
}

constexpr int PROMPTER_MAX_N = 2048;   // Nz + Nx <= 2048 (384^2 + 384^2 -> 1152)
constexpr int PROMPTER_SCRATCH = 3072; // floats: bitonic sort buffer (<= PROMPTER_MAX_N), then [3][D] accumulators (D <= 1024)

static __global__ void __launch_bounds__(256) prompter_pool_kernel(const PrompterParams p) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ float sm[];
  const int n = p.Nz + p.Nx;
  float* tok = sm;                       // [D]
  float* sim = tok + p.D;                // [n]
  float* w_tgt = sim + n;                // [n]
  float* w_bgd = w_tgt + n;              // [n]
  float* w_dis = w_bgd + n;              // [n]
  float* sorted = w_dis + n;               // [PROMPTER_SCRATCH]
  float* red = sorted + PROMPTER_SCRATCH;  // [8]
  uint8_t* tmask = reinterpret_cast<uint8_t*>(red + 8);  // [n]
  uint8_t* dmask = tmask + n;                              // [n]
  __shared__ float s_thr;

  const int b = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
  const long long f = p.flag[b];
  const float* xb = p.tokens + static_cast<long long>(b) * p.bstride;
  const float* xc = p.tokens + static_cast<long long>((b + p.ctx_rot) % p.B) * p.bstride;
  const float* vis = xb;
  const float* txt = xb + static_cast<long long>(p.Nv) * p.D;

  // ---- token ----
  float nrm = 0.f;
  for (int i = threadIdx.x; i < p.D; i += blockDim.x) {
    float tv;
    if (p.txt_mean) {
      float num = 0.f, den = 0.f;
      for (int j = 0; j < p.T; ++j) {
        const float m = p.text_mask[b * p.T + j];
        num += txt[static_cast<long long>(j) * p.D + i] * m;
        den += m;
      }
      tv = num / den;
    } else {
      tv = txt[i];
    }
    const float v = vis[i];
    const float t = (f == 0) ? v : (f == 1 ? tv : (v + tv) / 2.0f);
    tok[i] = t;
    nrm += t * t;
  }
  nrm = block_reduce(nrm, false, red);
  const float inv_tok = 1.0f / fmaxf(sqrtf(nrm), 1e-12f);
  for (int j = threadIdx.x; j < n; j += blockDim.x)
    tmask[j] = (j < p.Nz) ? p.template_mask[b * p.Nz + j] : p.context_mask[b * p.Nx + (j - p.Nz)];
  __syncthreads();

  // ---- similarity: one warp per target row ----
  for (int j = warp; j < n; j += nwarp) {
    const float* row = (j < p.Nz) ? xb + static_cast<long long>(1 + j) * p.D
                                  : xc + static_cast<long long>(1 + p.Nz + (j - p.Nz)) * p.D;
    float d = 0.f, rr = 0.f;
    for (int i = lane * 4; i < p.D; i += 128) {
      const float4 r4 = *reinterpret_cast<const float4*>(row + i);
      const float4 t4 = *reinterpret_cast<const float4*>(tok + i);
      d += r4.x * t4.x + r4.y * t4.y + r4.z * t4.z + r4.w * t4.w;
      rr += r4.x * r4.x + r4.y * r4.y + r4.z * r4.z + r4.w * r4.w;
    }
    d = warp_sum(d); rr = warp_sum(rr);
    if (lane == 0) sim[j] = d * inv_tok / fmaxf(sqrtf(rr), 1e-12f) * p.logit_scale_exp;
  }
  __syncthreads();

  // ---- target / background distributions ----
  masked_softmax(sim, n, w_tgt, red, 0, tmask, nullptr);
  masked_softmax(sim, n, w_bgd, red, 1, tmask, nullptr);  // bgd_score (before the split)

  // ---- ascending bitonic sort of bgd_score, sequential running sum, threshold ----
  int npad = 1;
  while (npad < n) npad <<= 1;
  for (int j = threadIdx.x; j < npad; j += blockDim.x) sorted[j] = (j < n) ? w_bgd[j] : INFINITY;
  __syncthreads();
  for (int k = 2; k <= npad; k <<= 1) {
    for (int s = k >> 1; s > 0; s >>= 1) {
      for (int i = threadIdx.x; i < npad; i += blockDim.x) {
        const int ixj = i ^ s;
        if (ixj > i) {
          const float a = sorted[i], c = sorted[ixj];
          const bool up = (i & k) == 0;
          if ((a > c) == up) { sorted[i] = c; sorted[ixj] = a; }
        }
      }
      __syncthreads();
    }
  }
  if (threadIdx.x == 0) {
    float cum = 0.f, thr = 1.0f;
    for (int j = 0; j < n; ++j) {
      cum += sorted[j];
      if (!(cum < 0.25f)) { thr = sorted[j]; break; }
    }
    s_thr = thr;
  }
  __syncthreads();
  const float thr = s_thr;
  for (int j = threadIdx.x; j < n; j += blockDim.x) dmask[j] = w_bgd[j] >= thr ? 1 : 0;
  __syncthreads();
  masked_softmax(sim, n, w_bgd, red, 2, tmask, dmask);  // pure background
  masked_softmax(sim, n, w_dis, red, 3, tmask, dmask);  // distractors

  // ---- three weighted sums over the target rows + query embeddings ----
  // warps stride over the rows (coalesced 16-byte reads of each row), lanes own columns; the per-warp partial sums are
  // then added in warp order through shared memory (fixed order -> deterministic).  `sorted` is reused as [3][D].
  {
    constexpr int MAXV = 8;  // D <= 1024: up to 8 float4 per lane
    const int nv = p.D / 128;
    float4 a0[MAXV], a1[MAXV], a2[MAXV];
#pragma unroll
    for (int k = 0; k < MAXV; ++k) a0[k] = a1[k] = a2[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int j = warp; j < n; j += nwarp) {
      const float* row = (j < p.Nz) ? xb + static_cast<long long>(1 + j) * p.D
                                    : xc + static_cast<long long>(1 + p.Nz + (j - p.Nz)) * p.D;
      const float wt = w_tgt[j], wd = w_dis[j], wb = w_bgd[j];
#pragma unroll
      for (int k = 0; k < MAXV; ++k) {
        if (k < nv) {
          const float4 v = *reinterpret_cast<const float4*>(row + k * 128 + lane * 4);
          a0[k].x += wt * v.x; a0[k].y += wt * v.y; a0[k].z += wt * v.z; a0[k].w += wt * v.w;
          a1[k].x += wd * v.x; a1[k].y += wd * v.y; a1[k].z += wd * v.z; a1[k].w += wd * v.w;
          a2[k].x += wb * v.x; a2[k].y += wb * v.y; a2[k].z += wb * v.z; a2[k].w += wb * v.w;
        }
      }
    }
    float* acc = sorted;  // [3][D] (PROMPTER_SCRATCH floats)
    __syncthreads();
    for (int w = 0; w < nwarp; ++w) {
      if (warp == w) {
#pragma unroll
        for (int k = 0; k < MAXV; ++k) {
          if (k < nv) {
            float4* d0 = reinterpret_cast<float4*>(acc + k * 128 + lane * 4);
            float4* d1 = reinterpret_cast<float4*>(acc + p.D + k * 128 + lane * 4);
            float4* d2 = reinterpret_cast<float4*>(acc + 2 * p.D + k * 128 + lane * 4);
            if (w == 0) { *d0 = a0[k]; *d1 = a1[k]; *d2 = a2[k]; }
            else {
              float4 t = *d0; t.x += a0[k].x; t.y += a0[k].y; t.z += a0[k].z; t.w += a0[k].w; *d0 = t;
              t = *d1; t.x += a1[k].x; t.y += a1[k].y; t.z += a1[k].z; t.w += a1[k].w; *d1 = t;
              t = *d2; t.x += a2[k].x; t.y += a2[k].y; t.z += a2[k].z; t.w += a2[k].w; *d2 = t;
            }
          }
        }
      }
      __syncthreads();
    }
    for (int i = threadIdx.x; i < p.D; i += blockDim.x) {
      const float q0 = p.query_embed[i] + tok[i];
      const float q1 = p.query_embed[p.D + i];
      const float q2 = p.query_embed[2 * p.D + i];
      const long long o = (static_cast<long long>(b) * 3) * p.D + i;
      p.src0[o] = q0; p.src0[o + p.D] = q1; p.src0[o + 2 * p.D] = q2;
      const float s0 = acc[i] + q0, s1 = acc[p.D + i] + q1, s2 = acc[2 * p.D + i] + q2;
      p.src[o] = s0; p.src[o + p.D] = s1; p.src[o + 2 * p.D] = s2;
      p.src_bf16[o] = __float2bfloat16(s0);
      p.src_bf16[o + p.D] = __float2bfloat16(s1);
      p.src_bf16[o + 2 * p.D] = __float2bfloat16(s2);
    }
  }
}

inline size_t prompter_smem_bytes(int D, int n) {
  return sizeof(float) * (static_cast<size_t>(D) + 4 * n + PROMPTER_SCRATCH + 8) + 2 * static_cast<size_t>(n) + 16;
}

// switcher of heads/utils.py:93-97: [src, src_, src][flag]
static __global__ void __launch_bounds__(256) prompt_select_kernel(const float* mlp_out, const float* src0,
                                                            const long long* flag, float* out, int per_seq, int B) {
  pdl_wait();
  pdl_trigger();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= per_seq * B) return;
  const int b = i / per_seq;
  out[i] = (flag[b] == 1) ? src0[i] : mlp_out[i];
}

}  // namespace uvlt
