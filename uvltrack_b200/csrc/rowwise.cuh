// HBM-bound row kernels: LayerNorm (both transformer dialects), BERT embedding, patch / 3x3 im2col, key-bias build.
// All use one warp per row, 128-bit loads/stores, warp-shuffle reductions, fp32 statistics.
#pragma once
#include "common.cuh"

namespace uvlt {

// ----------------------------------------------------------------------------------------------
// LayerNorm.  Reads the fp32 residual stream, writes the bf16 GEMM A-operand.
//   ViT  (block.py:30-31, nn.LayerNorm eps 1e-6, pre-LN): the stream itself is not modified, except that the fusion
//        layers first add the modal embedding to it (mae_vit.py:196) -> `add0/add1` + write-back (dst_mode 1),
//        and the first fusion layer also concatenates image and text streams (src0 / src1).
//   BERT (bert_backbone.py:240-244, eps 1e-12, post-LN): the normalised value replaces the stream (dst_mode 2).
// ----------------------------------------------------------------------------------------------
struct LnParams {
  float* x;            // fp32 token stream; row (b, i) lives at x + b * x_bstride + (x_row_off + i) * D
  long long x_bstride; // elements between sequences
  int x_row_off;       // first row of the normalised range inside a sequence
  int rows;            // rows per sequence
  const float* add0;   // [D] added to rows [0, split) of every sequence before normalising, or nullptr
  const float* add1;   // [D] added to rows [split, rows), or nullptr
  int split;
  int dst_mode;        // 0: stream untouched, 1: write (x + add) back [pre-LN fusion layers], 2: write LN(x) back [post-LN]
  __nv_bfloat16* dst_bf16;  // compact [B*rows, D] GEMM A operand, or nullptr
  const float* gamma;
  const float* beta;
  float eps;
  int total_rows;      // B * rows
  // split-K partial products of the GEMM that last updated these rows (gemm.cuh): added to the stream here, in a
  // fixed order, and the sum is written back (same [B, N, D] offsets as x)
  const float* partials;
  int n_partials;
  int partial_rows;        // only rows [0, partial_rows) of each sequence's range carry partials
  long long partial_stride;
};

constexpr int LN_MAX_PARTIALS = 5;  // fc2 is cut at most six ways (pick_splits)

// PART: the rows carry split-K partials of the preceding fc2.  Their loads are all issued BEFORE the first add (one L2
// round trip instead of one per partial: the batch-1 LayerNorm after a six-way fc2 took 6.5 us against 4.0 us without
// partials, profiles/r02_b1_launches.md); the adds keep the fixed order s = 0 .. n-1.  A separate instantiation, so that
// the large-batch LayerNorm (no partials, HBM bound) keeps its registers and occupancy.
template <int NV, bool PART>  // D = NV * 128  (768 -> 6, 1024 -> 8)
__global__ void __launch_bounds__(256) layernorm_kernel(const LnParams p) {
  pdl_wait();
  pdl_trigger();
  constexpr int D = NV * 128;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = blockIdx.x * (blockDim.x >> 5) + warp;
  if (r >= p.total_rows) return;
  const int b = r / p.rows, i = r - b * p.rows;
  float* src = p.x + static_cast<long long>(b) * p.x_bstride + static_cast<long long>(p.x_row_off + i) * D;
  const float* add = (i < p.split) ? p.add0 : p.add1;
  float4 x[NV];
#pragma unroll
  for (int k = 0; k < NV; ++k) x[k] = *reinterpret_cast<const float4*>(src + k * 128 + lane * 4);
  if (PART) {
    const int np = i < p.partial_rows ? p.n_partials : 0;  // <= LN_MAX_PARTIALS (checked on the host)
    float4 pv[LN_MAX_PARTIALS][NV];
#pragma unroll
    for (int sp = 0; sp < LN_MAX_PARTIALS; ++sp) {
      if (sp < np) {
        const float* ps = p.partials + sp * p.partial_stride + (src - p.x);
#pragma unroll
        for (int k = 0; k < NV; ++k) pv[sp][k] = *reinterpret_cast<const float4*>(ps + k * 128 + lane * 4);
      }
    }
#pragma unroll
    for (int sp = 0; sp < LN_MAX_PARTIALS; ++sp) {
      if (sp < np) {
#pragma unroll
        for (int k = 0; k < NV; ++k) {
          x[k].x += pv[sp][k].x; x[k].y += pv[sp][k].y; x[k].z += pv[sp][k].z; x[k].w += pv[sp][k].w;
        }
      }
    }
  }
  if (add) {
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(add + k * 128 + lane * 4));
      x[k].x += a.x; x[k].y += a.y; x[k].z += a.z; x[k].w += a.w;
    }
  }
  float* dstf = src;
  if ((p.dst_mode == 1 && add) || (p.n_partials > 0 && p.dst_mode != 2)) {
#pragma unroll
    for (int k = 0; k < NV; ++k) *reinterpret_cast<float4*>(dstf + k * 128 + lane * 4) = x[k];
  }
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < NV; ++k) s += (x[k].x + x[k].y) + (x[k].z + x[k].w);
  const float mean = warp_sum(s) * (1.0f / D);
  float v = 0.f;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const float a = x[k].x - mean, b2 = x[k].y - mean, c = x[k].z - mean, d = x[k].w - mean;
    v += (a * a + b2 * b2) + (c * c + d * d);
  }
  const float rstd = rsqrtf(warp_sum(v) * (1.0f / D) + p.eps);
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const float4 g = __ldg(reinterpret_cast<const float4*>(p.gamma + k * 128 + lane * 4));
    const float4 be = __ldg(reinterpret_cast<const float4*>(p.beta + k * 128 + lane * 4));
    float4 y;
    y.x = (x[k].x - mean) * rstd * g.x + be.x;
    y.y = (x[k].y - mean) * rstd * g.y + be.y;
    y.z = (x[k].z - mean) * rstd * g.z + be.z;
    y.w = (x[k].w - mean) * rstd * g.w + be.w;
    if (p.dst_mode == 2) *reinterpret_cast<float4*>(dstf + k * 128 + lane * 4) = y;
    if (p.dst_bf16) {
      uint2 u;
      u.x = pack_bf16x2(y.x, y.y);
      u.y = pack_bf16x2(y.z, y.w);
      *reinterpret_cast<uint2*>(p.dst_bf16 + static_cast<long long>(r) * D + k * 128 + lane * 4) = u;
    }
  }
}

// ----------------------------------------------------------------------------------------------
// BERT embedding (bert_backbone.py:260-274): word[ids] + position[t] + token_type[0] -> LayerNorm(eps 1e-12)
// ----------------------------------------------------------------------------------------------
struct BertEmbedParams {
  const long long* ids;     // [B, T]
  const float* word;        // [vocab, D]
  const float* pos;         // [max_pos, D]
  const float* type0;       // [D]   (token_type_ids are all zero on this path)
  const float* gamma;
  const float* beta;
  float* dst_f32;           // token stream; row (b, t) at dst_f32 + b * dst_bstride + (dst_row_off + t) * D
  long long dst_bstride;
  int dst_row_off;
  __nv_bfloat16* dst_bf16;  // compact [B*T, D] or nullptr
  int T, total_rows, vocab;
};

template <int NV>
__global__ void __launch_bounds__(256) bert_embed_kernel(const BertEmbedParams p) {
  pdl_wait();
  pdl_trigger();
  constexpr int D = NV * 128;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = blockIdx.x * (blockDim.x >> 5) + warp;
  if (r >= p.total_rows) return;
  const int t = r % p.T, bb = r / p.T;
  float* dstf = p.dst_f32 + static_cast<long long>(bb) * p.dst_bstride + static_cast<long long>(p.dst_row_off + t) * D;
  long long id = p.ids[r];
  id = id < 0 ? 0 : (id >= p.vocab ? p.vocab - 1 : id);
  const float* w = p.word + id * D;
  const float* ps = p.pos + static_cast<long long>(t) * D;
  float4 x[NV];
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(w + k * 128 + lane * 4));
    const float4 b = __ldg(reinterpret_cast<const float4*>(ps + k * 128 + lane * 4));
    const float4 c = __ldg(reinterpret_cast<const float4*>(p.type0 + k * 128 + lane * 4));
    // same association order as the reference: (word + position) + token_type
    x[k].x = (a.x + b.x) + c.x; x[k].y = (a.y + b.y) + c.y; x[k].z = (a.z + b.z) + c.z; x[k].w = (a.w + b.w) + c.w;
  }
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < NV; ++k) s += (x[k].x + x[k].y) + (x[k].z + x[k].w);
  const float mean = warp_sum(s) * (1.0f / D);
  float v = 0.f;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const float a = x[k].x - mean, b2 = x[k].y - mean, c = x[k].z - mean, d = x[k].w - mean;
    v += (a * a + b2 * b2) + (c * c + d * d);
  }
  const float rstd = rsqrtf(warp_sum(v) * (1.0f / D) + 1e-12f);
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const float4 g = __ldg(reinterpret_cast<const float4*>(p.gamma + k * 128 + lane * 4));
    const float4 be = __ldg(reinterpret_cast<const float4*>(p.beta + k * 128 + lane * 4));
    float4 y;
    y.x = (x[k].x - mean) * rstd * g.x + be.x;
    y.y = (x[k].y - mean) * rstd * g.y + be.y;
    y.z = (x[k].z - mean) * rstd * g.z + be.z;
    y.w = (x[k].w - mean) * rstd * g.w + be.w;
    *reinterpret_cast<float4*>(dstf + k * 128 + lane * 4) = y;
    if (p.dst_bf16) {
      uint2 u;
      u.x = pack_bf16x2(y.x, y.y);
      u.y = pack_bf16x2(y.z, y.w);
      *reinterpret_cast<uint2*>(p.dst_bf16 + static_cast<long long>(r) * D + k * 128 + lane * 4) = u;
    }
  }
}

// ----------------------------------------------------------------------------------------------
// Patch im2col (mae_vit.py:92,99: Conv2d(3, D, 16, stride 16) == GEMM over 3*16*16 = 768 inputs).
// Output row (b, p) with p in [0, Nz) template patches then [Nz, Nz+Nx) search patches, row-major (h, w);
// column order c*256 + ky*16 + kx matches weight.view(D, -1).  Also writes the cls rows of the token stream.
// Each image comes either as fp32 NCHW already normalised (forward_test operator API) or as the tracker's raw
// uint8 HWC RGB crop, in which case Preprocessor_wo_mask (lib/test/tracker/tracker_utils.py:25-29:
// ((x / 255) - mean) / std in fp32) is fused here.  One warp per patch.
// ----------------------------------------------------------------------------------------------
struct PatchParams {
  const float* tmpl;       // [B, 3, Hz, Hz] or nullptr
  const uint8_t* tmpl_u8;  // [B, Hz, Hz, 3] or nullptr
  const float* srch;       // [B, 3, Hx, Hx] or nullptr
  const uint8_t* srch_u8;  // [B, Hx, Hx, 3] or nullptr
  int B, Hz, Hx;           // image sides (multiples of 16)
  __nv_bfloat16* out;      // [B*(Nz+Nx), 768]
  const float* cls;        // [D]
  float* x_stream;         // token stream; row 0 of each sequence receives the cls token
  long long x_bstride;
  int D;
};

static __global__ void __launch_bounds__(256) patch_im2col_kernel(const PatchParams p) {
  pdl_wait();
  pdl_trigger();
  const int gz = p.Hz >> 4, gx = p.Hx >> 4;
  const int Nz = gz * gz, Nx = gx * gx;
  const int rows = p.B * (Nz + Nx);
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r < rows) {
    const int b = r / (Nz + Nx), pi = r - b * (Nz + Nx);
    const bool is_z = pi < Nz;
    const int side = is_z ? p.Hz : p.Hx;
    const int g = is_z ? gz : gx;
    const int q = is_z ? pi : pi - Nz;
    const int py = q / g, px = q - py * g;
    const float* f32 = is_z ? p.tmpl : p.srch;
    const uint8_t* u8 = is_z ? p.tmpl_u8 : p.srch_u8;
    __nv_bfloat16* dst = p.out + static_cast<long long>(r) * 768;
    if (f32) {
      const float* img = f32 + static_cast<long long>(b) * 3 * side * side;
      // 3 channels x 16 rows x 4 float4 = 192 float4 per patch -> 6 per lane
#pragma unroll
      for (int h = 0; h < 6; ++h) {
        const int idx = lane + h * 32;
        const int c = idx >> 6, ky = (idx >> 2) & 15, kx = (idx & 3) * 4;
        const float4 v = __ldg(reinterpret_cast<const float4*>(
            img + (static_cast<long long>(c) * side + py * 16 + ky) * side + px * 16 + kx));
        uint2 u;
        u.x = pack_bf16x2(v.x, v.y);
        u.y = pack_bf16x2(v.z, v.w);
        *reinterpret_cast<uint2*>(dst + c * 256 + ky * 16 + kx) = u;
      }
    } else {
      // HWC bytes: a patch row is 48 contiguous bytes; lane -> (ky = lane / 2, 8 pixels = 24 bytes)
      const uint8_t* img = u8 + static_cast<long long>(b) * side * side * 3;
      const int ky = lane >> 1, kx0 = (lane & 1) * 8;
      const uint8_t* src = img + (static_cast<long long>(py * 16 + ky) * side + px * 16 + kx0) * 3;
      const uint2 w0 = __ldg(reinterpret_cast<const uint2*>(src));
      const uint2 w1 = __ldg(reinterpret_cast<const uint2*>(src + 8));
      const uint2 w2 = __ldg(reinterpret_cast<const uint2*>(src + 16));
      const uint32_t words[6] = {w0.x, w0.y, w1.x, w1.y, w2.x, w2.y};
      const float mean[3] = {0.485f, 0.456f, 0.406f};
      const float stdv[3] = {0.229f, 0.224f, 0.225f};
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int byte = i * 3 + c;
          const float raw = static_cast<float>((words[byte >> 2] >> ((byte & 3) * 8)) & 0xffu);
          v[i] = (raw / 255.0f - mean[c]) / stdv[c];
        }
        uint4 u;
        u.x = pack_bf16x2(v[0], v[1]); u.y = pack_bf16x2(v[2], v[3]);
        u.z = pack_bf16x2(v[4], v[5]); u.w = pack_bf16x2(v[6], v[7]);
        *reinterpret_cast<uint4*>(dst + c * 256 + ky * 16 + kx0) = u;
      }
    }
  } else {
    // trailing warps: cls token rows
    const int w = r - rows;
    if (w < p.B) {
      float* dst = p.x_stream + static_cast<long long>(w) * p.x_bstride;
      for (int i = lane * 4; i < p.D; i += 128)
        *reinterpret_cast<float4*>(dst + i) = __ldg(reinterpret_cast<const float4*>(p.cls + i));
    }
  }
}

// ----------------------------------------------------------------------------------------------
// 3x3 / pad 1 im2col on an S x S token grid for the box-head conv towers (heads/utils.py:126-130).
// Source: token-major activations [B, S*S, ld] (+ row offset inside each element for the fp32 token stream),
// G channel groups of C channels (one per tower).  Destination [G][B*S*S][9*C] bf16, column order (ky, kx, c) --
// the conv weights are repacked to match at load time.  One warp per (group, pixel, tap).
// ----------------------------------------------------------------------------------------------
struct Im2col3Params {
  const void* src;
  int src_f32;            // 1: fp32 source (token stream), 0: bf16
  long long src_bstride;  // elements between batch elements
  long long src_row_off;  // first token row of the S*S grid inside an element
  long long src_ld;       // row pitch (elements)
  int G, C, S, B;
  __nv_bfloat16* dst;     // [G][B*S*S][9*C]
};

static __global__ void __launch_bounds__(256) im2col3x3_kernel(const Im2col3Params p) {
  pdl_wait();
  pdl_trigger();
  const int lane = threadIdx.x & 31;
  const long long wg = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int SS = p.S * p.S;
  const long long total = static_cast<long long>(p.G) * p.B * SS * 9;
  if (wg >= total) return;
  const int tap = static_cast<int>(wg % 9);
  const long long t1 = wg / 9;
  const int pix = static_cast<int>(t1 % (static_cast<long long>(p.B) * SS));
  const int g = static_cast<int>(t1 / (static_cast<long long>(p.B) * SS));
  const int b = pix / SS, pp = pix - b * SS;
  const int y = pp / p.S, x = pp - y * p.S;
  const int yy = y + tap / 3 - 1, xx = x + tap % 3 - 1;
  __nv_bfloat16* dst = p.dst + (static_cast<long long>(g) * p.B * SS + pix) * (9LL * p.C) + static_cast<long long>(tap) * p.C;
  const bool inside = (yy >= 0 && yy < p.S && xx >= 0 && xx < p.S);
  const long long srow = static_cast<long long>(b) * p.src_bstride + (p.src_row_off + yy * p.S + xx) * p.src_ld +
                         static_cast<long long>(g) * p.C;
  for (int c = lane * 8; c < p.C; c += 256) {
    uint4 u = make_uint4(0, 0, 0, 0);
    if (inside) {
      if (p.src_f32) {
        const float* s = reinterpret_cast<const float*>(p.src) + srow + c;
        const float4 a = *reinterpret_cast<const float4*>(s);
        const float4 d = *reinterpret_cast<const float4*>(s + 4);
        u.x = pack_bf16x2(a.x, a.y); u.y = pack_bf16x2(a.z, a.w);
        u.z = pack_bf16x2(d.x, d.y); u.w = pack_bf16x2(d.z, d.w);
      } else {
        u = *reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p.src) + srow + c);
      }
    }
    *reinterpret_cast<uint4*>(dst + c) = u;
  }
}

// Search rows of the fp32 token stream -> bf16 [B, S*S, D]: the A operand of the box head's first implicit-GEMM
// convolution (replaces the 9x larger im2col matrix)
static __global__ void __launch_bounds__(256) search_to_bf16_kernel(const float* x, long long x_bstride, int row_off,
                                                                    int rows, int D, int B, __nv_bfloat16* dst) {
  pdl_wait();
  pdl_trigger();
  const long long per_seq8 = static_cast<long long>(rows) * D / 8;
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= per_seq8 * B) return;
  const int b = static_cast<int>(i / per_seq8);
  const long long o = (i - b * per_seq8) * 8;
  const float* s = x + b * x_bstride + static_cast<long long>(row_off) * D + o;
  const float4 a = *reinterpret_cast<const float4*>(s);
  const float4 d = *reinterpret_cast<const float4*>(s + 4);
  uint4 u;
  u.x = pack_bf16x2(a.x, a.y); u.y = pack_bf16x2(a.z, a.w);
  u.z = pack_bf16x2(d.x, d.y); u.w = pack_bf16x2(d.z, d.w);
  *reinterpret_cast<uint4*>(dst + (static_cast<long long>(b) * rows * D + o)) = u;
}

// ----------------------------------------------------------------------------------------------
// Key-ignore masks -> additive attention biases (modality_unified_feature_extractor.py:43-50, bert_backbone.py:746-748)
//   visual [B, Nv]: cls, template ignored iff flag == 1          (masked_fill -1e10 dialect)
//   joint  [B, N ]: + text j ignored iff text_mask[j] == 0 or flag == 0
//   bert   [B, T ]: (1 - text_mask) * -10000                      (additive dialect)
// ----------------------------------------------------------------------------------------------
struct BiasParams {
  const long long* flag;   // [B]
  const float* text_mask;  // [B, T]
  int B, Nz, Nx, T;
  float* bias_vis;    // [B, 1+Nz+Nx]
  float* bias_joint;  // [B, 1+Nz+Nx+T]
  float* bias_bert;   // [B, T]
  // optional staging of the per-call inputs into engine-owned buffers (so a captured CUDA graph sees fixed addresses)
  long long* flag_copy;    // [B] or nullptr
  float* mask_copy;        // [B, T] or nullptr
  const float* prompt;     // [B, 3, D] or nullptr
  float* prompt_copy;      // [B, 3, D]
  int prompt_elems;        // B * 3 * D
};

static __global__ void __launch_bounds__(256) build_bias_kernel(const BiasParams p) {
  pdl_wait();
  pdl_trigger();
  const int Nv = 1 + p.Nz + p.Nx, N = Nv + p.T;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (p.prompt) {
    for (int j = i; j < p.prompt_elems; j += gridDim.x * blockDim.x) p.prompt_copy[j] = p.prompt[j];
  }
  if (p.flag_copy && i < p.B) p.flag_copy[i] = p.flag[i];
  if (i >= p.B * N) return;
  const int b = i / N, k = i - b * N;
  const long long f = p.flag[b];
  constexpr float NEG = -1e10f;
  if (k < Nv) {
    const float v = (k < 1 + p.Nz && f == 1) ? NEG : 0.0f;
    p.bias_vis[b * Nv + k] = v;
    p.bias_joint[i] = v;
  } else {
    const int j = k - Nv;
    const float m = p.text_mask[b * p.T + j];
    const float keep = m * (f != 0 ? 1.0f : 0.0f);      // reference: text.mask * (flag != 0), then .bool()
    p.bias_joint[i] = (keep != 0.0f) ? 0.0f : NEG;
    p.bias_bert[b * p.T + j] = (1.0f - m) * -10000.0f;
    if (p.mask_copy) p.mask_copy[b * p.T + j] = m;
  }
}

// ----------------------------------------------------------------------------------------------
// host launch helpers (return 0 on success; the caller reports cudaGetLastError)
// ----------------------------------------------------------------------------------------------
inline int launch_layernorm(const LnParams& p, int D, cudaStream_t s) {
  const int blocks = (p.total_rows + 7) / 8;
  if (p.n_partials > LN_MAX_PARTIALS || p.n_partials < 0) return 1;
  const bool part = p.n_partials > 0;
  if (D == 768 && part) UVLT_LAUNCH((layernorm_kernel<6, true>), dim3(blocks), dim3(256), 0, s, p);
  else if (D == 768) UVLT_LAUNCH((layernorm_kernel<6, false>), dim3(blocks), dim3(256), 0, s, p);
  else if (D == 1024 && part) UVLT_LAUNCH((layernorm_kernel<8, true>), dim3(blocks), dim3(256), 0, s, p);
  else if (D == 1024) UVLT_LAUNCH((layernorm_kernel<8, false>), dim3(blocks), dim3(256), 0, s, p);
  else return 1;
  return cudaGetLastError() != cudaSuccess;
}
inline int launch_bert_embed(const BertEmbedParams& p, int D, cudaStream_t s) {
  const int blocks = (p.total_rows + 7) / 8;
  if (D == 768) UVLT_LAUNCH(bert_embed_kernel<6>, dim3(blocks), dim3(256), 0, s, p);
  else if (D == 1024) UVLT_LAUNCH(bert_embed_kernel<8>, dim3(blocks), dim3(256), 0, s, p);
  else return 1;
  return cudaGetLastError() != cudaSuccess;
}
inline int launch_patch_im2col(const PatchParams& p, cudaStream_t s) {
  const int Np = (p.Hz / 16) * (p.Hz / 16) + (p.Hx / 16) * (p.Hx / 16);
  const long long warps = static_cast<long long>(p.B) * Np + p.B;
  UVLT_LAUNCH(patch_im2col_kernel, dim3(static_cast<unsigned>((warps + 7) / 8)), dim3(256), 0, s, p);
  return cudaGetLastError() != cudaSuccess;
}
inline int launch_im2col3x3(const Im2col3Params& p, cudaStream_t s) {
  const long long warps = 9LL * p.G * p.B * p.S * p.S;
  UVLT_LAUNCH(im2col3x3_kernel, dim3(static_cast<unsigned>((warps + 7) / 8)), dim3(256), 0, s, p);
  return cudaGetLastError() != cudaSuccess;
}
inline int launch_build_bias(const BiasParams& p, cudaStream_t s) {
  const int total = p.B * (1 + p.Nz + p.Nx + p.T);
  UVLT_LAUNCH(build_bias_kernel, dim3((total + 255) / 256), dim3(256), 0, s, p);
  return cudaGetLastError() != cudaSuccess;
}

// ---------------------------------------------------------------------------------------------------------------
// split-K reduction with the epilogue the split GEMM could not apply: out = act(sum_s part[s] + bias), bf16.
// Partials are added in the fixed order s = 0 .. splits-1 (the result does not depend on the grid).
// ---------------------------------------------------------------------------------------------------------------
struct SplitReduceParams {
  const float* part;      // [splits][M, N]
  long long stride;       // elements between partials
  int splits;
  const float* bias;      // [N] or nullptr
  __nv_bfloat16* out;     // [M, N]
  long long total;        // M * N  (N % 4 == 0)
  int N;
  int relu;
};

static __global__ void __launch_bounds__(256) splitk_reduce_kernel(const SplitReduceParams p) {
  pdl_wait();
  const long long i = (static_cast<long long>(blockIdx.x) * 256 + threadIdx.x) * 4;
  if (i >= p.total) return;
  float4 acc = *reinterpret_cast<const float4*>(p.part + i);
  for (int s = 1; s < p.splits; ++s) {
    const float4 v = *reinterpret_cast<const float4*>(p.part + s * p.stride + i);
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  if (p.bias) {
    const float4 b = *reinterpret_cast<const float4*>(p.bias + static_cast<int>(i % p.N));
    acc.x += b.x; acc.y += b.y; acc.z += b.z; acc.w += b.w;
  }
  if (p.relu) {
    acc.x = fmaxf(acc.x, 0.f); acc.y = fmaxf(acc.y, 0.f); acc.z = fmaxf(acc.z, 0.f); acc.w = fmaxf(acc.w, 0.f);
  }
  __nv_bfloat162 lo = __floats2bfloat162_rn(acc.x, acc.y), hi = __floats2bfloat162_rn(acc.z, acc.w);
  uint2 o;
  o.x = *reinterpret_cast<uint32_t*>(&lo);
  o.y = *reinterpret_cast<uint32_t*>(&hi);
  *reinterpret_cast<uint2*>(p.out + i) = o;
}
inline int launch_splitk_reduce(const SplitReduceParams& p, cudaStream_t s) {
  const long long threads = p.total / 4;
  UVLT_LAUNCH(splitk_reduce_kernel, dim3(static_cast<unsigned>((threads + 255) / 256)), dim3(256), 0, s, p);
  return cudaGetLastError() != cudaSuccess;
}

}  // namespace uvlt
