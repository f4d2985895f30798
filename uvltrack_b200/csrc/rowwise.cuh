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
  const float* src0;   // [B, rows0, D]
  const float* src1;   // [B, rows1, D] or nullptr
  int rows0, rows1;    // rows per batch element in src0 / src1 (output has rows0 + rows1 rows per element)
  const float* add0;   // [D] added to output rows [0, split) of every element, or nullptr
  const float* add1;   // [D] added to output rows [split, rows0+rows1), or nullptr
  int split;
  float* dst_f32;      // [B, rows0+rows1, D] or nullptr
  int dst_mode;        // 0 none, 1 write (x + add) [pre-norm], 2 write normalised value [post-LN]
  __nv_bfloat16* dst_bf16;  // [B*(rows0+rows1), D] or nullptr
  const float* gamma;
  const float* beta;
  float eps;
  int total_rows;      // B * (rows0 + rows1)
};

template <int NV>  // D = NV * 128  (768 -> 6, 1024 -> 8)
__global__ void __launch_bounds__(256) layernorm_kernel(const LnParams p) {
  constexpr int D = NV * 128;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = blockIdx.x * (blockDim.x >> 5) + warp;
  if (r >= p.total_rows) return;
  const int rows = p.rows0 + p.rows1;
  const int b = r / rows, i = r - b * rows;
  const float* src = (i < p.rows0) ? p.src0 + (static_cast<long long>(b) * p.rows0 + i) * D
                                   : p.src1 + (static_cast<long long>(b) * p.rows1 + (i - p.rows0)) * D;
  const float* add = (i < p.split) ? p.add0 : p.add1;
  float4 x[NV];
#pragma unroll
  for (int k = 0; k < NV; ++k) x[k] = *reinterpret_cast<const float4*>(src + k * 128 + lane * 4);
  if (add) {
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(add + k * 128 + lane * 4));
      x[k].x += a.x; x[k].y += a.y; x[k].z += a.z; x[k].w += a.w;
    }
  }
  float* dstf = p.dst_f32 ? p.dst_f32 + static_cast<long long>(r) * D : nullptr;
  if (dstf && p.dst_mode == 1) {
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
    if (dstf && p.dst_mode == 2) *reinterpret_cast<float4*>(dstf + k * 128 + lane * 4) = y;
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
  float* dst_f32;           // [B*T, D]
  __nv_bfloat16* dst_bf16;  // [B*T, D]
  int T, total_rows, vocab;
};

template <int NV>
__global__ void __launch_bounds__(256) bert_embed_kernel(const BertEmbedParams p) {
  constexpr int D = NV * 128;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = blockIdx.x * (blockDim.x >> 5) + warp;
  if (r >= p.total_rows) return;
  const int t = r % p.T;
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
    *reinterpret_cast<float4*>(p.dst_f32 + static_cast<long long>(r) * D + k * 128 + lane * 4) = y;
    uint2 u;
    u.x = pack_bf16x2(y.x, y.y);
    u.y = pack_bf16x2(y.z, y.w);
    *reinterpret_cast<uint2*>(p.dst_bf16 + static_cast<long long>(r) * D + k * 128 + lane * 4) = u;
  }
}

// ----------------------------------------------------------------------------------------------
// Patch im2col (mae_vit.py:92,99: Conv2d(3, D, 16, stride 16) == GEMM over 3*16*16 = 768 inputs).
// Output row (b, p) with p in [0, Nz) template patches then [Nz, Nz+Nx) search patches, row-major (h, w);
// column order c*256 + ky*16 + kx matches weight.view(D, -1).  Also writes the cls rows of the token stream.
// One warp per (row, channel): 16 rows x 16 px of one patch channel.
// ----------------------------------------------------------------------------------------------
struct PatchParams {
  const float* tmpl;   // [B, 3, Hz, Hz]
  const float* srch;   // [B, 3, Hx, Hx]
  int B, Hz, Hx;       // image sides (multiples of 16)
  __nv_bfloat16* out;  // [B*(Nz+Nx), 768]
  const float* cls;    // [D]
  float* x_stream;     // [B, 1+Nz+Nx, D] (only row 0 of each element is written here)
  int D;
};

__global__ void __launch_bounds__(256) patch_im2col_kernel(const PatchParams p) {
  const int gz = p.Hz >> 4, gx = p.Hx >> 4;
  const int Nz = gz * gz, Nx = gx * gx;
  const int rows = p.B * (Nz + Nx);
  const int warp_global = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (warp_global < rows * 3) {
    const int r = warp_global / 3, c = warp_global - r * 3;
    const int b = r / (Nz + Nx), pi = r - b * (Nz + Nx);
    const float* img;
    int side, py, px;
    if (pi < Nz) {
      img = p.tmpl + (static_cast<long long>(b) * 3 + c) * p.Hz * p.Hz;
      side = p.Hz; py = pi / gz; px = pi - py * gz;
    } else {
      const int q = pi - Nz;
      img = p.srch + (static_cast<long long>(b) * 3 + c) * p.Hx * p.Hx;
      side = p.Hx; py = q / gx; px = q - py * gx;
    }
    __nv_bfloat16* dst = p.out + static_cast<long long>(r) * 768 + c * 256;
    // 256 px = 64 float4; lane handles float4 index lane and lane+32: ky = idx/4, kx4 = idx%4
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int idx = lane + h * 32;
      const int ky = idx >> 2, kx = (idx & 3) * 4;
      const float4 v = __ldg(reinterpret_cast<const float4*>(img + static_cast<long long>(py * 16 + ky) * side + px * 16 + kx));
      uint2 u;
      u.x = pack_bf16x2(v.x, v.y);
      u.y = pack_bf16x2(v.z, v.w);
      *reinterpret_cast<uint2*>(dst + ky * 16 + kx) = u;
    }
  } else {
    // trailing warps: cls token rows
    const int w = warp_global - rows * 3;
    if (w < p.B) {
      float* dst = p.x_stream + static_cast<long long>(w) * (1 + Nz + Nx) * p.D;
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

__global__ void __launch_bounds__(256) im2col3x3_kernel(const Im2col3Params p) {
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
};

__global__ void __launch_bounds__(256) build_bias_kernel(const BiasParams p) {
  const int Nv = 1 + p.Nz + p.Nx, N = Nv + p.T;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
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
  }
}

}  // namespace uvlt
