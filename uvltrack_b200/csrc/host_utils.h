// Host-side helpers shared by the op-level entry points and the engine: error plumbing, TMA descriptor encoding,
// kernel launch wrappers.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "attention.cuh"
#include "attention2.cuh"
#include "attention3.cuh"
#include "gemm.cuh"

namespace uvlt {

// thread-local last error text for the C ABI (uvlt_last_error)
void set_error(const std::string& msg);
const char* get_error();

#define UVLT_CUDA_OK(expr)                                                                         \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess) {                                                                       \
      ::uvlt::set_error(std::string(#expr) + " failed: " + cudaGetErrorString(_e));                \
      return 1;                                                                                    \
    }                                                                                              \
  } while (0)

// 3-D tiled bf16 tensor map with 128-byte swizzle.  dims/strides innermost first; box = {64, box_rows, 1}.
// Returns 0 on success.
int make_tma_bf16_3d(CUtensorMap* out, const void* base, uint64_t dim0, uint64_t dim1, uint64_t dim2,
                     uint64_t stride1_bytes, uint64_t stride2_bytes, uint32_t box_rows);

struct GemmLaunch {
  CUtensorMap tma_a, tma_w;
  CUtensorMap tma_a_half;  // 64-row boxes of A for the multicast variant
  CUtensorMap tma_out;     // bf16 output matrix, 128 x 64 boxes (CTA-pair kernel: TMA bulk stores)
  bool multicast;
  bool two_sm;  // CTA-pair kernel (tcgen05 cta_group::2, 256 x 256 pair tile); bn == 256
  GemmShape shape;
  GemmEpilogue ep;
  int bn;      // 32 / 64 / 128
  int groups;  // grid.z
};

// Builds the tensor maps for A [groups][M,K] and W [groups][N,K] (both bf16, K contiguous).
// splits: 1 = no split-K; > 1 requires an fp32 output with ep.split_out set and K % (64 * splits) == 0.
int gemm_prepare(GemmLaunch* g, const void* A, long long a_ld, long long a_gstride, const void* W, long long w_ld,
                 long long w_gstride, int M, int N, int K, int groups, int bn, const GemmEpilogue& ep, int splits = 1);
// Implicit-GEMM 3x3 / pad 1 convolution over an S x S token grid (GemmShape::conv_S): src = bf16 feature map
// [B, S*S, src_ld] whose channels [g * Cin, (g + 1) * Cin) feed group g; W [groups][N, 9 * Cin] in (ky, kx, c) column
// order.  Same epilogues / split-K as gemm_prepare.  conv3x3_implicit_ok(S, Cin): the shapes this path supports (other
// sizes go through im2col3x3_kernel + a plain GEMM).
bool conv3x3_implicit_ok(int S, int Cin);
int gemm_prepare_conv3x3(GemmLaunch* g, const void* src, long long src_ld, int S, int B, int Cin, const void* W,
                         long long w_gstride, int N, int groups, int bn, const GemmEpilogue& ep, int splits = 1);
int make_tma_bf16_4d(CUtensorMap* out, const void* base, const uint64_t dims[4], const uint64_t strides_bytes[3],
                     const uint32_t box[4]);
// split-K factor for an [M, N, K] fp32-output GEMM whose consumer can add partials: > 1 only when the unsplit grid
// would leave most SMs idle (small M) and K is long
int pick_splits(int M, int N, int K);
// split-K factor of the box head's first conv GEMM (raw partials + splitk_reduce_kernel); 1 = do not split
int pick_head_splits(int M, int N, int K);
int gemm_launch(const GemmLaunch& g, cudaStream_t stream);
int pick_bn(int M, int N, int groups, bool out_f32, int act);
// true: the persistent CTA-pair kernel takes this GEMM (rule and measurements in host_utils.cu)
bool use_2sm(int M, int N, int K, int groups, int splits);
extern int g_gemm_multicast;  // UVLT_MULTICAST=0 disables the cluster / TMA-multicast GEMM variant

struct AttnLaunch {
  CUtensorMap tma_qkv;  // 128-row boxes (Q tile)
  CUtensorMap tma_kv;   // 64-row boxes (K / V tiles)
  AttnParams p;
  int B;
  bool split;  // key-split cluster variant (two CTAs per query tile)
  // second-generation kernel (attention2.cuh): two query tiles (or two key halves of one tile) per CTA, P in TMEM
  bool v2;
  bool poly;   // a quarter of the exponentials on the FMA pipe
  Attn2Params p2;
  CUtensorMap tma_o;  // output [B, n, H*64] bf16, 128-row boxes (TMA bulk store of a query tile)
  // third-generation kernel (attention3.cuh): persistent, two threads per query row; large grids only (small grids keep
  // the key-split cluster variant of the first kernel)
  bool v3;
  int grid3;   // CTAs: min(SMs, work items)
  int var3;    // softmax arithmetic variant (UVLT_ATTN_POLY, attention3.cuh)
  Attn3Params p3;
};
// capacity_batch: the batch size the split decision is made for (the engine passes its max_batch, so that a sequence's
// result does not depend on how many sequences share the call; 0 = use B)
int attn_prepare(AttnLaunch* a, const void* qkv, int B, int n, int H, const float* bias, void* out, const void* vt,
                 int n_pad, int capacity_batch = 0);
int attn_launch(const AttnLaunch& a, cudaStream_t stream);
// the second- and third-generation kernels live in their own translation unit (attention_big.cu), see there
int attn23_init_attributes();
int attn23_launch(const AttnLaunch& a, cudaStream_t stream);

void runtime_switches(int32_t* out6);  // see uvlt_runtime_switches (include/uvlt.h)
int init_kernel_attributes();  // cudaFuncSetAttribute(max dynamic smem) for every instantiation; idempotent

}  // namespace uvlt
