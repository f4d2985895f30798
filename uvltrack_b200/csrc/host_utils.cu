#include "host_utils.h"

#include <cudaTypedefs.h>

#include <algorithm>
#include <mutex>

namespace uvlt {

static thread_local std::string g_err;
void set_error(const std::string& msg) { g_err = msg; }
const char* get_error() { return g_err.c_str(); }

// cuTensorMapEncodeTiled is fetched through the runtime so the library has no link-time libcuda dependency
// (it must still dlopen on a GPU-less host for the symbol-export test).
static PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  });
  return fn;
}

int make_tma_bf16_3d(CUtensorMap* out, const void* base, uint64_t dim0, uint64_t dim1, uint64_t dim2,
                     uint64_t stride1_bytes, uint64_t stride2_bytes, uint32_t box_rows) {
  auto fn = get_encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled driver entry point unavailable (no CUDA driver?)");
    return 1;
  }
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (stride1_bytes & 15) || (stride2_bytes & 15)) {
    set_error("TMA operand base/strides must be 16-byte aligned");
    return 1;
  }
  cuuint64_t dims[3] = {dim0, dim1, dim2};
  cuuint64_t strides[2] = {stride1_bytes, stride2_bytes};
  cuuint32_t box[3] = {64, box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult " + std::to_string(static_cast<int>(r)));
    return 1;
  }
  return 0;
}

int init_kernel_attributes() {
  static int status = -1;
  static std::mutex mu;
  std::lock_guard<std::mutex> lk(mu);
  if (status == 0) return 0;
  UVLT_CUDA_OK(cudaFuncSetAttribute(gemm_bf16_tn_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    GemmSmem<32>::total(GemmSmem<32>::STAGES_1CTA)));
  UVLT_CUDA_OK(cudaFuncSetAttribute(gemm_bf16_tn_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    GemmSmem<64>::total(GemmSmem<64>::STAGES_1CTA)));
  UVLT_CUDA_OK(cudaFuncSetAttribute(gemm_bf16_tn_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    GemmSmem<128>::total(GemmSmem<128>::STAGES_1CTA)));
  UVLT_CUDA_OK(cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AttnSmem::TOTAL));
  UVLT_CUDA_OK(cudaFuncSetAttribute(attention_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                                    cudaSharedmemCarveoutMaxShared));
  status = 0;
  return 0;
}

int pick_bn(int M, int N, int groups) {
  // Widest tile that still gives >= 90 CTAs (measured on B200 at M = 513: wider tiles cut the L2 -> SM operand traffic,
  // which bounds these launches, until fewer than ~2/3 of the SMs have work); below that, the narrowest legal tile.
  const int m_tiles = (M + GEMM_BM - 1) / GEMM_BM;
  const int cands[3] = {128, 64, 32};
  int best = 0;
  for (int bn : cands) {
    if (N % bn) continue;
    best = bn;
    if (static_cast<long long>(m_tiles) * (N / bn) * groups >= 90) break;
  }
  return best;
}

int gemm_prepare(GemmLaunch* g, const void* A, long long a_ld, long long a_gstride, const void* W, long long w_ld,
                 long long w_gstride, int M, int N, int K, int groups, int bn, const GemmEpilogue& ep) {
  if (bn == 0) bn = pick_bn(M, N, groups);
  if (bn != 32 && bn != 64 && bn != 128) {
    set_error("gemm: N must be a multiple of 32");
    return 1;
  }
  if (K % GEMM_BK || N % bn || M <= 0) {
    set_error("gemm: need K % 64 == 0, N % BN == 0, M > 0 (got M=" + std::to_string(M) + " N=" + std::to_string(N) +
              " K=" + std::to_string(K) + ")");
    return 1;
  }
  g->shape = GemmShape{M, N, K, 0};
  {
    // ring depth: a grid that fits one CTA per SM is latency bound -> put as much of K in flight as shared memory
    // allows; larger grids keep ~96 KB rings so two CTAs share an SM and overlap epilogue with mainloop
    const long long tiles = static_cast<long long>((M + GEMM_BM - 1) / GEMM_BM) * (N / bn) * groups;
    const int s1 = bn == 32 ? GemmSmem<32>::STAGES_1CTA : bn == 64 ? GemmSmem<64>::STAGES_1CTA : GemmSmem<128>::STAGES_1CTA;
    const int s2 = bn == 32 ? GemmSmem<32>::STAGES_2CTA : bn == 64 ? GemmSmem<64>::STAGES_2CTA : GemmSmem<128>::STAGES_2CTA;
    int st = tiles <= 148 ? s1 : s2;
    st = std::min(st, K / GEMM_BK);
    g->shape.stages = std::max(st, 1);
  }
  g->ep = ep;
  g->bn = bn;
  g->groups = groups;
  if (groups == 1) { a_gstride = static_cast<long long>(M) * a_ld; w_gstride = static_cast<long long>(N) * w_ld; }
  if (make_tma_bf16_3d(&g->tma_a, A, K, M, groups, a_ld * 2, a_gstride * 2, GEMM_BM)) return 1;
  if (make_tma_bf16_3d(&g->tma_w, W, K, N, groups, w_ld * 2, w_gstride * 2, bn)) return 1;
  return 0;
}

int gemm_launch(const GemmLaunch& g, cudaStream_t stream) {
  dim3 grid(g.shape.N / g.bn, (g.shape.M + GEMM_BM - 1) / GEMM_BM, g.groups);
  switch (g.bn) {
    case 32:
      UVLT_LAUNCH(gemm_bf16_tn_kernel<32>, dim3(grid), dim3(GEMM_THREADS), GemmSmem<32>::total(g.shape.stages), stream, g.tma_a, g.tma_w, g.shape, g.ep);
      break;
    case 64:
      UVLT_LAUNCH(gemm_bf16_tn_kernel<64>, dim3(grid), dim3(GEMM_THREADS), GemmSmem<64>::total(g.shape.stages), stream, g.tma_a, g.tma_w, g.shape, g.ep);
      break;
    default:
      UVLT_LAUNCH(gemm_bf16_tn_kernel<128>, dim3(grid), dim3(GEMM_THREADS), GemmSmem<128>::total(g.shape.stages), stream, g.tma_a, g.tma_w, g.shape, g.ep);
      break;
  }
  UVLT_CUDA_OK(cudaGetLastError());
  return 0;
}

int attn_prepare(AttnLaunch* a, const void* qkv, int B, int n, int H, const float* bias, void* out, const void* vt,
                 int n_pad) {
  (void)n_pad;
  if (vt) {
    set_error("attention: the separate V^T operand of the bring-up kernel is no longer supported (pass NULL)");
    return 1;
  }
  if (n > ATT_MAX_KV || n <= 0) {
    set_error("attention: sequence length out of range");
    return 1;
  }
  const long long D3 = 3LL * H * ATT_D;
  // [B, n, 3D]: rows >= n are out of bounds inside each batch element -> TMA zero fill (no cross-sequence reads)
  if (make_tma_bf16_3d(&a->tma_qkv, qkv, D3, n, B, D3 * 2, static_cast<uint64_t>(n) * D3 * 2, ATT_BQ)) return 1;
  a->p.n = n;
  a->p.H = H;
  a->p.scale_log2 = 0.125f * 1.4426950408889634f;  // head_dim 64
  a->p.bias = bias;
  a->p.out = reinterpret_cast<__nv_bfloat16*>(out);
  a->B = B;
  return 0;
}

int attn_launch(const AttnLaunch& a, cudaStream_t stream) {
  dim3 grid((a.p.n + ATT_BQ - 1) / ATT_BQ, a.p.H, a.B);
  UVLT_LAUNCH(attention_kernel, grid, dim3(ATT_THREADS), AttnSmem::TOTAL, stream, a.tma_qkv, a.p);
  UVLT_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace uvlt
