#include "host_utils.h"

#include <cstdlib>

#include <cudaTypedefs.h>

#include <algorithm>
#include <mutex>
#include <set>

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

int make_tma_bf16_4d(CUtensorMap* out, const void* base, const uint64_t dims[4], const uint64_t strides_bytes[3],
                     const uint32_t box[4]) {
  auto fn = get_encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled driver entry point unavailable (no CUDA driver?)");
    return 1;
  }
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (strides_bytes[0] & 15) || (strides_bytes[1] & 15) || (strides_bytes[2] & 15)) {
    set_error("TMA operand base/strides must be 16-byte aligned");
    return 1;
  }
  cuuint64_t d[4] = {dims[0], dims[1], dims[2], dims[3]};
  cuuint64_t st[3] = {strides_bytes[0], strides_bytes[1], strides_bytes[2]};
  cuuint32_t bx[4] = {box[0], box[1], box[2], box[3]};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), d, st, bx, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (4-D) failed with CUresult " + std::to_string(static_cast<int>(r)));
    return 1;
  }
  return 0;
}

static int g_num_sms = 148;
// UVLT_ATTN_SPLIT=0 disables the key-split cluster variant of the attention kernel (A/B timing)
static int g_attn_split = [] {
  const char* e = getenv("UVLT_ATTN_SPLIT");
  return (e && e[0] == '0') ? 0 : 1;
}();

// Attention kernel selection.  Default (UVLT_ATTN_V unset or 0): the third-generation kernel (attention3.cuh: persistent
// CTAs, two threads per query row, P in tensor memory) for large grids of mid-length sequences, the first kernel
// (attention.cuh, three CTAs per SM, 64-key blocks; its key-split cluster variant for small grids) everywhere else.
// UVLT_ATTN_V=1 / 2 / 3 forces one generation (2 = attention2.cuh, the one-CTA-per-SM predecessor of 3; 3 still leaves
// small grids to the key-split variant unless UVLT_ATTN_SPLIT=0).
// Measured on B200 (tools/kernel_sweep.py attn, profiles/r02_attention.md), us per launch, v1 / v3:
//   H=12: B=32 n=553 76.6 / 70.0, n=513 72.7 / 68.2, n=361 35.3 / 35.0   B=16 n=553 41.9 / 36.9   B=8 n=553 26.8 / 22.3
//         B=4 n=553 18.7 / 13.4
//   H=16 (UVLTrack-L): B=8 n=1193 89.9 / 89.3, B=4 51.9 / 54.6, B=2 32.2 / 37.6 (few work items per CTA: tail imbalance)
// UVLT_ATTN_POLY: softmax arithmetic variant of v3 (0..5, attention3.cuh; default 2 = packed FFMA2 / FADD2, every
// exponential on the MUFU), or for v2 1 = a quarter of the exponentials on the FMA pipe.
static int g_attn_v = [] {
  const char* e = getenv("UVLT_ATTN_V");
  return (e && e[0] >= '1' && e[0] <= '3') ? e[0] - '0' : 0;  // 0 = automatic
}();
// UVLT_ATTN_PINGPONG=0: the two slots of the third-generation kernel do not take turns on the MUFU (A/B timing)
static int g_attn_pingpong = [] {
  const char* e = getenv("UVLT_ATTN_PINGPONG");
  return (e && e[0] == '0') ? 0 : 1;
}();
// UVLT_ATTN_GRID=<n> caps the persistent grid of the third-generation kernel (tests: many work items per CTA)
static int g_attn_grid_cap = [] {
  const char* e = getenv("UVLT_ATTN_GRID");
  return e ? atoi(e) : 0;
}();
static int g_attn_poly = [] {
  const char* e = getenv("UVLT_ATTN_POLY");
  return (e && e[0] >= '0' && e[0] <= '9') ? e[0] - '0' : -1;  // v2: 0 / 1; v3: variant 0..5 (attention3.cuh); -1 = default
}();

// cudaFuncSetAttribute applies to the CURRENT device: the opt-in is tracked per device ordinal, so a process that creates
// engines on several GPUs gets it on each of them (one process per GPU remains the supported layout, dp.py).
int init_kernel_attributes() {
  static std::set<int> done;
  static std::mutex mu;
  std::lock_guard<std::mutex> lk(mu);
  int cur_dev = 0;
  UVLT_CUDA_OK(cudaGetDevice(&cur_dev));
  if (done.count(cur_dev)) return 0;
#define UVLT_GEMM_ATTR(BN, EPI)                                                                                     \
  UVLT_CUDA_OK(cudaFuncSetAttribute(gemm_bf16_tn_kernel<BN, EPI, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                    GemmSmem<BN>::total(GemmSmem<BN>::STAGES_2CTA)));                                 \
  UVLT_CUDA_OK(cudaFuncSetAttribute(gemm_bf16_tn_kernel<BN, EPI, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                                    GemmSmem<BN>::total(GemmSmem<BN>::STAGES_2CTA)))
#define UVLT_GEMM_ATTR_BN(BN) \
  UVLT_GEMM_ATTR(BN, EPI_BF16); UVLT_GEMM_ATTR(BN, EPI_BF16_GELU); UVLT_GEMM_ATTR(BN, EPI_BF16_RELU); UVLT_GEMM_ATTR(BN, EPI_F32)
  UVLT_GEMM_ATTR_BN(32);
  UVLT_GEMM_ATTR_BN(64);
  UVLT_GEMM_ATTR_BN(128);
  UVLT_GEMM_ATTR_BN(256);
#define UVLT_GEMM2_ATTR(EPI)                                                                                    \
  UVLT_CUDA_OK(cudaFuncSetAttribute(gemm_bf16_tn_2sm_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                    gemm2_smem_bytes(GEMM2_MAX_STAGES)))
  UVLT_GEMM2_ATTR(EPI_BF16); UVLT_GEMM2_ATTR(EPI_BF16_GELU); UVLT_GEMM2_ATTR(EPI_BF16_RELU); UVLT_GEMM2_ATTR(EPI_F32);
  UVLT_CUDA_OK(cudaFuncSetAttribute(attention_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, AttnSmem::TOTAL));
  UVLT_CUDA_OK(cudaFuncSetAttribute(attention_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                    cudaSharedmemCarveoutMaxShared));
  UVLT_CUDA_OK(cudaFuncSetAttribute(attention_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    AttnSmem::TOTAL_SPLIT));
  UVLT_CUDA_OK(cudaFuncSetAttribute(attention_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                    cudaSharedmemCarveoutMaxShared));
  if (attn23_init_attributes()) return 1;  // attention_big.cu
  {
    int dev = 0, sms = 0;
    UVLT_CUDA_OK(cudaGetDevice(&dev));
    UVLT_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    g_num_sms = sms > 1 ? sms : 148;
  }
  done.insert(cur_dev);
  return 0;
}

// Off by default: measured on B200 (profiles/r01_gemm_multicast.txt) the cluster variant is 5-20 % SLOWER at every
// batch size -- the L2 already merges the two CTAs' near-simultaneous requests for the same A tile, so multicast saves no
// L2 bandwidth and only adds the cross-CTA slot handshake.  UVLT_MULTICAST=1 enables it for experiments.
int g_gemm_multicast = [] {
  const char* e = getenv("UVLT_MULTICAST");
  return (e && e[0] == '1') ? 1 : 0;
}();

// CTA-pair persistent GEMM (gemm_bf16_tn_2sm_kernel): UVLT_GEMM_2SM=0 never, 1 by the rule below (default), 2 wherever
// legal.  Measured on B200 (tools/kernel_sweep.py gemm, profiles/r01_gemm_2sm.md), us per GEMM, best one-CTA tile vs pair:
//   M = 16416: qkv 56.1 -> 49.0, proj 35.5 -> 30.1, fc1 99.5 -> 83.1, fc2 87.4 -> 77.5
//   M =  8208: qkv 27.3 -> 27.7, proj 17.2 -> 19.6, fc1 53.1 -> 46.7, fc2 42.9 -> 41.2
//   M =  4104: qkv 19.1 -> 19.2, proj 13.7 -> 13.1, fc1 28.3 -> 26.4, fc2 28.4 -> 24.4       M = 2052: one-CTA tiles win
// The persistent kernel loses when its last round of tiles leaves most of the 74 TPCs idle (proj at M = 8208: 99 tiles).
int g_gemm_2sm = [] {
  const char* e = getenv("UVLT_GEMM_2SM");
  return e ? std::atoi(e) : 1;
}();
static int g_gemm_2sm_stages = [] {
  const char* e = getenv("UVLT_GEMM_2SM_STAGES");
  return e ? std::min(std::max(std::atoi(e), 2), GEMM2_MAX_STAGES) : GEMM2_MAX_STAGES;
}();

bool use_2sm(int M, int N, int K, int groups, int splits) {
  if (!g_gemm_2sm || N % GEMM2_BN || splits != 1 || groups != 1) return false;
  if (g_gemm_2sm >= 2) return true;
  if (M < 4096) return false;
  const long long m_pairs = (M + 2 * GEMM_BM - 1) / (2 * GEMM_BM);
  const long long tiles = m_pairs * (N / GEMM2_BN) * groups;
  const long long clusters = std::min<long long>(tiles, g_num_sms / 2);
  const long long rounds = (tiles + clusters - 1) / clusters;
  return K >= 2048 || 4 * tiles >= 3 * rounds * clusters;  // last-round efficiency >= 0.75
}

int pick_bn(int M, int N, int groups, bool out_f32, int act) {
  // 128 x 256 tiles halve the A-operand smem reads per FLOP and cut the L2 -> SM operand traffic by a quarter: measured
  // 862 -> 1029 TFLOP/s on the qkv GEMM at M = 16416, but slower for the fp32-output GEMMs (N = 768: only three N tiles,
  // two-pass epilogue), for the GELU epilogue (128-wide tiles at three CTAs per SM hide it better) and whenever the grid
  // is below ~2.5 waves of 2 x 148 CTAs (profiles/r01_gemm_bn256.txt)
  if (!out_f32 && act == ACT_NONE && N % 256 == 0 &&
      static_cast<long long>((M + GEMM_BM - 1) / GEMM_BM) * (N / 256) * groups >= 740)
    return 256;
  // Widest tile that still gives one CTA per SM (148) (measured on B200 at M = 513: wider tiles cut the L2 -> SM operand traffic,
  // which bounds these launches, until fewer than ~2/3 of the SMs have work); below that, the narrowest legal tile.
  const int m_tiles = (M + GEMM_BM - 1) / GEMM_BM;
  const int cands[3] = {128, 64, 32};
  int best = 0;
  for (int bn : cands) {
    if (N % bn) continue;
    best = bn;
    if (static_cast<long long>(m_tiles) * (N / bn) * groups >= 148) break;
  }
  return best;
}

int pick_splits(int M, int N, int K) {
  // measured on B200 (tools/kernel_sweep.py splitk): at M = 513 the K = 3072 GEMM takes 17 us unsplit (L2 -> SM operand
  // traffic of a few CTAs), 8.1 us cut in four and 7.6 us cut in six with BN = 64 tiles (BN = 128: 9.7 / 9.4); at
  // M = 1026 four-way is best (9.8 us) and six-way overflows the 3 x 148 resident CTAs (13.5 us)
  const int m_tiles = (M + GEMM_BM - 1) / GEMM_BM;
  if (N % 64 || K < 2048 || K % GEMM_BK) return 1;
  const int tiles = m_tiles * (N / 64);
  const int kb = K / GEMM_BK;
  int best = 1;
  for (int s : {2, 3, 4, 6})
    if (kb % s == 0 && kb / s >= 8 && tiles * s <= 3 * 148) best = s;
  return best;
}

int pick_head_splits(int M, int N, int K) {
  // the head's first conv GEMM (K = 9 * D): at max_batch 1 it is 32 BN = 64 tiles of 108 k-blocks on 148 SMs (33 us);
  // cut in nine it takes 8.2 us (+ the reduce kernel).  Smallest split that fills ~2 CTAs per SM, >= 8 k-blocks each.
  if (const char* v = std::getenv("UVLT_HEAD_SPLITS")) return std::max(1, std::atoi(v));
  if (N % 64 || K % GEMM_BK) return 1;
  const int tiles = ((M + GEMM_BM - 1) / GEMM_BM) * (N / 64);
  const int kb = K / GEMM_BK;
  int best = 1;
  for (int s = 2; s <= 12; ++s) {
    if (kb % s || kb / s < 8 || tiles * s > 3 * 148) continue;
    best = s;
    if (tiles * s >= 256) break;
  }
  return best;
}

int gemm_prepare(GemmLaunch* g, const void* A, long long a_ld, long long a_gstride, const void* W, long long w_ld,
                 long long w_gstride, int M, int N, int K, int groups, int bn, const GemmEpilogue& ep, int splits) {
  if (splits < 1) splits = 1;
  if (splits > 1) {
    if (!ep.out_f32 || !ep.split_out || K % (GEMM_BK * splits)) {
      set_error("gemm: split-K needs an fp32 output, a partial buffer and K % (64 * splits) == 0");
      return 1;
    }
    if (bn == 0) bn = 64;
  }
  g->two_sm = false;
  if (bn == 512 || (bn == 0 && use_2sm(M, N, K, groups, splits))) {
    if (N % GEMM2_BN || splits != 1 || groups != 1) {
      set_error("gemm: the CTA-pair kernel needs N % 256 == 0, one group and no split-K");
      return 1;
    }
    g->two_sm = true;
    bn = GEMM2_BN;
  }
  if (bn == 0) bn = pick_bn(M, N, groups, ep.out_f32 != 0, ep.act);
  if (bn != 32 && bn != 64 && bn != 128 && bn != 256) {
    set_error("gemm: N must be a multiple of 32");
    return 1;
  }
  if (K % GEMM_BK || N % bn || M <= 0) {
    set_error("gemm: need K % 64 == 0, N % BN == 0, M > 0 (got M=" + std::to_string(M) + " N=" + std::to_string(N) +
              " K=" + std::to_string(K) + ")");
    return 1;
  }
  // epilogue flavours the kernel is instantiated for (gemm.cuh): bf16 out = bias + activation only;
  // fp32 out = bias + optional residual + optional row remap, no activation
  if (!ep.out_f32 && (ep.resid || ep.in_rows_per_b > 0)) {
    set_error("gemm: a residual / row remap needs an fp32 output");
    return 1;
  }
  if (ep.out_f32 && ep.act != ACT_NONE) {
    set_error("gemm: activations are only fused with a bf16 output");
    return 1;
  }
  if ((ep.in_rows_per_b > 0 && ep.in_rows_per_b < 8) || (ep.resid_period > 0 && ep.resid_period < 8)) {
    set_error("gemm: row remap / residual periods must be >= 8 rows");
    return 1;
  }
  g->shape = GemmShape{M, N, K, 0, splits, groups};
  {
    const long long tiles = static_cast<long long>((M + GEMM_BM - 1) / GEMM_BM) * (N / bn) * groups * splits;
    const bool thr = tiles > 2 * 148;  // more than one wave of two CTAs per SM: throughput regime (gemm.cuh)
    const int st = bn == 32 ? GemmSmem<32>::stages_for(thr) : bn == 64 ? GemmSmem<64>::stages_for(thr)
                   : bn == 128 ? GemmSmem<128>::stages_for(thr) : GemmSmem<256>::stages_for(thr);
    g->shape.stages = g->two_sm ? g_gemm_2sm_stages : std::max(std::min(st, K / GEMM_BK / splits), 1);
  }
  g->ep = ep;
  g->bn = bn;
  g->groups = groups;
  if (groups == 1) { a_gstride = static_cast<long long>(M) * a_ld; w_gstride = static_cast<long long>(N) * w_ld; }
  if (make_tma_bf16_3d(&g->tma_a, A, K, M, groups, a_ld * 2, a_gstride * 2, GEMM_BM)) return 1;
  // pairs of CTAs along N share the A tile through TMA multicast (64-row halves) whenever the N tiles pair up
  g->multicast = !g->two_sm && g_gemm_multicast && ((N / bn) % 2 == 0);
  if (g->multicast && make_tma_bf16_3d(&g->tma_a_half, A, K, M, groups, a_ld * 2, a_gstride * 2, GEMM_BM / 2)) return 1;
  // the CTA-pair kernel stores its bf16 output tiles with TMA (128 x 64 boxes of the [M, out_ld] matrix)
  g->tma_out = g->tma_a;
  if (g->two_sm && !ep.out_f32 &&
      make_tma_bf16_3d(&g->tma_out, ep.out, N, M, 1, ep.out_ld * 2, static_cast<uint64_t>(M) * ep.out_ld * 2, GEMM_BM))
    return 1;
  // the CTA-pair kernel stages one 128-row half of the 256-wide W tile per CTA
  if (make_tma_bf16_3d(&g->tma_w, W, K, N, groups, w_ld * 2, w_gstride * 2, g->two_sm ? GEMM2_BN / 2 : bn)) return 1;
  return 0;
}

bool conv3x3_implicit_ok(int S, int Cin) {
  return S > 0 && 128 % S == 0 && (S * S) % 128 == 0 && Cin % GEMM_BK == 0;
}

int gemm_prepare_conv3x3(GemmLaunch* g, const void* src, long long src_ld, int S, int B, int Cin, const void* W,
                         long long w_gstride, int N, int groups, int bn, const GemmEpilogue& ep, int splits) {
  if (!conv3x3_implicit_ok(S, Cin)) {
    set_error("conv3x3: implicit GEMM needs 128 % S == 0, S * S % 128 == 0 and Cin % 64 == 0");
    return 1;
  }
  const int M = B * S * S, K = 9 * Cin;
  // the plain path builds every map and picks the kernel; its A map (over a fictitious [M, K] matrix at `src`) is then
  // replaced by the 4-D window map.  a_ld = K keeps the 3-D encoder's stride checks satisfied.
  if (gemm_prepare(g, src, K, static_cast<long long>(M) * K, W, K, w_gstride, M, N, K, groups, bn, ep, splits)) return 1;
  if (g->multicast) g->multicast = false;  // the multicast variant loads half boxes of A: not defined for window boxes
  const uint64_t dims[4] = {static_cast<uint64_t>(src_ld), static_cast<uint64_t>(S), static_cast<uint64_t>(S),
                            static_cast<uint64_t>(B)};
  const uint64_t strides[3] = {static_cast<uint64_t>(src_ld) * 2, static_cast<uint64_t>(S) * src_ld * 2,
                               static_cast<uint64_t>(S) * S * src_ld * 2};
  const uint32_t box[4] = {64, static_cast<uint32_t>(S), static_cast<uint32_t>(128 / S), 1};
  if (make_tma_bf16_4d(&g->tma_a, src, dims, strides, box)) return 1;
  g->shape.conv_S = S;
  g->shape.conv_cb = Cin / GEMM_BK;
  return 0;
}

template <int BN, bool MC>
static void gemm_launch_bn(const GemmLaunch& g, dim3 grid, cudaStream_t stream) {
  const size_t smem = GemmSmem<BN>::total(g.shape.stages);
  const dim3 block(GEMM_THREADS);
  const CUtensorMap& ta = MC ? g.tma_a_half : g.tma_a;
  const int cl = MC ? 2 : 1;
  if (g.ep.out_f32) {
    (void)launch_kc(gemm_bf16_tn_kernel<BN, EPI_F32, MC>, grid, block, smem, stream, cl, ta, g.tma_w, g.shape, g.ep);
  } else if (g.ep.act == ACT_GELU) {
    (void)launch_kc(gemm_bf16_tn_kernel<BN, EPI_BF16_GELU, MC>, grid, block, smem, stream, cl, ta, g.tma_w, g.shape, g.ep);
  } else if (g.ep.act == ACT_RELU) {
    (void)launch_kc(gemm_bf16_tn_kernel<BN, EPI_BF16_RELU, MC>, grid, block, smem, stream, cl, ta, g.tma_w, g.shape, g.ep);
  } else {
    (void)launch_kc(gemm_bf16_tn_kernel<BN, EPI_BF16, MC>, grid, block, smem, stream, cl, ta, g.tma_w, g.shape, g.ep);
  }
}

static void gemm_launch_2sm(const GemmLaunch& g, cudaStream_t stream) {
  // persistent: one CTA pair per TPC (or fewer, when there are fewer 256 x 256 tiles)
  const long long m_pairs = (g.shape.M + 2 * GEMM_BM - 1) / (2 * GEMM_BM);
  const long long tiles = m_pairs * (g.shape.N / GEMM2_BN) * g.groups;
  const int clusters = static_cast<int>(std::min<long long>(tiles, g_num_sms / 2));
  const dim3 grid(2 * clusters), block(GEMM2_THREADS);
#define UVLT_GEMM2_LAUNCH(EPI)                                                                                  \
  (void)launch_kc(gemm_bf16_tn_2sm_kernel<EPI>, grid, block, gemm2_smem_bytes(g.shape.stages), stream, 2, g.tma_a, \
                  g.tma_w, g.tma_out, g.shape, g.ep)
  if (g.ep.out_f32) UVLT_GEMM2_LAUNCH(EPI_F32);
  else if (g.ep.act == ACT_GELU) UVLT_GEMM2_LAUNCH(EPI_BF16_GELU);
  else if (g.ep.act == ACT_RELU) UVLT_GEMM2_LAUNCH(EPI_BF16_RELU);
  else UVLT_GEMM2_LAUNCH(EPI_BF16);
#undef UVLT_GEMM2_LAUNCH
}

int gemm_launch(const GemmLaunch& g, cudaStream_t stream) {
  if (g.two_sm) {
    gemm_launch_2sm(g, stream);
    UVLT_CUDA_OK(cudaGetLastError());
    return 0;
  }
  dim3 grid(g.shape.N / g.bn, (g.shape.M + GEMM_BM - 1) / GEMM_BM, g.groups * g.shape.splits);
  if (g.multicast) {
    switch (g.bn) {
      case 32: gemm_launch_bn<32, true>(g, grid, stream); break;
      case 64: gemm_launch_bn<64, true>(g, grid, stream); break;
      case 256: gemm_launch_bn<256, true>(g, grid, stream); break;
      default: gemm_launch_bn<128, true>(g, grid, stream); break;
    }
  } else {
    switch (g.bn) {
      case 32: gemm_launch_bn<32, false>(g, grid, stream); break;
      case 64: gemm_launch_bn<64, false>(g, grid, stream); break;
      case 256: gemm_launch_bn<256, false>(g, grid, stream); break;
      default: gemm_launch_bn<128, false>(g, grid, stream); break;
    }
  }
  UVLT_CUDA_OK(cudaGetLastError());
  return 0;
}

int attn_prepare(AttnLaunch* a, const void* qkv, int B, int n, int H, const float* bias, void* out, const void* vt,
                 int n_pad, int capacity_batch) {
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
  if (make_tma_bf16_3d(&a->tma_kv, qkv, D3, n, B, D3 * 2, static_cast<uint64_t>(n) * D3 * 2, ATT_BKV)) return 1;
  a->p.n = n;
  a->p.H = H;
  a->p.scale_log2 = 0.125f * 1.4426950408889634f;  // head_dim 64
  a->p.bias = bias;
  a->p.out = reinterpret_cast<__nv_bfloat16*>(out);
  a->B = B;
  // small grids (every CTA has an SM to itself): two CTAs per query tile, half of the key blocks each
  const int cb = capacity_batch > 0 ? capacity_batch : B;
  a->split = g_attn_split && cb * H * ((n + ATT_BQ - 1) / ATT_BQ) <= g_num_sms && (n + ATT_BKV - 1) / ATT_BKV >= 2;
  a->v2 = g_attn_v == 2;
  a->poly = g_attn_poly == 1;
  a->var3 = (g_attn_poly >= 0 && g_attn_poly <= 5) ? g_attn_poly : 2;
  a->p2.n = n;
  a->p2.H = H;
  a->p2.scale_log2 = a->p.scale_log2;
  a->p2.bias = bias;
  a->p2.out = a->p.out;
  a->p2.zero = 0;
  if (make_tma_bf16_3d(&a->tma_o, out, static_cast<uint64_t>(H) * ATT_D, n, B, static_cast<uint64_t>(H) * ATT_D * 2,
                       static_cast<uint64_t>(n) * H * ATT_D * 2, AT2_BQ))
    return 1;
  // small grids (decided from the engine's capacity, so that a sequence's result does not depend on the batch it shares
  // a call with): one query tile per CTA, the two slots split its key blocks
  a->p2.split_all = (g_attn_split && cb * H * ((n + AT2_BQ - 1) / AT2_BQ) <= g_num_sms) ? 1 : 0;
  // third generation: persistent CTAs over (batch, head, tile pair) items; only where the grid fills the GPU (decided from
  // the engine's capacity like the key split, so that a sequence's result does not depend on its batch)
  {
    const int ntiles = (n + AT3_BQ - 1) / AT3_BQ;
    const int items = B * H * ((ntiles >> 1) + (ntiles & 1));
    // automatic: where v3 measured faster (table above): grids beyond one CTA per SM, four or five query tiles
    a->v3 = !a->split && (g_attn_v == 3 || (g_attn_v == 0 && n > 384 && n <= 640));
    a->grid3 = items < g_num_sms ? items : g_num_sms;
    if (g_attn_grid_cap > 0 && a->grid3 > g_attn_grid_cap) a->grid3 = g_attn_grid_cap;
    a->p3.n = n;
    a->p3.H = H;
    a->p3.B = B;
    a->p3.scale_log2 = a->p.scale_log2;
    a->p3.bias = bias;
    a->p3.zero = 0;
    a->p3.pingpong = g_attn_pingpong;
  }
  return 0;
}

void runtime_switches(int32_t* out6) {
  out6[0] = g_pdl_enabled;
  out6[1] = g_gemm_multicast;
  out6[2] = g_gemm_2sm;
  out6[3] = g_attn_v;
  out6[4] = g_attn_split;
  out6[5] = (g_attn_poly >= 0 && g_attn_poly <= 5) ? g_attn_poly : 2;
}

int attn_launch(const AttnLaunch& a, cudaStream_t stream) {
  if (a.v3 || a.v2) return attn23_launch(a, stream);  // attention_big.cu
  dim3 grid((a.p.n + ATT_BQ - 1) / ATT_BQ, a.p.H, a.B);
  if (a.split) {
    grid.x *= 2;
    (void)launch_kc(attention_kernel<true>, grid, dim3(ATT_THREADS), AttnSmem::TOTAL_SPLIT, stream, 2, a.tma_qkv, a.tma_kv,
                    a.p);
  } else {
    UVLT_LAUNCH(attention_kernel<false>, grid, dim3(ATT_THREADS), AttnSmem::TOTAL, stream, a.tma_qkv, a.tma_kv, a.p);
  }
  UVLT_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace uvlt

#ifdef UVLT_TRACE
#include "attention_big.cu"  // trace builds: one module, one g_trace buffer
// debug builds only (make TRACE=1): copy out and reset the in-kernel timeline; out = [n][3] uint64 (tag, clk, ns)
extern "C" __attribute__((visibility("default"))) int uvlt_debug_trace(unsigned long long* out, int max_recs) {
  cudaDeviceSynchronize();
  unsigned int n = 0;
  cudaMemcpyFromSymbol(&n, uvlt::g_trace_n, sizeof(n));
  if (n > 4096u) n = 4096u;
  if (static_cast<int>(n) > max_recs) n = max_recs;
  cudaMemcpyFromSymbol(out, uvlt::g_trace, sizeof(uvlt::TraceRec) * n);
  unsigned int zero = 0;
  cudaMemcpyToSymbol(uvlt::g_trace_n, &zero, sizeof(zero));
  return static_cast<int>(n);
}
#endif
