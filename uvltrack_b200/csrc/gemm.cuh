// TMA + tcgen05 GEMM:  C[g][M,N] = epilogue( A[g][M,K] (bf16, K-major) x W[g][N,K]^T (bf16, K-major) )
//
// This one kernel serves every dense contraction of the UVLTrack forward:
//   K1 patch-embed (conv16x16/s16 as GEMM), K3 qkv, K5 out-proj (+residual), K6 fc1 (+GELU) / fc2 (+residual),
//   K8 BERT dense layers, K11 the 3x3 conv towers of the box head (im2col + grouped GEMM, BN folded, ReLU).
//
// Structure (one 128 x BN output tile per CTA, 192 threads):
//   warp 0      TMA producer: A tile [128 x 64] and W tile [BN x 64] per k-block into a STAGES-deep smem ring
//   warp 1      MMA issuer : one thread issues tcgen05.mma (M=128, N=BN, K=16) x4 per k-block, accumulator in TMEM
//   warps 2..5  epilogue   : tcgen05.ld the fp32 accumulator (thread = row), bias / GELU / ReLU / residual,
//                            vectorised global stores (bf16 or fp32)
// Shared memory is sized so two CTAs are co-resident per SM: one CTA's epilogue overlaps the other's mainloop.
#pragma once
#include "common.cuh"

namespace uvlt {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;
constexpr int GEMM_THREADS = 192;

enum { ACT_NONE = 0, ACT_GELU = 1, ACT_RELU = 2 };

struct GemmEpilogue {
  const float* bias;         // [groups][N] or nullptr
  long long bias_gstride;    // elements between groups
  const float* resid;        // fp32 residual or nullptr; row index = mapped output row, or (row % resid_period)
  long long resid_ld;        // row pitch (elements)
  int resid_period;          // >0: residual row = input row % period (e.g. positional embedding table)
  int act;                   // ACT_*
  void* out;                 // bf16 or fp32
  int out_f32;               // 1: fp32 output, 0: bf16 output
  long long out_ld;          // row pitch (elements)
  long long out_gstride;     // elements between groups (column offset for tower-concatenated outputs)
  // optional row remap: out_row = (r / in_rows_per_b) * out_rows_per_b + out_row_off + r % in_rows_per_b
  int in_rows_per_b;         // 0 = identity
  int out_rows_per_b;
  int out_row_off;
};

struct GemmShape {
  int M, N, K;
  int stages;  // depth of the smem operand ring (2..GEMM_MAX_STAGES), chosen on the host from the grid size
};

constexpr int GEMM_MAX_STAGES = 12;

template <int BN>
struct GemmSmem {
  static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;
  static constexpr int B_BYTES = BN * GEMM_BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_BYTES = 256;  // 2 * GEMM_MAX_STAGES + 1 mbarriers + the TMEM slot
  // dynamic smem for a ring of `stages`: [<=1023 B align slack][ring, 1024-aligned][barriers]
  static constexpr int total(int stages) { return stages * STAGE_BYTES + 1024 + BAR_BYTES; }
  // throughput configuration: ~96 KB of ring -> two CTAs per SM, one CTA's epilogue overlaps the other's mainloop
  static constexpr int STAGES_2CTA = (96 * 1024) / STAGE_BYTES > 8 ? 8 : (96 * 1024) / STAGE_BYTES;
  // latency configuration (grid <= one CTA per SM): as much of K in flight as fits in 227 KB
  static constexpr int STAGES_1CTA =
      (224 * 1024 - 1024 - BAR_BYTES) / STAGE_BYTES > GEMM_MAX_STAGES ? GEMM_MAX_STAGES
                                                                     : (224 * 1024 - 1024 - BAR_BYTES) / STAGE_BYTES;
};

template <int BN>
__global__ void __launch_bounds__(GEMM_THREADS, 2)
gemm_bf16_tn_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_w,
                    const GemmShape shape, const GemmEpilogue ep) {
  using S = GemmSmem<BN>;
  const int STAGES = shape.stages;
  constexpr uint32_t TMEM_COLS = BN < 32 ? 32 : BN;

  extern __shared__ uint8_t smem_raw[];
  // 128B-swizzled TMA/UMMA tiles need 1024 B alignment
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * S::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + GEMM_MAX_STAGES;
  uint64_t* acc_bar = empty_bar + GEMM_MAX_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * BN;
  const int m0 = blockIdx.y * GEMM_BM;
  const int g = blockIdx.z;
  const int num_kb = shape.K / GEMM_BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_w);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(acc_bar, 1);
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_acc = *tmem_slot;
  pdl_wait();     // everything above overlapped the predecessor's tail; A / residual are only read below
  pdl_trigger();

  if (warp == 0) {
    if (lane == 0) {
      // ---------------- TMA producer ----------------
      int s = 0;
      uint32_t ph = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&empty_bar[s], ph ^ 1);
        uint8_t* a_dst = smem + s * S::STAGE_BYTES;
        uint8_t* b_dst = a_dst + S::A_BYTES;
        mbar_expect_tx(&full_bar[s], S::STAGE_BYTES);
        tma_load_3d(a_dst, &tma_a, &full_bar[s], kb * GEMM_BK, m0, g);
        tma_load_3d(b_dst, &tma_w, &full_bar[s], kb * GEMM_BK, n0, g);
        if (++s == STAGES) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ---------------- MMA issuer ----------------
      constexpr uint32_t idesc = umma_idesc_bf16(GEMM_BM, BN, 0);
      int s = 0;
      uint32_t ph = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem + s * S::STAGE_BYTES);
        const uint32_t b_addr = a_addr + S::A_BYTES;
        const uint64_t adesc = umma_smem_desc_sw128(a_addr, 1024, 0);
        const uint64_t bdesc = umma_smem_desc_sw128(b_addr, 1024, 0);
#pragma unroll
        for (int k = 0; k < GEMM_BK / 16; ++k) {
          // advance 16 bf16 = 32 B inside the 128 B swizzle atom: +2 in the (addr >> 4) field
          umma_bf16_ss(tmem_acc, adesc + 2 * k, bdesc + 2 * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(&empty_bar[s]);  // frees the smem slot when these MMAs have drained
        if (++s == STAGES) { s = 0; ph ^= 1; }
      }
      umma_commit(acc_bar);  // accumulator complete
    }
  } else {
    // ---------------- epilogue warps ----------------
    const int lane_grp = warp & 3;  // TMEM lane quadrant this warp may access
    const int row_in_tile = lane_grp * 32 + lane;
    const int r = m0 + row_in_tile;
    mbar_wait(acc_bar, 0);
    tc_fence_after();
    const bool row_ok = r < shape.M;
    long long out_row = r;
    if (ep.in_rows_per_b > 0) {
      out_row = static_cast<long long>(r / ep.in_rows_per_b) * ep.out_rows_per_b + ep.out_row_off +
                (r % ep.in_rows_per_b);
    }
    const float* bias = ep.bias ? ep.bias + g * ep.bias_gstride + n0 : nullptr;
    const float* resid = nullptr;
    if (ep.resid) {
      const long long rr = ep.resid_period > 0 ? (r % ep.resid_period) : out_row;
      resid = ep.resid + rr * ep.resid_ld + n0;
    }
    constexpr int CH = BN < 32 ? BN : 32;
#pragma unroll 1
    for (int c = 0; c < BN; c += 32) {
      uint32_t v[32];
      tmem_ld32(tmem_acc + (static_cast<uint32_t>(lane_grp * 32) << 16) + c, v);
      tmem_wait_ld();
      if (!row_ok) continue;
      float f[32];
#pragma unroll
      for (int i = 0; i < CH; ++i) f[i] = __uint_as_float(v[i]);
      if (bias) {
#pragma unroll
        for (int i = 0; i < CH; i += 4) {
          const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias + c + i));
          f[i] += b4.x; f[i + 1] += b4.y; f[i + 2] += b4.z; f[i + 3] += b4.w;
        }
      }
      if (ep.act == ACT_GELU) {
#pragma unroll
        for (int i = 0; i < CH; ++i) f[i] = gelu_erf(f[i]);
      } else if (ep.act == ACT_RELU) {
#pragma unroll
        for (int i = 0; i < CH; ++i) f[i] = fmaxf(f[i], 0.0f);
      }
      if (resid) {
#pragma unroll
        for (int i = 0; i < CH; i += 4) {
          const float4 r4 = *reinterpret_cast<const float4*>(resid + c + i);
          f[i] += r4.x; f[i + 1] += r4.y; f[i + 2] += r4.z; f[i + 3] += r4.w;
        }
      }
      if (ep.out_f32) {
        float* o = reinterpret_cast<float*>(ep.out) + g * ep.out_gstride + out_row * ep.out_ld + n0 + c;
#pragma unroll
        for (int i = 0; i < CH; i += 4)
          *reinterpret_cast<float4*>(o + i) = make_float4(f[i], f[i + 1], f[i + 2], f[i + 3]);
      } else {
        __nv_bfloat16* o =
            reinterpret_cast<__nv_bfloat16*>(ep.out) + g * ep.out_gstride + out_row * ep.out_ld + n0 + c;
#pragma unroll
        for (int i = 0; i < CH; i += 8) {
          uint4 u;
          u.x = pack_bf16x2(f[i], f[i + 1]);
          u.y = pack_bf16x2(f[i + 2], f[i + 3]);
          u.z = pack_bf16x2(f[i + 4], f[i + 5]);
          u.w = pack_bf16x2(f[i + 6], f[i + 7]);
          *reinterpret_cast<uint4*>(o + i) = u;
        }
      }
    }
    tc_fence_before();
  }

  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_acc, TMEM_COLS);
  }
}

}  // namespace uvlt
