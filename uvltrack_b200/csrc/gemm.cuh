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
// Shared memory is sized so two (latency regime) or three (throughput regime) CTAs are co-resident per SM: one CTA's
// epilogue overlaps the others' mainloops.
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
  // split-K (GemmShape::splits > 1, fp32 flavour only): split 0 writes `out` (bias + residual as usual), split s > 0
  // writes its raw partial product to split_out + (s - 1) * split_stride at the same [row, col] offsets; the consumer
  // (the LayerNorm that follows, rowwise.cuh) adds the partials in a fixed order -> deterministic, no atomics
  float* split_out;
  long long split_stride;
};

struct GemmShape {
  int M, N, K;
  int stages;  // depth of the smem operand ring (2..GEMM_MAX_STAGES), chosen on the host from the grid size
  int splits;  // K is cut into `splits` equal ranges, one CTA each (blockIdx.z = group * splits + split)
  int groups;  // read by the persistent CTA-pair kernel only (the one-tile kernels take the group from blockIdx.z)
  // Implicit-GEMM 3x3 / pad 1 convolution over an S x S token grid (conv_S > 0): A is NOT an [M, K] matrix but the bf16
  // feature map [B, S, S, groups * Cin] behind a 4-D tensor map with box {64 channels, S, 128 / S rows, 1}.  k-block kb
  // covers tap = kb / conv_cb (ky = tap / 3, kx = tap % 3) and channels [(kb % conv_cb) * 64, +64) of group g; the 128
  // output rows of an M tile are 128 / S whole image rows of one sequence, and the tap shifts the box by (kx - 1, ky - 1):
  // out-of-range coordinates are the zero padding (TMA zero fill).  W keeps the (ky, kx, c) column order of the im2col
  // formulation, so K = 9 * Cin and everything but the A load is unchanged.  Requires 128 % S == 0 and S * S % 128 == 0.
  int conv_S;
  int conv_cb;  // Cin / 64
};

constexpr int GEMM_MAX_STAGES = 12;

template <int BN>
struct GemmSmem {
  static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;
  static constexpr int B_BYTES = BN * GEMM_BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_BYTES = 256;   // 2 * GEMM_MAX_STAGES + 1 mbarriers + the TMEM slot
  static constexpr int BIAS_BYTES = BN * 4;
  // staging tile of the epilogue (aliases the operand ring): fp32 flavour in passes of <= 128 columns, bf16 whole tile
  static constexpr int EPI_BYTES = GEMM_BM * (BN > 128 ? 128 : BN) * 4;
  // dynamic smem for a ring of `stages`: [ring, 1024-aligned][barriers][bias]
  static constexpr int total(int stages) {
    return (stages * STAGE_BYTES > EPI_BYTES ? stages * STAGE_BYTES : EPI_BYTES) + BAR_BYTES + BIAS_BYTES;
  }
  // Ring depth.  Latency regime (grid <= one wave of two CTAs per SM): ~100 KB, as much of K in flight as two
  // co-resident CTAs allow, and under programmatic dependent launch the next kernel's CTAs become resident (and prefetch
  // their weight tiles) while this one runs.  Throughput regime: ~70 KB so that THREE CTAs share an SM -- the epilogue
  // of a tile is a latency chain of ~2.7k cycles on four lone warps, and a third CTA hides more of it
  // (measured at M = 16416: proj 43.5 -> 36.4 us, fc1 116 -> 100.5 us, fc2 92.6 -> 87.4 us).
  static constexpr int stages_for(bool throughput) {
    const int budget = (throughput && BN <= 128 ? 72 : 108) * 1024;
    return budget / STAGE_BYTES > 8 ? 8 : budget / STAGE_BYTES;
  }
  static constexpr int STAGES_2CTA = stages_for(false);
};

// Epilogue flavours (compile-time, so that each instantiation carries only its own code: these kernels run every code
// path once per CTA, i.e. with a cold instruction cache at batch 1 -- code size is latency here).
enum { EPI_BF16 = 0, EPI_BF16_GELU = 1, EPI_BF16_RELU = 2, EPI_F32 = 3 };
//   EPI_BF16*  : out bf16 [M, out_ld] (+ group column offset), bias, optional activation
//   EPI_F32    : out fp32 with the optional row remap into the [B, N, D] residual stream, bias, optional fp32 residual
//                (same-row in-place residual, or a periodic table such as the positional embedding)

__device__ __forceinline__ float4 gelu4(float4 v) {
  gelu_erf_x2(v.x, v.y);
  gelu_erf_x2(v.z, v.w);
  return v;
}

// Epilogue of one 128 x BN accumulator tile, run by the four epilogue warps (warp & 3 = TMEM lane quadrant).
// Stage 1 (thread = accumulator row): TMEM -> registers, + bias (+ activation, -> bf16) -> staging tile in the (now
// idle) operand ring, 16-byte chunks XOR-swizzled with the row so that both stages are bank-conflict free.
// Stage 2 (lanes across columns): coalesced residual read / add, store -- whole rows per warp instruction (a
// row-per-thread store costs 32 L1 wavefronts per instruction instead of 2-4).
// STAGE1_ONLY: stop after the staging tile is written (the caller stores it with one TMA bulk store; bf16 flavours: the
// staging layout [128 rows][BN = 64 bf16], 16-byte chunks XOR-swizzled with row % 8, IS the 128B-swizzle TMA layout).
template <int BN, int EPI, bool STAGE1_ONLY = false>
__device__ __forceinline__ void gemm_epilogue(uint8_t* smem, const float* s_bias, uint64_t* acc_bar, uint32_t acc_parity,
                                              uint32_t tmem_acc, int m0, int n0, int g, int sp, const GemmShape& shape,
                                              const GemmEpilogue& ep, int warp, int lane TRACE_PARAMS) {
  const int lane_grp = warp & 3;  // TMEM lane quadrant this warp may access
  constexpr bool F32 = (EPI == EPI_F32);
  constexpr int ESZ = F32 ? 4 : 2;            // staged element size: the bf16 flavours convert in stage 1
  constexpr int PW = (F32 && BN > 128) ? 128 : BN;  // columns per epilogue pass (the staging tile is <= 64 KB)
  constexpr int NPASS = BN / PW;
  constexpr int ROWB = PW * ESZ;              // bytes per staged row
  constexpr int CPR = ROWB / 16;              // 16-byte chunks per staged row
  constexpr int RPI = 32 / CPR;               // rows covered by one warp instruction in stage 2
  constexpr int ITERS = 32 / RPI;
  constexpr int SWZ = (CPR < 8 ? CPR : 8) - 1;  // XOR swizzle of the chunk index with the row, kept inside the row
  const uint32_t stage_w = smem_u32(smem) + lane_grp * 32 * ROWB;  // this warp's 32 rows of the staging tile
  const int j2 = lane % CPR;
  const int col0 = n0 + j2 * (16 / ESZ);
  const int r_first = m0 + lane_grp * 32 + lane / CPR;  // this lane's rows are r_first + it * RPI
  // position of the first row inside its sequence / inside the periodic residual table, computed while the mainloop
  // runs; advanced incrementally afterwards (the periods are >= 8 rows on this path, checked on the host)
  int seq_q0 = 0, seq_rem0 = r_first, per_rem0 = 0;
  if (F32) {
    if (ep.in_rows_per_b > 0) { seq_q0 = r_first / ep.in_rows_per_b; seq_rem0 = r_first - seq_q0 * ep.in_rows_per_b; }
    if (ep.resid_period > 0) per_rem0 = r_first % ep.resid_period;
  }
  // fp32 flavour: the residual rows of this tile do not depend on the accumulator -- fetch (up to eight row groups of) them
  // BEFORE waiting for it.  ncu on the CTA-pair out-proj GEMM (M = 17696, K = 768): 53 % of all warp stalls were the
  // long-scoreboard wait on exactly these loads, issued four at a time after the accumulator had landed
  // (profiles/r02_b32_gemm_full.md); the kernel ran at 632 TFLOP/s with DRAM at 26 % and the tensor pipe at 32 %.
  constexpr int NPRE = (F32 && NPASS == 1 && ITERS <= 8) ? ITERS : 0;  // 32-column passes (wider ones would spill)
  float4 rpre[NPRE > 0 ? NPRE : 1];
  if (NPRE > 0) {
    int q = seq_q0, rem = seq_rem0, per = per_rem0;
    const int rows_left = shape.M - r_first;
#pragma unroll
    for (int it = 0; it < NPRE; ++it) {
      long long rr;
      if (ep.in_rows_per_b > 0) {
        if (rem >= ep.in_rows_per_b) { rem -= ep.in_rows_per_b; ++q; }
        rr = static_cast<long long>(q) * ep.out_rows_per_b + ep.out_row_off + rem;
      } else {
        rr = rem;
      }
      rem += RPI;
      if (ep.resid_period > 0) {
        if (per >= ep.resid_period) per -= ep.resid_period;
        rr = per;
        per += RPI;
      }
      rpre[it] = (ep.resid && it * RPI < rows_left && sp == 0)
                     ? *reinterpret_cast<const float4*>(ep.resid + rr * ep.resid_ld + col0)
                     : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  mbar_wait(acc_bar, acc_parity);
  tc_fence_after();
  if (threadIdx.x == 64) TRACE_PT(0x105);
#pragma unroll 1
  for (int pass = 0; pass < NPASS; ++pass) {
  const int col = col0 + pass * PW;
  int seq_q = seq_q0, seq_rem = seq_rem0, per_rem = per_rem0;
  {
    const uint32_t t_row = tmem_acc + (static_cast<uint32_t>(lane_grp * 32) << 16) + pass * PW;
    const uint32_t my_row = stage_w + lane * ROWB;
    uint32_t va[32], vb[32];
    // 32 accumulator columns: + bias (+ activation, -> bf16) -> swizzled staging row
    auto stage1 = [&](uint32_t (&v)[32], int c) {
      if (F32) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 b4 = *reinterpret_cast<const float4*>(s_bias + pass * PW + c + q * 4);
          const int j = ((c >> 2) + q) ^ (lane & SWZ);
          sts128(my_row + j * 16, __uint_as_float(v[q * 4]) + b4.x, __uint_as_float(v[q * 4 + 1]) + b4.y,
                 __uint_as_float(v[q * 4 + 2]) + b4.z, __uint_as_float(v[q * 4 + 3]) + b4.w);
        }
      } else {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float4 x0 = *reinterpret_cast<const float4*>(s_bias + pass * PW + c + q * 8);
          float4 x1 = *reinterpret_cast<const float4*>(s_bias + pass * PW + c + q * 8 + 4);
          x0.x += __uint_as_float(v[q * 8]);     x0.y += __uint_as_float(v[q * 8 + 1]);
          x0.z += __uint_as_float(v[q * 8 + 2]); x0.w += __uint_as_float(v[q * 8 + 3]);
          x1.x += __uint_as_float(v[q * 8 + 4]); x1.y += __uint_as_float(v[q * 8 + 5]);
          x1.z += __uint_as_float(v[q * 8 + 6]); x1.w += __uint_as_float(v[q * 8 + 7]);
          if (EPI == EPI_BF16_GELU) { x0 = gelu4(x0); x1 = gelu4(x1); }
          if (EPI == EPI_BF16_RELU) {
            x0.x = fmaxf(x0.x, 0.f); x0.y = fmaxf(x0.y, 0.f); x0.z = fmaxf(x0.z, 0.f); x0.w = fmaxf(x0.w, 0.f);
            x1.x = fmaxf(x1.x, 0.f); x1.y = fmaxf(x1.y, 0.f); x1.z = fmaxf(x1.z, 0.f); x1.w = fmaxf(x1.w, 0.f);
          }
          const int j = ((c >> 3) + q) ^ (lane & SWZ);
          sts128(my_row + j * 16, __uint_as_float(pack_bf16x2(x0.x, x0.y)), __uint_as_float(pack_bf16x2(x0.z, x0.w)),
                 __uint_as_float(pack_bf16x2(x1.x, x1.y)), __uint_as_float(pack_bf16x2(x1.z, x1.w)));
        }
      }
    };
    tmem_ld32(t_row, va);
    tmem_wait_ld_dep(va);
#pragma unroll 1
    for (int c = 0; c < PW; c += 64) {
      if (c + 32 < PW) tmem_ld32(t_row + c + 32, vb);
      stage1(va, c);
      if (c + 32 < PW) {
        tmem_wait_ld_dep(vb);
        if (c + 64 < PW) tmem_ld32(t_row + c + 64, va);
        stage1(vb, c + 32);
        if (c + 64 < PW) tmem_wait_ld_dep(va);
      }
    }
  }
  tc_fence_before();
  __syncwarp();
  if (threadIdx.x == 64) TRACE_PT(0x108);
  if (STAGE1_ONLY) return;
  {
    const int rows_left = shape.M - r_first;
    const uint32_t src = stage_w + (lane / CPR) * ROWB;
    constexpr int UN = ITERS < 4 ? ITERS : 4;
    // fully unrolled when residual rows were prefetched (rpre must be indexed statically to stay in registers)
#pragma unroll(NPRE > 0 ? 16 : 1)
    for (int it0 = 0; it0 < ITERS; it0 += UN) {
      float4 v[UN];
#pragma unroll
      for (int u = 0; u < UN; ++u) {
        const int row_l = (it0 + u) * RPI + lane / CPR;
        v[u] = lds128(src + (it0 + u) * RPI * ROWB + ((j2 ^ (row_l & SWZ)) * 16));
      }
      if (F32) {
        long long orow[UN];
        float4 r4[UN];
#pragma unroll
        for (int u = 0; u < UN; ++u) {
          const bool ok = (it0 + u) * RPI < rows_left;
          if (ep.in_rows_per_b > 0) {
            if (seq_rem >= ep.in_rows_per_b) { seq_rem -= ep.in_rows_per_b; ++seq_q; }
            orow[u] = static_cast<long long>(seq_q) * ep.out_rows_per_b + ep.out_row_off + seq_rem;
          } else {
            orow[u] = seq_rem;
          }
          seq_rem += RPI;
          long long rr = orow[u];
          if (ep.resid_period > 0) {
            if (per_rem >= ep.resid_period) per_rem -= ep.resid_period;
            rr = per_rem;
            per_rem += RPI;
          }
          if (NPRE > 0 && it0 + u < NPRE)
            r4[u] = rpre[it0 + u];
          else
            r4[u] = (ep.resid && ok && sp == 0) ? *reinterpret_cast<const float4*>(ep.resid + rr * ep.resid_ld + col)
                                                : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < UN; ++u) {
          if ((it0 + u) * RPI >= rows_left) continue;
          v[u].x += r4[u].x; v[u].y += r4[u].y; v[u].z += r4[u].z; v[u].w += r4[u].w;
          float* const obase = sp == 0 ? reinterpret_cast<float*>(ep.out) : ep.split_out + (sp - 1) * ep.split_stride;
          *reinterpret_cast<float4*>(obase + g * ep.out_gstride + orow[u] * ep.out_ld + col) = v[u];
        }
      } else {
        __nv_bfloat16* const o = reinterpret_cast<__nv_bfloat16*>(ep.out) + g * ep.out_gstride + col;
#pragma unroll
        for (int u = 0; u < UN; ++u) {
          if ((it0 + u) * RPI >= rows_left) continue;
          *reinterpret_cast<float4*>(o + static_cast<long long>(r_first + (it0 + u) * RPI) * ep.out_ld) = v[u];
        }
      }
    }
  }
  __syncwarp();  // stage 2 has read this pass before the next pass overwrites the staging rows
  }  // pass
}

// MC (launched as clusters of two CTAs along N): the two CTAs compute neighbouring N tiles of the same M tile, so they
// need the same A tile.  Each CTA fetches one 64-row half of it and TMA-multicasts the half into both CTAs' rings: the
// L2 -> SM operand traffic, which bounds this kernel at every batch size (12 TB/s at 128x128 tiles), drops by a quarter.
// A ring slot may be refilled only when BOTH CTAs have consumed it: the MMA commit arrives on both CTAs' empty barrier.
template <int BN, int EPI, bool MC>
__global__ void __launch_bounds__(GEMM_THREADS, 3)
gemm_bf16_tn_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_w,
                    const GemmShape shape, const GemmEpilogue ep) {
  using S = GemmSmem<BN>;
  const int STAGES = shape.stages;
  constexpr uint32_t TMEM_COLS = BN < 32 ? 32 : BN;

  extern __shared__ __align__(1024) uint8_t gemm_smem[];  // 128B-swizzled TMA/UMMA tiles need 1024 B alignment
  uint8_t* const smem = gemm_smem;
  const int ring_bytes = STAGES * S::STAGE_BYTES > S::EPI_BYTES ? STAGES * S::STAGE_BYTES : S::EPI_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + ring_bytes);
  uint64_t* empty_bar = full_bar + GEMM_MAX_STAGES;
  uint64_t* acc_bar = empty_bar + GEMM_MAX_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_bar + 1);
  float* s_bias = reinterpret_cast<float*>(smem + ring_bytes + S::BAR_BYTES);

  const int warp = uniform_warp_idx();
  const int lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * BN;
  const int m0 = blockIdx.y * GEMM_BM;
  const int g = blockIdx.z / shape.splits;
  const int sp = blockIdx.z - g * shape.splits;
  const int num_kb = shape.K / GEMM_BK / shape.splits;
  const int kb0 = sp * num_kb;  // first k-block of this split
  const int pre = num_kb < STAGES ? num_kb : STAGES;  // k-blocks whose weight tile is requested before pdl_wait
  const uint32_t crank = MC ? cluster_ctarank() : 0u;
  constexpr int A_HALF = S::A_BYTES / 2;
  TRACE_DECL;
  if (threadIdx.x == 0) TRACE_PT(0x100);

  // ---- prologue: touches only weights (W tiles, bias), so under PDL it overlaps the predecessor kernel ----
  if (warp == 0) {
    if (lane == 0) {
      if (smem_u32(smem) & 1023u) __trap();  // dynamic shared memory must be 1024-byte aligned
      tma_prefetch_desc(&tma_a);
      tma_prefetch_desc(&tma_w);
#pragma unroll 1
      for (int s = 0; s < STAGES; ++s) {
        mbar_init(&full_bar[s], 1);
        mbar_init(&empty_bar[s], MC ? 2 : 1);
      }
      mbar_init(acc_bar, 1);
      fence_mbar_init();
    }
    __syncwarp();
    if (elect_one_sync()) {
#pragma unroll 1
      for (int kb = 0; kb < pre; ++kb) {
        mbar_expect_tx(&full_bar[kb], S::STAGE_BYTES);  // A + W bytes; the A half is issued after pdl_wait
        tma_load_3d(smem + kb * S::STAGE_BYTES + S::A_BYTES, &tma_w, &full_bar[kb], (kb0 + kb) * GEMM_BK, n0, g);
      }
    }
    __syncwarp();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  if (threadIdx.x >= 64) {
    for (int i = threadIdx.x - 64; i < BN; i += GEMM_THREADS - 64)
      s_bias[i] = (ep.bias && sp == 0) ? __ldg(ep.bias + g * ep.bias_gstride + n0 + i) : 0.0f;
  }
  tc_fence_before();
  __syncthreads();
  if (MC) cluster_sync_all();  // the peer's barriers must be initialised before our multicast can signal them
  tc_fence_after();
  const uint32_t tmem_acc = *tmem_slot;
  if (threadIdx.x == 0) TRACE_PT(0x101);
  pdl_wait();     // A and the residual come from earlier kernels of the chain: nothing below may move above this
  pdl_trigger();
  if (threadIdx.x == 0) TRACE_PT(0x102);

  // A tile of k-block `kbg` (global k-block index): a [128 x 64] box of the [M, K] matrix, or in convolution mode the
  // window of the feature map shifted by the tap of this k-block (see GemmShape::conv_S)
  auto load_a = [&](uint8_t* dst, uint64_t* bar, int kbg) {
    if (shape.conv_S > 0) {
      const int tap = kbg / shape.conv_cb;
      const int c0 = (g * shape.conv_cb + (kbg - tap * shape.conv_cb)) * GEMM_BK;
      const int SS = shape.conv_S * shape.conv_S;
      const int b = m0 / SS;
      const int y0 = (m0 - b * SS) / shape.conv_S;
      tma_load_4d(dst, &tma_a, bar, c0, tap % 3 - 1, y0 + tap / 3 - 1, b);
    } else {
      tma_load_3d(dst, &tma_a, bar, kbg * GEMM_BK, m0, g);
    }
  };
  // The producer and MMA warps stay CONVERGED (uniform loop counters, all lanes wait on the barriers) and only the
  // asynchronous instruction is issued by one elected lane: under `if (lane == 0)` ptxas wraps every tcgen05.mma / TMA
  // instruction in an ELECT + R2UR.BROADCAST waterfall loop (see elect_one_sync in common.cuh).
  if (warp == 0) {
    {
      // ---------------- TMA producer ----------------
      if (elect_one_sync()) {
#pragma unroll 1
        for (int kb = 0; kb < pre; ++kb) {
          if (MC)
            tma_load_3d_mc(smem + kb * S::STAGE_BYTES + crank * A_HALF, &tma_a, &full_bar[kb], (kb0 + kb) * GEMM_BK,
                           m0 + crank * (GEMM_BM / 2), g, 0x3);
          else
            load_a(smem + kb * S::STAGE_BYTES, &full_bar[kb], kb0 + kb);
        }
      }
      __syncwarp();
      int s = 0;              // pre == STAGES whenever the loop below runs
      uint32_t ph = 0;
#pragma unroll 1
      for (int kb = pre; kb < num_kb; ++kb) {
        mbar_wait(&empty_bar[s], ph);
        if (elect_one_sync()) {
          uint8_t* a_dst = smem + s * S::STAGE_BYTES;
          mbar_expect_tx(&full_bar[s], S::STAGE_BYTES);
          if (MC)
            tma_load_3d_mc(a_dst + crank * A_HALF, &tma_a, &full_bar[s], (kb0 + kb) * GEMM_BK, m0 + crank * (GEMM_BM / 2), g,
                           0x3);
          else
            load_a(a_dst, &full_bar[s], kb0 + kb);
          tma_load_3d(a_dst + S::A_BYTES, &tma_w, &full_bar[s], (kb0 + kb) * GEMM_BK, n0, g);
        }
        __syncwarp();
        if (++s == STAGES) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    {
      // ---------------- MMA issuer ----------------
      constexpr uint32_t idesc = umma_idesc_bf16(GEMM_BM, BN, 0);
      int s = 0;
      uint32_t ph = 0;
#pragma unroll 1
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        if (kb == 0 && lane == 0) TRACE_PT(0x103);
        const uint32_t a_addr = smem_u32(smem + s * S::STAGE_BYTES);
        const uint32_t b_addr = a_addr + S::A_BYTES;
        const uint64_t adesc = umma_smem_desc_sw128(a_addr, 1024, 0);
        const uint64_t bdesc = umma_smem_desc_sw128(b_addr, 1024, 0);
        if (elect_one_sync()) {
#pragma unroll
          for (int k = 0; k < GEMM_BK / 16; ++k) {
            // advance 16 bf16 = 32 B inside the 128 B swizzle atom: +2 in the (addr >> 4) field
            umma_bf16_ss(tmem_acc, adesc + 2 * k, bdesc + 2 * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
          }
          // frees the smem slot (in both CTAs of the cluster) when these MMAs have drained
          if (MC) umma_commit_mc(&empty_bar[s], 0x3);
          else umma_commit(&empty_bar[s]);
          if (kb == num_kb - 1) umma_commit(acc_bar);  // accumulator complete
        }
        __syncwarp();
        if (++s == STAGES) { s = 0; ph ^= 1; }
      }
      if (lane == 0) TRACE_PT(0x104);
    }
  } else {
    // ---------------- epilogue warps ----------------
    gemm_epilogue<BN, EPI>(smem, s_bias, acc_bar, 0, tmem_acc, m0, n0, g, sp, shape, ep, warp, lane TRACE_ARGS);
    if (threadIdx.x == 64) TRACE_PT(0x106);
  }

  __syncthreads();
  if (MC) cluster_sync_all();  // the peer's last commits / multicasts target this CTA's shared memory
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_acc, TMEM_COLS);
  }
  if (threadIdx.x == 64) TRACE_PT(0x107);
  if (threadIdx.x == 0 || threadIdx.x == 32 || threadIdx.x == 64) TRACE_FLUSH();
}

// ---------------------------------------------------------------------------------------------------------------
// CTA-pair persistent variant (tcgen05 cta_group::2) for the throughput regime.
//
// A cluster of two CTAs on the two SMs of a TPC computes 256 x 256 output tiles, one after the other (static
// round-robin over the grid's clusters).  Each CTA stages its own 128 rows of A and ONE HALF (128 rows) of the W tile
// per k-block; the leader's single thread issues tcgen05.mma.cta_group::2 (M = 256, N = 256), which reads both CTAs'
// shared memory and writes both CTAs' TMEM (128 lanes x 256 columns each): half the L2 -> SM operand bytes per FLOP of
// the one-CTA 128 x 128 tile.  The accumulator is double-buffered in TMEM (2 x 256 columns = all of it), so the sixteen
// epilogue warps drain tile i while the producer / MMA warps are already in the mainloop of tile i + 1, and the
// per-tile set-up / tear-down latency chains of the one-tile-per-CTA kernel (barrier init, TMEM allocation, pipeline
// fill, cluster sync: ~8 us per 128 x 256 tile against a 3.3 us mainloop, in-kernel timeline in
// profiles/r01_gemm_2sm.md) are paid once per kernel.
//   full_bar[s]      in the LEADER: its producer posts expect_tx for both CTAs' bytes, both CTAs' TMA loads signal it
//   empty_bar[s]     in each CTA: the leader's tcgen05.commit multicasts the arrive to both
//   acc_full[b]      in each CTA: multicast commit after the tile's last MMA
//   acc_empty[b]     in the LEADER: all epilogue warps of the pair arrive once their TMEM reads are done
// ---------------------------------------------------------------------------------------------------------------
constexpr int GEMM2_BN = 256;       // columns of the pair tile = accumulator columns per CTA and tile
constexpr int GEMM2_EPI_GROUPS = 4;  // epilogue warp groups (4 warps = the 4 TMEM lane quadrants), 256 / GROUPS columns each
constexpr int GEMM2_THREADS = 64 + GEMM2_EPI_GROUPS * 128;  // warp 0 TMA, warp 1 MMA, then the epilogue groups
constexpr int GEMM2_STAGE_BYTES = GEMM_BM * GEMM_BK * 2 + (GEMM2_BN / 2) * GEMM_BK * 2;  // A + half of W: 32 KB
// each epilogue group stages its 128 x 64 sub-tile (bf16; fp32 in two 32-column passes) in its own 16 KB
constexpr int GEMM2_GROUP_COLS = GEMM2_BN / GEMM2_EPI_GROUPS;
constexpr int GEMM2_EPI_BYTES = GEMM2_EPI_GROUPS * GEMM_BM * GEMM2_GROUP_COLS * 2;
constexpr int GEMM2_MAX_STAGES = 5;
// ring + staging tiles + barriers + a zero bias row + per epilogue group two 64-float bias rows (double-buffered by tile)
constexpr int GEMM2_BIAS_BYTES = GEMM2_EPI_GROUPS * 2 * GEMM2_GROUP_COLS * 4;
constexpr int gemm2_smem_bytes(int stages) {
  return stages * GEMM2_STAGE_BYTES + GEMM2_EPI_BYTES + 256 + 128 * 4 + GEMM2_BIAS_BYTES;
}
static_assert(gemm2_smem_bytes(GEMM2_MAX_STAGES) <= 227 * 1024, "CTA-pair GEMM: shared memory budget");

template <int EPI>
__global__ void __launch_bounds__(GEMM2_THREADS, 1)
gemm_bf16_tn_2sm_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_w,
                        const __grid_constant__ CUtensorMap tma_out, const GemmShape shape, const GemmEpilogue ep) {
  constexpr int BN = GEMM2_BN;
  constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;
  constexpr int STAGE_BYTES = GEMM2_STAGE_BYTES;
  const int STAGES = shape.stages;

  extern __shared__ __align__(1024) uint8_t gemm_smem[];
  uint8_t* const smem = gemm_smem;
  uint8_t* const stage_base = smem + STAGES * STAGE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(stage_base + GEMM2_EPI_BYTES);
  uint64_t* empty_bar = full_bar + GEMM2_MAX_STAGES;
  uint64_t* acc_full = empty_bar + GEMM2_MAX_STAGES;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  float* s_zero = reinterpret_cast<float*>(stage_base + GEMM2_EPI_BYTES + 256);  // bias of a bias-less GEMM (unused)
  float* s_bias2 = s_zero + 128;  // [groups][2][64]: the current / next tile's bias slice of each epilogue group

  const int warp = uniform_warp_idx();
  const int lane = threadIdx.x & 31;
  const uint32_t crank = cluster_ctarank();  // 0 = leader (issues the MMAs)
  const int num_kb = shape.K / GEMM_BK;
  const int m_pairs = (shape.M + 2 * GEMM_BM - 1) / (2 * GEMM_BM);
  const int n_tiles = shape.N / BN;
  const int tiles_per_group = m_pairs * n_tiles;
  const int total_tiles = tiles_per_group * shape.groups;
  const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
  TRACE_DECL;
  if (threadIdx.x == 0) TRACE_PT(0x100);

  if (warp == 0 && lane == 0) {
    if (smem_u32(smem) & 1023u) __trap();
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_w);
#pragma unroll 1
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&acc_full[0], 1);
    mbar_init(&acc_full[1], 1);
    mbar_init(&acc_empty[0], 2 * 4 * GEMM2_EPI_GROUPS);
    mbar_init(&acc_empty[1], 2 * 4 * GEMM2_EPI_GROUPS);
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc_2sm(tmem_slot, 2 * BN);
    tmem_relinquish_2sm();
  }
  if (threadIdx.x < 128) s_zero[threadIdx.x] = 0.0f;
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // both CTAs' barriers exist before any TMA / commit can signal them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) TRACE_PT(0x101);
  pdl_wait();
  pdl_trigger();
  if (threadIdx.x == 0) TRACE_PT(0x102);

  // tile t -> (group, m pair, n tile): n fastest, so the ~74 clusters running at the same time cover a BAND of M with
  // every N tile of it: each A tile is fetched from HBM once and re-read from L2 by the clusters of the other N tiles,
  // and the whole weight matrix (<= 4.7 MB) stays L2 resident.  With m fastest (round 1) one wave streamed the full A
  // operand per N tile: fc2 at M = 17696 read its 108.7 MB A operand three times (ncu dram__bytes_read 318.6 MB against
  // 167.8 MB algorithmic, profiles/r01_final_b32_gemm_full.md).
  auto tile_coords = [&](int t, int& g, int& m0, int& n0) {
    g = t / tiles_per_group;
    const int r = t - g * tiles_per_group;
    const int mp = r / n_tiles;
    m0 = (mp * 2 + static_cast<int>(crank)) * GEMM_BM;
    n0 = (r - mp * n_tiles) * BN;
  };

  if (warp == 0) {
    {
      // ---------------- TMA producer (both CTAs; converged warp, one elected lane issues) ----------------
      int s = 0;
      uint32_t ph = 1;  // fresh barriers: the first pass over the ring does not wait
#pragma unroll 1
      for (int t = cluster_id; t < total_tiles; t += n_clusters) {
        int g, m0, n0;
        tile_coords(t, g, m0, n0);
#pragma unroll 1
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[s], ph);  // the leader's MMAs have drained this slot in BOTH CTAs
          if (elect_one_sync()) {
            uint8_t* a_dst = smem + s * STAGE_BYTES;
            if (crank == 0) mbar_expect_tx(&full_bar[s], 2 * STAGE_BYTES);
            if (shape.conv_S > 0) {  // implicit-GEMM convolution: shifted window of the feature map (GemmShape::conv_S)
              const int tap = kb / shape.conv_cb;
              const int SS = shape.conv_S * shape.conv_S;
              const int b = m0 / SS;
              tma_load_4d_2sm(a_dst, &tma_a, &full_bar[s], (g * shape.conv_cb + (kb - tap * shape.conv_cb)) * GEMM_BK,
                              tap % 3 - 1, (m0 - b * SS) / shape.conv_S + tap / 3 - 1, b);
            } else {
              tma_load_3d_2sm(a_dst, &tma_a, &full_bar[s], kb * GEMM_BK, m0, g);
            }
            tma_load_3d_2sm(a_dst + A_BYTES, &tma_w, &full_bar[s], kb * GEMM_BK, n0 + static_cast<int>(crank) * (BN / 2), g);
          }
          __syncwarp();
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (crank == 0) {
      // ---------------- MMA issuer (leader only; converged warp, one elected lane issues) ----------------
      constexpr uint32_t idesc = umma_idesc_bf16(2 * GEMM_BM, BN, 0);
      int s = 0;
      uint32_t ph = 0;
      int it = 0;
#pragma unroll 1
      for (int t = cluster_id; t < total_tiles; t += n_clusters, ++it) {
        const int buf = it & 1;
        mbar_wait(&acc_empty[buf], ((it >> 1) & 1) ^ 1);  // both CTAs' epilogues have drained this accumulator
        tc_fence_after();
        const uint32_t tmem_acc = tmem_base + buf * BN;
#pragma unroll 1
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          if (it == 0 && kb < 2 && lane == 0) TRACE_PT(0x103);
          const uint32_t a_addr = smem_u32(smem + s * STAGE_BYTES);
          const uint64_t adesc = umma_smem_desc_sw128(a_addr, 1024, 0);
          const uint64_t bdesc = umma_smem_desc_sw128(a_addr + A_BYTES, 1024, 0);
          if (elect_one_sync()) {
#pragma unroll
            for (int k = 0; k < GEMM_BK / 16; ++k)
              umma_bf16_ss_2sm(tmem_acc, adesc + 2 * k, bdesc + 2 * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
            umma_commit_2sm(&empty_bar[s], 0x3);
            if (kb == num_kb - 1) umma_commit_2sm(&acc_full[buf], 0x3);
          }
          __syncwarp();
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
        if (it < 4 && lane == 0) TRACE_PT(0x104);
      }
    }
  } else {
    // ---------------- epilogue: group q (4 warps) owns columns [q * 64, +64) of the tile ----------------
    constexpr int GC = GEMM2_GROUP_COLS;
    const int half = (warp - 2) >> 2;
    uint8_t* const stage = stage_base + half * (GEMM2_EPI_BYTES / GEMM2_EPI_GROUPS);
    const uint32_t leader_empty = smem_u32(acc_empty) & 0xFEFFFFFFu;  // the pair's even CTA (see tma_load_3d_2sm)
    // bf16 flavours: the group's 128 x 64 staging tile goes out with ONE TMA bulk store issued by the group's first
    // thread -- with sixteen warps pushing row stores through the LSU while the TMA loads saturate the L2 slices, the
    // store stage alone took ~5k cycles per tile and made the GELU epilogue (7.4k cycles of math) the bound of fc1
    constexpr bool TMA_ST = (EPI != EPI_F32);
    const bool store_thread = ((warp - 2) & 3) == 0 && lane == 0;
    const int bar_id = 1 + half;  // named barrier of this epilogue group (0 is __syncthreads)
    if (TMA_ST && store_thread) tma_prefetch_desc(&tma_out);
    int it = 0;
#pragma unroll 1
    for (int t = cluster_id; t < total_tiles; t += n_clusters, ++it) {
      int g, m0, n0;
      tile_coords(t, g, m0, n0);
      const int buf = it & 1;
      const uint32_t par = (it >> 1) & 1;
      const uint32_t tmem_acc = tmem_base + buf * BN + half * GC;
      const int nh = n0 + half * GC;
      // The tile's 64 bias values go to shared memory before the accumulator is waited for: read from global memory in
      // stage 1 (this kernel leaves ~3 KB of L1) every FADD of the bias was a long-scoreboard stall, ~15 % of the qkv
      // kernel's stall samples (profiles/r02_b32_gemm_full.md).  Double-buffered by tile: a warp of the group may still
      // be in stage 1 of the previous tile.
      float* const sb = s_bias2 + (half * 2 + buf) * GC;
      const int tg = threadIdx.x - 64 - half * 128;  // thread index inside the epilogue group
      if (tg < GC) sb[tg] = ep.bias ? __ldg(ep.bias + g * ep.bias_gstride + nh + tg) : 0.0f;
      const float* bias = sb;
      if (!TMA_ST) asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
      if (TMA_ST) {
        if (store_thread) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // staging tile free again
        asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
        gemm_epilogue<GC, EPI, true>(stage, bias, &acc_full[buf], par, tmem_acc, m0, nh, g, 0, shape, ep, warp,
                                     lane TRACE_ARGS);
      } else {
        constexpr int CW = GC / 2;  // fp32 rows: two 32-column passes through the same 16 KB
#pragma unroll 1
        for (int c = 0; c < GC; c += CW)
          gemm_epilogue<CW, EPI>(stage, bias + c, &acc_full[buf], par, tmem_acc + c, m0, nh + c, g, 0, shape, ep, warp,
                                 lane TRACE_ARGS);
      }
      // this warp's TMEM reads of the tile are complete (the epilogue ends its stage 1 with a tcgen05 fence, which is what
      // orders them before the leader's next MMAs into this buffer).  RELAXED arrive: with .release the fp32 flavour
      // waited here for its global stores of the tile to be performed cluster-wide (MEMBAR.ALL.CTA + ERRBAR: 12 % of
      // the out-proj kernel's stall samples, profiles/r02_b32_gemm_full.md) although nobody consumes them through this
      // barrier.
      if (lane == 0)
        asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(leader_empty + buf * 8) : "memory");
      if (TMA_ST) {
        fence_proxy_async_smem();  // generic-proxy staging writes -> visible to the TMA (async proxy)
        asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
        if (store_thread) {
          asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                           reinterpret_cast<uint64_t>(&tma_out)),
                       "r"(smem_u32(stage)), "r"(nh), "r"(m0), "r"(g)
                       : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
      if (threadIdx.x == 64 && it < 4) TRACE_PT(0x106);
    }
    if (TMA_ST && store_thread) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // writes complete before exit
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // neither CTA may release its TMEM / shared memory while the pair still uses it
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, 2 * BN);
  }
  if (threadIdx.x == 64) TRACE_PT(0x107);
  if (threadIdx.x == 0 || threadIdx.x == 32 || threadIdx.x == 64) TRACE_FLUSH();
}

}  // namespace uvlt
