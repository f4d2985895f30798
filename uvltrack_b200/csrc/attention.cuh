// Fused multi-head attention over one packed QKV activation:
//     O[b, q, h, :] = softmax_k( Q[b,q,h,:] . K[b,k,h,:] * scale + bias[b,k] ) @ V[b,k,h,:]
// for the joint [cls | template | search | text] sequence of the ViT blocks (reference block.py:47-61: masked keys are
// filled with -1e10) and for the BERT text layers (bert_backbone.py:299-325: additive -10000 mask).  Both mask
// dialects are an additive per-key fp32 bias here (adding -1e10 to an O(10) fp32 score is exactly -1e10).
//
// QKV layout in HBM: bf16 [B, n, 3*H*64] exactly as the qkv GEMM writes it (token-major, [which][head][64]).
// One CTA = one (batch, head, 128-query tile), 64 keys per block, 192 threads:
//   warp 0      TMA producer : Q tile once, then K_j / V_j tiles (64 keys x 64) through a 2-stage ring
//   warp 1      MMA issuer   : S_j = Q K_j^T (tcgen05 M=128, N<=64, K=64) -> TMEM cols [0,64); O += P_j V_j (M=128, N=64,
//                              K<=64) -> TMEM cols [64,128).  V is consumed straight from its [key][64] tile as an MN-major
//                              B operand.  All descriptors are built before the loop.
//   warps 2..5  softmax      : thread = query row.  One wide tcgen05.ld brings the 64 scores of the block into registers
//                              and S is released at once (QK_{j+1} overlaps the softmax of block j); row maximum (3-input
//                              max), exp2 against the running reference (FFMA + ex2.approx + FADD per score; the reference
//                              moves only when the maximum grew by > 2^8, then O / l are rescaled in TMEM), P -> bf16 ->
//                              swizzled smem (A operand of the PV MMA).  Masked / partial blocks use the same registers
//                              with the additive per-key bias.
// The per-block chain of a CTA is latency bound (every barrier wait / arrive / fence costs a lone warp 100-250 cycles of
// issue stall, profiles/r01_attention_experiments.md), so the kernel is built for residency instead: 64.3 KB of shared
// memory, 128 TMEM columns and <= 112 registers per thread -> THREE CTAs per SM whose chains interleave.
#pragma once
#include "common.cuh"

namespace uvlt {

constexpr int ATT_BQ = 128;    // queries per CTA
constexpr int ATT_BKV = 64;    // keys per block
constexpr int ATT_D = 64;      // head dim (both UVLTrack-B and -L)
constexpr int ATT_THREADS = 192;
constexpr int ATT_STAGES = 2;
constexpr int ATT_MAX_KV = 2048;  // 32 key blocks (bit mask of biased blocks)

struct AttnParams {
  int n;               // sequence length (queries == keys)
  int H;               // heads
  float scale_log2;    // head_dim^-0.5 * log2(e)
  const float* bias;   // [B, n] additive key bias (natural-log domain), or nullptr
  __nv_bfloat16* out;  // [B, n, H*64]
};

struct AttnSmem {
  static constexpr int Q_BYTES = ATT_BQ * ATT_D * 2;       // 16 KB
  static constexpr int KV_BYTES = ATT_BKV * ATT_D * 2;     // 8 KB each
  static constexpr int P_BYTES = ATT_BQ * ATT_BKV * 2;     // 16 KB per buffer
  static constexpr int OFF_Q = 0;
  static constexpr int OFF_K = OFF_Q + Q_BYTES;
  static constexpr int OFF_V = OFF_K + ATT_STAGES * KV_BYTES;
  static constexpr int OFF_P = OFF_V + ATT_STAGES * KV_BYTES;
  static constexpr int OFF_BAR = OFF_P + P_BYTES;
  static constexpr int TOTAL = OFF_BAR + 256;  // 65792 B: 3 x (TOTAL + 1 KB reserved) <= 228 KB per SM
  // key-split variant (cluster of two CTAs per query tile): the partner's partial output lands here
  static constexpr int OFF_XO = TOTAL;                          // [16 x (128 rows x float4)] = 32 KB, fp32 O partial
  static constexpr int OFF_XML = OFF_XO + ATT_BQ * ATT_D * 4;   // [128] float2 (reference, row sum)
  static constexpr int TOTAL_SPLIT = OFF_XML + ATT_BQ * 8;
};

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}

constexpr float ATT_LOG2E = 1.4426950408889634f;

// MODE 0: full unbiased block (fast path)   MODE 1: partial block, no bias   MODE 2: biased (and possibly partial)
// maximum of the (biased, log2-scaled) scores of one 32-column chunk
template <int MODE>
__device__ __forceinline__ float chunk_max(const uint32_t (&v)[32], float scale, const float* bias_c, int lim) {
  float m0 = -INFINITY, m1 = -INFINITY;
  if (MODE == 0) {
    // raw maximum; the positive scale is applied once afterwards
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      m0 = fmax3(m0, __uint_as_float(v[i]), __uint_as_float(v[i + 1]));
      m1 = fmax3(m1, __uint_as_float(v[i + 2]), __uint_as_float(v[i + 3]));
    }
    return fmaxf(m0, m1) * scale;
  }
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    // columns >= lim were never written by the MMA (stale TMEM) or belong to keys >= n
    float x = __uint_as_float(v[i]) * scale;
    if (MODE == 2) x = fmaf(__uint_as_float(v[i]), scale, (i < lim ? __ldg(bias_c + i) : 0.0f) * ATT_LOG2E);
    m0 = fmaxf(m0, i < lim ? x : -INFINITY);
  }
  return m0;
}

// probabilities of one 32-column chunk -> bf16 pairs; returns the fp32 row-sum contribution
template <int MODE>
__device__ __forceinline__ float chunk_exp(const uint32_t (&v)[32], uint32_t (&pk)[16], float scale, float neg_m,
                                           const float* bias_c, int lim) {
  float s0 = 0.f, s1 = 0.f;
#pragma unroll
  for (int i = 0; i < 32; i += 2) {
    float p0, p1;
    if (MODE == 2) {
      const float b0 = (i < lim ? __ldg(bias_c + i) : 0.0f) * ATT_LOG2E;
      const float b1 = (i + 1 < lim ? __ldg(bias_c + i + 1) : 0.0f) * ATT_LOG2E;
      // scale and reference in ONE fma exactly as in the unbiased modes, the bias added afterwards: a key whose bias
      // is 0 gets bit-identical probabilities whichever mode its block runs in (so masking the text keys and dropping
      // them give the same image rows, engine option skip_text)
      p0 = ex2_approx(fmaf(__uint_as_float(v[i]), scale, neg_m) + b0);
      p1 = ex2_approx(fmaf(__uint_as_float(v[i + 1]), scale, neg_m) + b1);
    } else {
      p0 = ex2_approx(fmaf(__uint_as_float(v[i]), scale, neg_m));
      p1 = ex2_approx(fmaf(__uint_as_float(v[i + 1]), scale, neg_m));
    }
    if (MODE != 0) {
      p0 = i < lim ? p0 : 0.0f;  // stale TMEM columns past the last real key must not reach P
      p1 = i + 1 < lim ? p1 : 0.0f;
    }
    s0 += p0;
    s1 += p1;
    pk[i >> 1] = pack_bf16x2(p0, p1);
  }
  return s0 + s1;
}

__device__ __forceinline__ float chunk_max_dyn(int mode, const uint32_t (&v)[32], float scale, const float* bias_c,
                                               int lim) {
  if (mode == 0) return chunk_max<0>(v, scale, nullptr, 32);
  if (mode == 1) return chunk_max<1>(v, scale, nullptr, lim);
  return chunk_max<2>(v, scale, bias_c, lim);
}
__device__ __forceinline__ float chunk_exp_dyn(int mode, const uint32_t (&v)[32], uint32_t (&pk)[16], float scale,
                                               float neg_m, const float* bias_c, int lim) {
  if (mode == 0) return chunk_exp<0>(v, pk, scale, neg_m, nullptr, 32);
  if (mode == 1) return chunk_exp<1>(v, pk, scale, neg_m, nullptr, lim);
  return chunk_exp<2>(v, pk, scale, neg_m, bias_c, lim);
}

// SPLIT (launched as clusters of two CTAs along x, small grids only): the two CTAs of a cluster share one query tile and
// each takes half of the key blocks -- at batch 1 the kernel is 60 serial per-CTA chains of 9 key blocks on 148 SMs, so
// halving the chain is worth more than the exchange.  CTA 1 sends its partial (O, reference, row sum) through
// distributed shared memory; CTA 0 merges the two partial softmaxes exactly and stores.
template <bool SPLIT>
static __global__ void __launch_bounds__(ATT_THREADS, 3)
attention_kernel(const __grid_constant__ CUtensorMap tma_q, const __grid_constant__ CUtensorMap tma_kv, const AttnParams p) {
  extern __shared__ __align__(1024) uint8_t att_smem[];  // 128B-swizzled TMA/UMMA tiles need 1024 B alignment
  uint8_t* const smem = att_smem;
  uint8_t* sQ = smem + AttnSmem::OFF_Q;
  uint8_t* sK = smem + AttnSmem::OFF_K;
  uint8_t* sV = smem + AttnSmem::OFF_V;
  uint8_t* sP = smem + AttnSmem::OFF_P;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + AttnSmem::OFF_BAR);
  uint64_t* q_full = bars + 0;
  uint64_t* kv_full = bars + 1;                  // [ATT_STAGES]  K_j and V_j landed
  uint64_t* kv_empty = kv_full + ATT_STAGES;     // [ATT_STAGES]  PV_j drained (K_j was consumed earlier by QK_j)
  uint64_t* s_full = kv_empty + ATT_STAGES;      // S_j landed in TMEM
  uint64_t* s_empty = s_full + 1;                // the softmax warps hold S_j in registers (128 arrivals)
  uint64_t* p_full = s_empty + 1;                // P_j in smem, O rescaled (128 arrivals)
  uint64_t* pv_done = p_full + 1;                // PV_j drained: P buffer reusable, O includes block j
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + 1);

  const int warp = uniform_warp_idx();
  const int lane = threadIdx.x & 31;
  const int q0 = (SPLIT ? (blockIdx.x >> 1) : blockIdx.x) * ATT_BQ;
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int D = p.H * ATT_D;
  const int nblk = (p.n + ATT_BKV - 1) / ATT_BKV;  // key blocks of the sequence
  // this CTA's key blocks [jb, je); loop counters below are local (jj = j - jb), masks / tails use the global j
  const uint32_t crank = SPLIT ? cluster_ctarank() : 0u;
  const int jb = (SPLIT && crank) ? (nblk + 1) / 2 : 0;
  const int je = (SPLIT && !crank) ? (nblk + 1) / 2 : nblk;

  if (warp == 0 && lane == 0) {
    if (smem_u32(smem) & 1023u) __trap();  // dynamic shared memory must be 1024-byte aligned
    tma_prefetch_desc(&tma_q);
    tma_prefetch_desc(&tma_kv);
    mbar_init(q_full, 1);
    for (int s = 0; s < ATT_STAGES; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(s_empty, 128);
    mbar_init(p_full, 128);
    mbar_init(pv_done, 1);
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 128);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  // SPLIT: CTA 1 stores its partial into CTA 0's shared memory at the end.  A cluster's CTAs are co-scheduled, but a
  // distributed-shared-memory access is only defined once the target CTA has started executing: one cluster barrier up
  // front makes that explicit (compute-sanitizer racecheck: "block that might not have entered yet").
  if (SPLIT) cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_S = tmem_base;             // 64 columns
  const uint32_t tmem_O = tmem_base + ATT_BKV;   // 64 columns
  pdl_wait();  // bias and qkv come from earlier kernels of the chain
  pdl_trigger();

  // producer and MMA warps stay converged, one elected lane issues the asynchronous instructions (elect_one_sync in
  // common.cuh: no ELECT / R2UR waterfall loop around every tcgen05.mma / TMA instruction)
  if (warp == 0) {
    {
      // ---------------- TMA producer ----------------
      if (elect_one_sync()) {
        mbar_expect_tx(q_full, AttnSmem::Q_BYTES);
        tma_load_3d(sQ, &tma_q, q_full, h * ATT_D, q0, b);
      }
      __syncwarp();
      int s = 0;
      uint32_t ph = 0;
#pragma unroll 1
      for (int j = jb; j < je; ++j) {
        mbar_wait(&kv_empty[s], ph ^ 1);
        if (elect_one_sync()) {
          mbar_expect_tx(&kv_full[s], 2 * AttnSmem::KV_BYTES);
          tma_load_3d(sK + s * AttnSmem::KV_BYTES, &tma_kv, &kv_full[s], D + h * ATT_D, j * ATT_BKV, b);
          tma_load_3d(sV + s * AttnSmem::KV_BYTES, &tma_kv, &kv_full[s], 2 * D + h * ATT_D, j * ATT_BKV, b);
        }
        __syncwarp();
        if (++s == ATT_STAGES) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    {
      // ---------------- MMA issuer ----------------
      const int last_valid = p.n - (nblk - 1) * ATT_BKV;  // keys in the last block
      const uint32_t idesc_full = umma_idesc_bf16(ATT_BQ, ATT_BKV, 0);
      const uint32_t idesc_last = umma_idesc_bf16(ATT_BQ, (last_valid + 15) & ~15, 0);  // UMMA N granularity 16
      constexpr uint32_t idesc_pv = umma_idesc_bf16(ATT_BQ, ATT_D, 1);
      const uint64_t qd = umma_smem_desc_sw128(smem_u32(sQ), 1024, 0);
      const uint64_t kd_base = umma_smem_desc_sw128(smem_u32(sK), 1024, 0);       // + stage * (KV_BYTES >> 4)
      // V: [key][64] rows of 128 B = MN-major B operand; 16 keys = 16 rows = 2048 B (+128 in the 16-byte address field)
      const uint64_t vd_base = umma_smem_desc_sw128(smem_u32(sV), 1024, 1024);
      // P: [128 x 64] K-major; 16 keys = 32 B inside the swizzle atom (+2 in the address field)
      const uint64_t pd_base = umma_smem_desc_sw128(smem_u32(sP), 1024, 0);       // + buffer * (P_BYTES >> 4)
      constexpr uint64_t KV_STEP = AttnSmem::KV_BYTES >> 4;
      int sq = 0;            // ring stage of the next QK
      uint32_t phq = 0;      // its kv_full parity
      auto issue_qk = [&](int j) {
        const uint64_t kd = kd_base + KV_STEP * sq;
        const uint32_t idesc = (j == nblk - 1) ? idesc_last : idesc_full;
        if (elect_one_sync()) {
          umma_bf16_ss(tmem_S, qd, kd, idesc, 0u);
          umma_bf16_ss(tmem_S, qd + 2, kd + 2, idesc, 1u);
          umma_bf16_ss(tmem_S, qd + 4, kd + 4, idesc, 1u);
          umma_bf16_ss(tmem_S, qd + 6, kd + 6, idesc, 1u);
          umma_commit(s_full);
        }
        __syncwarp();
        if (++sq == ATT_STAGES) { sq = 0; phq ^= 1; }
      };
      mbar_wait(q_full, 0);
      mbar_wait(&kv_full[0], 0);
      tc_fence_after();
      issue_qk(jb);
      int sv = 0;  // ring stage of the next PV
#pragma unroll 1
      for (int j = jb; j < je; ++j) {
        const int jj = j - jb;
        if (j + 1 < je) {
          // the softmax warps hold S_j in registers -> overwrite it with S_{j+1} while they compute P_j
          mbar_wait(&kv_full[sq], phq);
          mbar_wait(s_empty, jj & 1);
          tc_fence_after();
          issue_qk(j + 1);
        }
        mbar_wait(p_full, jj & 1);
        tc_fence_after();
        const int ksteps = (j == nblk - 1) ? ((last_valid + 15) >> 4) : (ATT_BKV / 16);
        const uint64_t vd = vd_base + KV_STEP * sv;
        const uint64_t pd = pd_base;
        if (elect_one_sync()) {
          umma_bf16_ss(tmem_O, pd, vd, idesc_pv, jj > 0 ? 1u : 0u);
          if (ksteps > 1) umma_bf16_ss(tmem_O, pd + 2, vd + 128, idesc_pv, 1u);
          if (ksteps > 2) umma_bf16_ss(tmem_O, pd + 4, vd + 256, idesc_pv, 1u);
          if (ksteps > 3) umma_bf16_ss(tmem_O, pd + 6, vd + 384, idesc_pv, 1u);
          umma_commit(&kv_empty[sv]);
          umma_commit(pv_done);
        }
        __syncwarp();
        if (++sv == ATT_STAGES) sv = 0;
      }
    }
  } else {
    // ---------------- softmax / correction / epilogue warps ----------------
    const int lane_grp = warp & 3;
    const int row = lane_grp * 32 + lane;  // query row inside the tile == TMEM lane
    const uint32_t lane_off = static_cast<uint32_t>(lane_grp * 32) << 16;
    const float scale = p.scale_log2;
    // which key blocks carry a non-zero bias (warp-uniform bit mask); all other full blocks take the fast path
    uint32_t biased = 0;
    if (p.bias) {
      // lane-strided over the whole sequence with NO vote inside the loop: the loads are independent and pipeline.  (One
      // __any_sync per block made this a chain of dependent L2 round trips -- 24 % of the stall samples of the batch-1
      // kernel, profiles/r02_b1_attn_full.md.)
      const float* bb = p.bias + static_cast<long long>(b) * p.n;
#pragma unroll 4
      for (int k = lane; k < p.n; k += 32)
        if (__ldg(bb + k) != 0.0f) biased |= 1u << (k / ATT_BKV);
      biased = __reduce_or_sync(0xffffffffu, biased);
    }
    float m_run = -INFINITY;  // reference of the running sum / output (scaled log2 domain): the largest score seen,
                              // moved only when the maximum grows by more than 2^8 (P stays <= 256; exact after O / l)
    float l_run = 0.0f;
    for (int j = jb; j < je; ++j) {
      const int jj = j - jb;
      const int kv_valid = min(ATT_BKV, p.n - j * ATT_BKV);
      const int nchunk = (((kv_valid + 15) & ~15) + 31) >> 5;  // 32-key chunks the PV MMA may read: must be written
      const int mode = ((biased >> j) & 1u) ? 2 : (kv_valid < ATT_BKV ? 1 : 0);  // block-uniform
      const float* bj = p.bias ? p.bias + static_cast<long long>(b) * p.n + j * ATT_BKV : nullptr;
      uint32_t v[2][32], pk[16];
      mbar_wait(s_full, jj & 1);
      tc_fence_after();
      // one wide TMEM load per block (every load -> wait round trip stalls the lone softmax warp of a scheduler for
      // ~150-200 cycles; three narrow loads per block made the kernel 35 % slower)
      tmem_ld64(tmem_S + lane_off, v[0], v[1]);
      tmem_wait_ld_dep2(v[0], v[1]);
      tc_fence_before();
      mbar_arrive(s_empty);  // the scores are in registers: QK_{j+1} may overwrite S

      // ---- block maximum and the exp reference ----
      float m_blk;
      if (mode == 0) {
        m_blk = fmaxf(chunk_max<0>(v[0], scale, nullptr, 32), chunk_max<0>(v[1], scale, nullptr, 32));
      } else {
        m_blk = chunk_max_dyn(mode, v[0], scale, bj, kv_valid);
        if (nchunk > 1) m_blk = fmaxf(m_blk, chunk_max_dyn(mode, v[1], scale, bj + 32, kv_valid - 32));
      }
      const float ref = (m_blk > m_run + 8.0f) ? m_blk : m_run;  // m_run = -inf on the first block -> m_blk
      const float alpha = (ref == m_run) ? 1.0f : ex2_approx(m_run - ref);  // 0 on the first block
      const float neg_ref = -ref;

      // ---- the P buffer (and O) are free once PV_{j-1} has drained ----
      if (jj > 0) mbar_wait(pv_done, (jj - 1) & 1);
      tc_fence_after();
      if (jj > 0 && __any_sync(0xffffffffu, alpha != 1.0f)) {  // bring the running output to the new reference (rare)
#pragma unroll
        for (int c = 0; c < ATT_D; c += 32) {
          uint32_t o[32];
          tmem_ld32(tmem_O + lane_off + c, o);
          tmem_wait_ld();
#pragma unroll
          for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
          tmem_st32(tmem_O + lane_off + c, o);
        }
        tmem_wait_st();
      }

      // ---- probabilities -> bf16 -> swizzled smem ----
      uint8_t* const p_row = sP + row * 128;
      float l_blk = 0.0f;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        if (c < nchunk) {
          if (mode == 0) l_blk += chunk_exp<0>(v[c], pk, scale, neg_ref, nullptr, 32);
          else l_blk += chunk_exp_dyn(mode, v[c], pk, scale, neg_ref, bj + c * 32, kv_valid - c * 32);
          // 32 keys = four 16 B chunks of this row; the chunk index is XOR-swizzled with row % 8 (128B swizzle)
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int chunk = (c * 4 + q) ^ (row & 7);
            *reinterpret_cast<uint4*>(p_row + chunk * 16) = make_uint4(pk[q * 4], pk[q * 4 + 1], pk[q * 4 + 2], pk[q * 4 + 3]);
          }
        }
      }
      l_run = l_run * alpha + l_blk;
      m_run = ref;
      fence_proxy_async_smem();  // generic-proxy P writes -> visible to the tensor core (async proxy)
      tc_fence_before();
      mbar_arrive(p_full);
    }
    // ---------------- epilogue: O / l -> bf16 ----------------
    mbar_wait(pv_done, (je - jb - 1) & 1);
    tc_fence_after();
    const int q = q0 + row;
    float inv_l = 1.0f / l_run;
    float a_own = 1.0f, a_peer = 0.0f;
    uint8_t* const xo = smem + AttnSmem::OFF_XO;
    if (SPLIT) {
      if (crank == 1) {
        // send the partial to CTA 0: chunk-major float4s, so that a warp writes 512 contiguous bytes per instruction
        uint32_t remote_o, remote_ml;
        asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(remote_o) : "r"(smem_u32(xo)));
        asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(remote_ml) : "r"(smem_u32(smem + AttnSmem::OFF_XML)));
#pragma unroll
        for (int c = 0; c < ATT_D; c += 32) {
          uint32_t o[32];
          tmem_ld32(tmem_O + lane_off + c, o);
          tmem_wait_ld();
#pragma unroll
          for (int i = 0; i < 32; i += 4)
            asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(remote_o + (((c + i) >> 2) * ATT_BQ + row) * 16),
                         "r"(o[i]), "r"(o[i + 1]), "r"(o[i + 2]), "r"(o[i + 3])
                         : "memory");
        }
        asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(remote_ml + row * 8), "f"(m_run), "f"(l_run) : "memory");
        tc_fence_before();
      }
      cluster_sync_all();  // release (CTA 1's stores) / acquire (CTA 0's loads); every thread of both CTAs takes part
      if (crank == 0) {
        const float2 ml = *reinterpret_cast<const float2*>(smem + AttnSmem::OFF_XML + row * 8);
        const float m = fmaxf(m_run, ml.x);
        a_own = ex2_approx(m_run - m);
        a_peer = ex2_approx(ml.x - m);
        inv_l = 1.0f / (l_run * a_own + ml.y * a_peer);
      }
    }
    if (!SPLIT || crank == 0) {
#pragma unroll
      for (int c = 0; c < ATT_D; c += 32) {
        uint32_t o[32];
        tmem_ld32(tmem_O + lane_off + c, o);
        tmem_wait_ld();
        if (SPLIT) {
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const float4 x = *reinterpret_cast<const float4*>(xo + (((c + i) >> 2) * ATT_BQ + row) * 16);
            o[i] = __float_as_uint(__uint_as_float(o[i]) * a_own + x.x * a_peer);
            o[i + 1] = __float_as_uint(__uint_as_float(o[i + 1]) * a_own + x.y * a_peer);
            o[i + 2] = __float_as_uint(__uint_as_float(o[i + 2]) * a_own + x.z * a_peer);
            o[i + 3] = __float_as_uint(__uint_as_float(o[i + 3]) * a_own + x.w * a_peer);
          }
        }
        if (q < p.n) {
          __nv_bfloat16* dst = p.out + (static_cast<long long>(b) * p.n + q) * D + h * ATT_D + c;
#pragma unroll
          for (int i = 0; i < 32; i += 8) {
            uint4 u;
            u.x = pack_bf16x2(__uint_as_float(o[i]) * inv_l, __uint_as_float(o[i + 1]) * inv_l);
            u.y = pack_bf16x2(__uint_as_float(o[i + 2]) * inv_l, __uint_as_float(o[i + 3]) * inv_l);
            u.z = pack_bf16x2(__uint_as_float(o[i + 4]) * inv_l, __uint_as_float(o[i + 5]) * inv_l);
            u.w = pack_bf16x2(__uint_as_float(o[i + 6]) * inv_l, __uint_as_float(o[i + 7]) * inv_l);
            *reinterpret_cast<uint4*>(dst + i) = u;
          }
        }
      }
    }
    tc_fence_before();
  }
  if (SPLIT && warp < 2) {  // the producer / MMA warps' share of the exchange barrier
    __syncwarp();
    cluster_sync_all();
  }

  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 128);
  }
}

}  // namespace uvlt
