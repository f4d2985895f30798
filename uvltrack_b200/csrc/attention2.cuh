// Fused multi-head attention, second generation (round 2): same contract as attention.cuh
//     O[b, q, h, :] = softmax_k( Q[b,q,h,:] . K[b,k,h,:] * scale + bias[b,k] ) @ V[b,k,h,:]
// (reference block.py:47-61 masked_fill(-1e10) and bert_backbone.py:299-325 additive -10000, both as an additive per-key
// fp32 bias), rebuilt around what bounded the first kernel (profiles/r01_attention_experiments.md): the probabilities
// took the long way (registers -> bf16 -> swizzled shared memory -> fence.proxy.async -> SS MMA), one query tile per CTA
// left the tensor pipe idle during every softmax, and each 64-key block paid ~1.3k cycles of fixed synchronisation.
//
// One CTA per SM, 384 threads (three warpgroups; setmaxnreg moves registers from the third to the softmax groups), ALL 512 TMEM columns, two independent "slots" that ping-pong on the tensor pipe:
//   warps 0..3  softmax of slot A      warps 4..7  softmax of slot B      (thread = query row = TMEM lane)
//   warp 8      TMA producer: Q tile(s) once, then K_j / V_j tiles of 128 keys through a 4-stage ring
//   warps 9,10  MMA issuers, one thread per slot: S_t = Q_t K_j^T (M=128, N<=128, K=64, SS) -> TMEM;  O_t += P_t V_j with the A
//               operand P_t read from TENSOR MEMORY (tcgen05.mma with a TMEM A operand, 8 x K=16) and V consumed in
//               place from its [key][64] tile as an MN-major B operand
// TMEM per slot (256 columns): S fp32 [128 x 128] | P bf16 [128 x 128] packed two per column (64 columns) | O fp32
// [128 x 64].  S, P and O do not alias, so QK_{j+1} is issued as soon as the softmax warps hold S_j in registers and
// the softmax of block j+1 starts right after block j; only the P store waits for PV_j (long done by then).
//   PAIR  mode: the two slots are two neighbouring query tiles of one (batch, head); every K/V tile is fetched once
//               and consumed by both (kv_empty counts two PV commits).
//   SPLIT mode: both slots work on the SAME query tile and take half of the key blocks each (own ring stages); slot B
//               hands its partial (O, reference, row sum) to slot A through shared memory and A merges the two partial
//               softmaxes exactly.  Used for the odd last tile of a sequence and for every tile when the grid is small
//               (batch 1-2: the per-CTA chain of key blocks is the latency of the kernel).
// Softmax per block (thread = row, 128 scores in registers): 3-input max, lazy reference (moves only when the maximum
// grew by > 2^8; O / l are then rescaled in TMEM), P = 2^(s*scale - ref): FFMA + ex2.approx + FADD per score; a fixed
// quarter of the columns (POLY) evaluates 2^x on the FMA pipe instead of the MUFU (Cody-Waite split + degree-3
// polynomial, max rel. error 7.7e-5, 25x below the bf16 rounding of P; coefficients and error study in
// profiles/r01_attention_experiments.md) -- at head dim 64 a 128 x 128 tile needs 1024 MUFU cycles against 512 tensor
// cycles, so the MUFU is the first wall.  P goes bf16 -> tcgen05.st -> TMEM (no shared-memory round trip, no
// fence.proxy.async).
#pragma once
#include "attention.cuh"

namespace uvlt {

constexpr int AT2_BQ = 128;
constexpr int AT2_BK = 128;
constexpr int AT2_THREADS = 384;  // three warpgroups: softmax A, softmax B, {TMA, MMA A, MMA B, idle}
constexpr int AT2_STAGES = 4;  // SPLIT mode keeps four K/V tiles in flight: (step i, i+1) x (slot A, B)

struct Attn2Smem {
  static constexpr int Q_BYTES = AT2_BQ * ATT_D * 2;      // 16 KB per query tile
  static constexpr int KV_BYTES = AT2_BK * ATT_D * 2;     // 16 KB each for K and V
  static constexpr int STAGE_BYTES = 2 * KV_BYTES;
  static constexpr int OFF_Q = 0;
  static constexpr int OFF_KV = OFF_Q + 2 * Q_BYTES;
  static constexpr int OFF_XO = OFF_KV + AT2_STAGES * STAGE_BYTES;  // SPLIT: slot B's fp32 partial O, chunk-major float4
  static constexpr int OFF_XML = OFF_XO + AT2_BQ * ATT_D * 4;        // [128] float2 (reference, row sum)
  static constexpr int OFF_BIAS = OFF_XML + AT2_BQ * 8;              // [ATT_MAX_KV] key bias * log2(e)
  static constexpr int OFF_BAR = OFF_BIAS + ATT_MAX_KV * 4;
  static constexpr int TOTAL = OFF_BAR + 256;
};

struct Attn2Params {
  int n;
  int H;
  float scale_log2;
  const float* bias;
  __nv_bfloat16* out;
  int split_all;  // 1: every CTA runs ONE query tile in SPLIT mode (small grids); 0: PAIR mode, odd last tile SPLIT
  int zero;       // always 0: an opaque branch condition that separates scheduling regions (see at2_chunk_ex2)
};

// D[tmem] (+)= A[tmem] * B[smem]: the A operand (M = 128 rows = lanes, K = 16 bf16 = 8 packed 32-bit columns) is read
// from tensor memory
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}

__device__ __forceinline__ float lds_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}

// 2^x on the FMA pipe: x = n + f, n = round(x), f in [-0.5, 0.5]; 2^f by a degree-3 polynomial (weighted least squares
// on Chebyshev nodes, max rel. error 7.7e-5), n added to the exponent field.  x is clamped at -125 (masked keys sit at
// -1.4e10) so that the exponent add cannot wrap; results below 2^-125 are irrelevant to a bf16 P <= 2^8.
__device__ __forceinline__ float ex2_poly3(float x) {
  x = fmaxf(x, -125.0f);
  const float t = x + 12582912.0f;  // 1.5 * 2^23: the integer part lands in the low mantissa bits (round to nearest)
  const float n = t - 12582912.0f;
  const float f = x - n;
  float p = fmaf(f, 0.05508868f, 0.24260405f);
  p = fmaf(p, f, 0.69327623f);
  p = fmaf(p, f, 0.99992895f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}

// maximum of one 32-column chunk of scaled (and biased) scores.  MODE 0: full unbiased block (raw maximum, scaled once by
// the caller); MODE 2: partial and / or biased block (sb = shared address of this chunk's bias * log2e, zero where no
// bias applies; columns >= lim are stale TMEM or keys >= n).
template <int MODE>
__device__ __forceinline__ float at2_chunk_max(const uint32_t (&v)[32], float scale, uint32_t sb, int lim) {
  if (MODE == 0) {
    float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      m0 = fmax3(m0, __uint_as_float(v[i]), __uint_as_float(v[i + 1]));
      m1 = fmax3(m1, __uint_as_float(v[i + 2]), __uint_as_float(v[i + 3]));
    }
    return fmaxf(m0, m1);
  }
  float m = -INFINITY;
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    const float x = fmaf(__uint_as_float(v[i]), scale, lds_f32(sb + i * 4));
    m = fmaxf(m, i < lim ? x : -INFINITY);
  }
  return m;
}

// The softmax of a 32-column chunk in two separately scheduled halves.  Written as "p = ex2(x); sum += p; pack(p)" ptxas
// places every consumer two instructions behind its MUFU whatever the source order (it also reorders volatile asm), and
// the in-order warp then stalls for the MUFU latency on every column: 16.5 cycles per column measured with the warp alone
// on its scheduler (in-kernel timeline, profiles/r02_attention.md) against the 8 the MUFU pipe needs.  So the
// exponentials are written back IN PLACE by at2_chunk_ex2 and consumed by at2_chunk_sum_pack one basic block later (the
// caller separates them with a branch ptxas cannot see through): MUFUs issue back to back, and a chunk's sums / packs
// fill the issue slots under the next chunk's MUFUs.
// MODE 0: full unbiased block.  MODE 2: partial and / or biased block (sb = shared address of this chunk's bias * log2e;
// scale and reference in ONE fma exactly as in mode 0 and the bias added afterwards, so a key whose bias is 0 gets
// bit-identical probabilities whichever mode its block runs in -- engine option skip_text).
// POLY: every fourth column takes the FMA-pipe exponential.
template <int MODE, bool POLY>
__device__ __forceinline__ void at2_chunk_ex2(uint32_t (&v)[32], float scale, float neg_ref, uint32_t sb, int lim) {
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    float x = fmaf(__uint_as_float(v[i]), scale, neg_ref);
    if (MODE == 2) x += lds_f32(sb + i * 4);
    float y = (POLY && (i & 3) == 3) ? ex2_poly3(x) : ex2_approx(x);
    if (MODE != 0) y = i < lim ? y : 0.0f;  // stale TMEM columns past the last real key must not reach P
    v[i] = __float_as_uint(y);
  }
}
__device__ __forceinline__ float at2_chunk_sum_pack(const uint32_t (&v)[32], uint32_t (&pk)[16]) {
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
  for (int i = 0; i < 32; i += 4) {
    s0 += __uint_as_float(v[i]);
    s1 += __uint_as_float(v[i + 1]);
    s2 += __uint_as_float(v[i + 2]);
    s3 += __uint_as_float(v[i + 3]);
    pk[i >> 1] = pack_bf16x2(__uint_as_float(v[i]), __uint_as_float(v[i + 1]));
    pk[(i >> 1) + 1] = pack_bf16x2(__uint_as_float(v[i + 2]), __uint_as_float(v[i + 3]));
  }
  return (s0 + s1) + (s2 + s3);
}

template <bool POLY>
static __global__ void __launch_bounds__(AT2_THREADS, 1)
attention2_kernel(const __grid_constant__ CUtensorMap tma_qkv, const __grid_constant__ CUtensorMap tma_out,
                  const Attn2Params p) {
  extern __shared__ __align__(1024) uint8_t att2_smem[];
  uint8_t* const smem = att2_smem;
  uint8_t* const sQ = smem + Attn2Smem::OFF_Q;
  uint8_t* const sKV = smem + Attn2Smem::OFF_KV;
  uint64_t* const bars = reinterpret_cast<uint64_t*>(smem + Attn2Smem::OFF_BAR);
  uint64_t* const q_full = bars;                      // [2]
  uint64_t* const kv_full = q_full + 2;               // [AT2_STAGES]
  uint64_t* const kv_empty = kv_full + AT2_STAGES;    // [AT2_STAGES] the PV MMAs that read the stage have drained
  uint64_t* const s_full = kv_empty + AT2_STAGES;     // [2] S_t landed in TMEM
  uint64_t* const s_free = s_full + 2;                // [2] the slot's 4 softmax warps hold S_t in registers
  uint64_t* const p_full = s_free + 2;                // [2] P_t stored (and O_t rescaled)
  uint64_t* const pv_done = p_full + 2;               // [2] PV_t drained: P_t reusable, O_t includes the block
  uint64_t* const stagger = pv_done + 2;              // slot A is half way through its first block (4 warp arrivals)
  uint32_t* const tmem_slot = reinterpret_cast<uint32_t*>(stagger + 1);

  const int warp = uniform_warp_idx();
  const int lane = threadIdx.x & 31;
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int D = p.H * ATT_D;
  const int nblk = (p.n + AT2_BK - 1) / AT2_BK;
  TRACE_DECL;
  if (lane == 0) TRACE_PT(0x200 + warp);
  const int ntiles = (p.n + AT2_BQ - 1) / AT2_BQ;
  // ---- slot configuration (uniform over the CTA) ----
  const int tile0 = p.split_all ? blockIdx.x : 2 * blockIdx.x;
  const bool lone = p.split_all || (tile0 + 1 >= ntiles);  // SPLIT mode: both slots on tile0, half of the key blocks each
  const int q0A = tile0 * AT2_BQ, q0B = lone ? q0A : q0A + AT2_BQ;
  const int jbB = lone ? (nblk + 1) / 2 : 0;                      // slot A always starts at key block 0
  const int nsA = lone ? (nblk + 1) / 2 : nblk, nsB = nblk - jbB;  // key blocks per slot; nsA >= nsB
  const int steps = nsA;
  auto q0_of = [&](int t) { return t ? q0B : q0A; };
  auto jb_of = [&](int t) { return t ? jbB : 0; };
  auto ns_of = [&](int t) { return t ? nsB : nsA; };
  // position of (step i, slot t) in the order K/V tiles travel through the ring
  auto ring_pos = [&](int i, int t) { return lone ? (i < nsB ? 2 * i + t : nsB + i) : i; };

  if (warp == 8 && lane == 0) {
    if (smem_u32(smem) & 1023u) __trap();
    tma_prefetch_desc(&tma_qkv);
    tma_prefetch_desc(&tma_out);
    mbar_init(&q_full[0], 1);
    mbar_init(&q_full[1], 1);
    for (int s = 0; s < AT2_STAGES; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], lone ? 1 : 2);
    }
    for (int t = 0; t < 2; ++t) {
      mbar_init(&s_full[t], 1);
      mbar_init(&s_free[t], 4);
      mbar_init(&p_full[t], 4);
      mbar_init(&pv_done[t], 1);
    }
    mbar_init(stagger, 4);
    fence_mbar_init();
  }
  if (warp == 9) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();  // qkv and the bias come from earlier kernels of the chain
  pdl_trigger();

  // Register budget: the kernel starts with 168 registers per thread (65536 / 384); the producer / MMA warpgroup gives
  // most of its share back and the two softmax warpgroups (128 scores per thread live in registers) take it.
  if (warp >= 8) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
  if (warp == 8) {
    {
      // ---------------- TMA producer (converged warp, one elected lane issues) ----------------
      if (elect_one_sync()) {
        mbar_expect_tx(&q_full[0], Attn2Smem::Q_BYTES);
        tma_load_3d(sQ, &tma_qkv, &q_full[0], h * ATT_D, q0A, b);
        if (!lone) {
          mbar_expect_tx(&q_full[1], Attn2Smem::Q_BYTES);
          tma_load_3d(sQ + Attn2Smem::Q_BYTES, &tma_qkv, &q_full[1], h * ATT_D, q0B, b);
        }
      }
      __syncwarp();
      int c = 0;
#pragma unroll 1
      for (int i = 0; i < steps; ++i) {
#pragma unroll 1
        for (int t = 0; t < (lone ? 2 : 1); ++t) {
          if (i >= ns_of(t)) continue;
          const int j = jb_of(t) + i;
          const int s = c % AT2_STAGES;
          const uint32_t ph = (c / AT2_STAGES) & 1;
          mbar_wait_trap(&kv_empty[s], ph ^ 1);
          if (elect_one_sync()) {
            uint8_t* const dst = sKV + s * Attn2Smem::STAGE_BYTES;
            mbar_expect_tx(&kv_full[s], Attn2Smem::STAGE_BYTES);
            tma_load_3d(dst, &tma_qkv, &kv_full[s], D + h * ATT_D, j * AT2_BK, b);
            tma_load_3d(dst + Attn2Smem::KV_BYTES, &tma_qkv, &kv_full[s], 2 * D + h * ATT_D, j * AT2_BK, b);
          }
          __syncwarp();
          ++c;
        }
      }
    }
  } else if (warp == 9 || warp == 10) {
    {
      // ---------------- MMA issuer of slot t (one converged warp per slot, one elected lane issues) ----------------
      // One issuer per slot: with a single thread walking [QK_A, QK_B, PV_A, PV_B] in a fixed order, each slot's PV waits
      // for the OTHER slot's softmax and the two slots fall into lock step -- both in their exponentials (tensor pipe
      // idle), then both waiting for their PV (MUFU idle): 36 % of the softmax warps' samples sat in s_full / pv_done
      // waits (ncu source view, profiles/r02_attention.md).  The tensor pipe interleaves the two threads' instructions.
      const int t = warp - 9;
      const int ns = ns_of(t), jb = jb_of(t);
      const int last_valid = p.n - (nblk - 1) * AT2_BK;  // keys in the sequence's last block
      const uint32_t idesc_full = umma_idesc_bf16(AT2_BQ, AT2_BK, 0);
      const uint32_t idesc_last = umma_idesc_bf16(AT2_BQ, (last_valid + 15) & ~15, 0);  // UMMA N granularity 16
      constexpr uint32_t idesc_pv = umma_idesc_bf16(AT2_BQ, ATT_D, 1);
      const uint64_t qd = umma_smem_desc_sw128(smem_u32(sQ + ((lone || t == 0) ? 0 : Attn2Smem::Q_BYTES)), 1024, 0);
      const uint64_t kd_base = umma_smem_desc_sw128(smem_u32(sKV), 1024, 0);
      // V: [key][64] rows of 128 B = MN-major B operand; 16 keys = 16 rows = 2048 B (+128 in the 16-byte address field)
      const uint64_t vd_base = umma_smem_desc_sw128(smem_u32(sKV + Attn2Smem::KV_BYTES), 1024, 1024);
      constexpr uint64_t STAGE_STEP = Attn2Smem::STAGE_BYTES >> 4;
      const uint32_t tS = tmem_base + t * 256;
      const uint32_t tP = tS + 128;
      const uint32_t tO = tS + 192;
      auto issue_qk = [&](int i) {
        const int c = ring_pos(i, t);
        mbar_wait_trap(&kv_full[c % AT2_STAGES], (c / AT2_STAGES) & 1);
        if (i > 0) mbar_wait_trap(&s_free[t], (i - 1) & 1);  // the softmax warps hold S_t(i-1) in registers
        tc_fence_after();
        const uint64_t kd = kd_base + STAGE_STEP * (c % AT2_STAGES);
        const uint32_t idesc = (jb + i == nblk - 1) ? idesc_last : idesc_full;
        if (elect_one_sync()) {
          umma_bf16_ss(tS, qd, kd, idesc, 0u);
          umma_bf16_ss(tS, qd + 2, kd + 2, idesc, 1u);
          umma_bf16_ss(tS, qd + 4, kd + 4, idesc, 1u);
          umma_bf16_ss(tS, qd + 6, kd + 6, idesc, 1u);
          umma_commit(&s_full[t]);
        }
        __syncwarp();
        if (lane == 0) TRACE_PT(0x300 + t * 0x100 + 0x10 + i);  // QK(i) issued
      };
      if (ns > 0) {
        mbar_wait_trap(&q_full[(lone || t == 0) ? 0 : 1], 0);
        // PAIR mode: start slot B half a block behind slot A, so that one slot's exponentials (MUFU) run under the other
        // slot's TMEM loads, maximum, barrier round trips and MMAs instead of both doing the same thing at the same time
        if (!lone && t == 1) mbar_wait_trap(stagger, 0);
        issue_qk(0);
      }
#pragma unroll 1
      for (int i = 0; i < ns; ++i) {
        if (i + 1 < ns) issue_qk(i + 1);
        mbar_wait_trap(&p_full[t], i & 1);
        tc_fence_after();
        if (lane == 0) TRACE_PT(0x300 + t * 0x100 + 0x20 + i);  // p_full(i) seen
        const int c = ring_pos(i, t);
        const uint64_t vd = vd_base + STAGE_STEP * (c % AT2_STAGES);
        const int ksteps = (jb + i == nblk - 1) ? ((last_valid + 15) >> 4) : (AT2_BK / 16);
        if (elect_one_sync()) {
          if (ksteps == AT2_BK / 16) {
#pragma unroll
            for (int k = 0; k < AT2_BK / 16; ++k) umma_bf16_ts(tO, tP + 8 * k, vd + 128 * k, idesc_pv, (i > 0 || k > 0) ? 1u : 0u);
          } else {
#pragma unroll 1
            for (int k = 0; k < ksteps; ++k) umma_bf16_ts(tO, tP + 8 * k, vd + 128 * k, idesc_pv, (i > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&kv_empty[c % AT2_STAGES]);
          umma_commit(&pv_done[t]);
        }
        __syncwarp();
        if (lane == 0) TRACE_PT(0x300 + t * 0x100 + 0x30 + i);  // PV(i) issued
      }
      if (lane == 0) TRACE_FLUSH();
    }
  }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
    // ---------------- softmax / correction / epilogue warps ----------------
    const int t = warp >> 2;           // slot
    const int quad = warp & 3;         // TMEM lane quadrant
    const int row = quad * 32 + lane;  // query row inside the tile == TMEM lane
    const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t tS = tmem_base + t * 256 + lane_off;
    const uint32_t tP = tS + 128;
    const uint32_t tO = tS + 192;
    const float scale = p.scale_log2;
    const int ns = ns_of(t), jb = jb_of(t);
    const uint32_t sbias = smem_u32(smem + Attn2Smem::OFF_BIAS);
    // key bias (times log2 e) of the whole sequence -> shared memory; which 128-key blocks carry a non-zero bias
    uint32_t biased = 0;
    {
      float* sb = reinterpret_cast<float*>(smem + Attn2Smem::OFF_BIAS);
      const float* bb = p.bias ? p.bias + static_cast<long long>(b) * p.n : nullptr;
      for (int k = threadIdx.x; k < nblk * AT2_BK; k += 256) sb[k] = (bb && k < p.n) ? __ldg(bb + k) * ATT_LOG2E : 0.0f;
      if (bb) {
        for (int j = 0; j < nblk; ++j) {
          bool nz = false;
          for (int i = lane; i < AT2_BK; i += 32) {
            const int k = j * AT2_BK + i;
            nz |= (k < p.n) && (__ldg(bb + k) != 0.0f);
          }
          if (__any_sync(0xffffffffu, nz)) biased |= 1u << j;
        }
      }
      asm volatile("bar.sync 2, 256;" ::: "memory");  // the 8 softmax warps: bias table complete
    }
    float m_run = -INFINITY;  // reference of the running sum / output (scaled log2 domain)
    float l_run = 0.0f;
    for (int i = 0; i < ns; ++i) {
      const int j = jb + i;
      const int kv_valid = min(AT2_BK, p.n - j * AT2_BK);
      const int nchunk = (((kv_valid + 15) & ~15) + 31) >> 5;  // 32-key chunks the PV MMA may read: must be written
      const int mode = (((biased >> j) & 1u) || kv_valid < AT2_BK) ? 2 : 0;  // block-uniform
      const uint32_t sbj = sbias + j * AT2_BK * 4;
      uint32_t v[4][32], pk[16];
      mbar_wait_trap(&s_full[t], i & 1);
      tc_fence_after();
      if (lane == 0 && quad == 0) TRACE_PT(0x500 + t * 0x100 + 0x10 + i);  // s_full(i) seen
      tmem_ld64(tS, v[0], v[1]);
      tmem_ld64(tS + 64, v[2], v[3]);
      tmem_wait_ld_dep2(v[0], v[1]);
      tmem_wait_ld_dep2(v[2], v[3]);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_free[t]);  // QK of the slot's next block may overwrite S
      if (lane == 0 && quad == 0) TRACE_PT(0x500 + t * 0x100 + 0x20 + i);  // S in registers

      // ---- block maximum and the exp reference ----
      float m_blk;
      if (mode == 0) {
        m_blk = fmaxf(fmaxf(at2_chunk_max<0>(v[0], scale, 0, 32), at2_chunk_max<0>(v[1], scale, 0, 32)),
                      fmaxf(at2_chunk_max<0>(v[2], scale, 0, 32), at2_chunk_max<0>(v[3], scale, 0, 32))) * scale;
      } else {
        m_blk = -INFINITY;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          if (c < nchunk) {
            m_blk = fmaxf(m_blk, at2_chunk_max<2>(v[c], scale, sbj + c * 128, kv_valid - c * 32));
          }
        }
      }
      const float ref = (m_blk > m_run + 8.0f) ? m_blk : m_run;  // m_run = -inf on the first block -> m_blk
      const float alpha = (ref == m_run) ? 1.0f : ex2_approx(m_run - ref);  // 0 on the first block
      const float neg_ref = -ref;
      // ---- exponentials (in place) one chunk ahead of the sums / packs / P stores; `p.zero` (always 0) gives ptxas a
      //      branch between the stages, i.e. separate scheduling regions (see at2_chunk_ex2) ----
      float l_blk = 0.0f;
      auto ex2_chunk = [&](int c) {
        if (c >= nchunk) return;
        if (mode == 0) at2_chunk_ex2<0, POLY>(v[c], scale, neg_ref, 0, 32);
        else at2_chunk_ex2<2, false>(v[c], scale, neg_ref, sbj + c * 128, kv_valid - c * 32);
      };
      auto store_chunk = [&](int c) {
        if (c >= nchunk) return;
        l_blk += at2_chunk_sum_pack(v[c], pk);
        tmem_st16(tP + c * 16, pk);
      };
      ex2_chunk(0);
      if (p.zero) break;
      ex2_chunk(1);
      if (i == 0 && t == 0 && lane == 0) mbar_arrive(stagger);  // PAIR mode: slot B starts half a block behind slot A
      if (i > 0) {
        // P_t and O_t are free once PV_t(i-1) has drained (two chunks of exponentials were computed meanwhile)
        if (lane == 0 && quad == 0) TRACE_PT(0x500 + t * 0x100 + 0x30 + i);  // max + two chunks done
        mbar_wait_trap(&pv_done[t], (i - 1) & 1);
        tc_fence_after();
        if (lane == 0 && quad == 0) TRACE_PT(0x500 + t * 0x100 + 0x40 + i);  // pv_done(i-1) seen
        if (__any_sync(0xffffffffu, alpha != 1.0f)) {  // bring the running output to the new reference (rare)
#pragma unroll
          for (int cc = 0; cc < ATT_D; cc += 32) {
            uint32_t o[32];
            tmem_ld32(tO + cc, o);
            tmem_wait_ld();
#pragma unroll
            for (int q = 0; q < 32; ++q) o[q] = __float_as_uint(__uint_as_float(o[q]) * alpha);
            tmem_st32(tO + cc, o);
          }
        }
      }
      store_chunk(0);
      if (p.zero) break;
      ex2_chunk(2);
      store_chunk(1);
      if (p.zero) break;
      ex2_chunk(3);
      store_chunk(2);
      if (p.zero) break;
      store_chunk(3);
      l_run = l_run * alpha + l_blk;
      m_run = ref;
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[t]);
      if (lane == 0 && quad == 0) TRACE_PT(0x500 + t * 0x100 + 0x50 + i);  // P stored
    }
    // ---------------- epilogue: O / l -> bf16 ----------------
    if (ns > 0) {
      mbar_wait_trap(&pv_done[t], (ns - 1) & 1);
      tc_fence_after();
    }
    float inv_l = 1.0f / l_run;
    float a_own = 1.0f, a_peer = 0.0f;
    uint8_t* const xo = smem + Attn2Smem::OFF_XO;
    bool merge = false;
    if (lone) {
      if (t == 1 && ns > 0) {
        // slot B -> slot A: chunk-major float4s (a warp writes 512 contiguous bytes per instruction)
#pragma unroll
        for (int c = 0; c < ATT_D; c += 32) {
          uint32_t o[32];
          tmem_ld32(tO + c, o);
          tmem_wait_ld();
#pragma unroll
          for (int q = 0; q < 32; q += 4)
            *reinterpret_cast<uint4*>(xo + (((c + q) >> 2) * AT2_BQ + row) * 16) = make_uint4(o[q], o[q + 1], o[q + 2], o[q + 3]);
        }
        *reinterpret_cast<float2*>(smem + Attn2Smem::OFF_XML + row * 8) = make_float2(m_run, l_run);
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");  // both slots' softmax warps
      merge = nsB > 0;
      if (t == 0 && merge) {
        const float2 ml = *reinterpret_cast<const float2*>(smem + Attn2Smem::OFF_XML + row * 8);
        const float m = fmaxf(m_run, ml.x);
        a_own = ex2_approx(m_run - m);
        a_peer = ex2_approx(ml.x - m);
        inv_l = 1.0f / (l_run * a_own + ml.y * a_peer);
      }
    }
    if (!lone || t == 0) {
      uint8_t* const stage = sQ + ((lone || t == 0) ? 0 : Attn2Smem::Q_BYTES);
#pragma unroll
      for (int c = 0; c < ATT_D; c += 32) {
        uint32_t o[32];
        tmem_ld32(tO + c, o);
        tmem_wait_ld();
        if (merge) {
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const float4 x = *reinterpret_cast<const float4*>(xo + (((c + i) >> 2) * AT2_BQ + row) * 16);
            o[i] = __float_as_uint(__uint_as_float(o[i]) * a_own + x.x * a_peer);
            o[i + 1] = __float_as_uint(__uint_as_float(o[i + 1]) * a_own + x.y * a_peer);
            o[i + 2] = __float_as_uint(__uint_as_float(o[i + 2]) * a_own + x.z * a_peer);
            o[i + 3] = __float_as_uint(__uint_as_float(o[i + 3]) * a_own + x.w * a_peer);
          }
        }
        // bf16 row -> staging tile in the slot's (no longer needed) Q buffer, 16-byte chunks XOR-swizzled with row % 8:
        // that IS the 128B-swizzle TMA layout, and the writes are bank-conflict free
        uint8_t* const st_row = stage + row * 128;
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          uint4 u;
          u.x = pack_bf16x2(__uint_as_float(o[i]) * inv_l, __uint_as_float(o[i + 1]) * inv_l);
          u.y = pack_bf16x2(__uint_as_float(o[i + 2]) * inv_l, __uint_as_float(o[i + 3]) * inv_l);
          u.z = pack_bf16x2(__uint_as_float(o[i + 4]) * inv_l, __uint_as_float(o[i + 5]) * inv_l);
          u.w = pack_bf16x2(__uint_as_float(o[i + 6]) * inv_l, __uint_as_float(o[i + 7]) * inv_l);
          *reinterpret_cast<uint4*>(st_row + ((((c + i) >> 3) ^ (row & 7)) << 4)) = u;
        }
      }
      // one TMA bulk store per tile (rows >= n are clipped by the tensor map): the row-per-thread global stores of the
      // first version cost ~4k cycles per CTA (32 L1 wavefronts per instruction)
      fence_proxy_async_smem();
      asm volatile("bar.sync %0, 128;" ::"r"(3 + t) : "memory");
      if (quad == 0 && lane == 0) {
        asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                         reinterpret_cast<uint64_t>(&tma_out)),
                     "r"(smem_u32(stage)), "r"(h * ATT_D), "r"(q0_of(t)), "r"(b)
                     : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // the writes are complete before the CTA exits
      }
    }
    tc_fence_before();
    if (lane == 0 && quad == 0) { TRACE_PT(0x500 + t * 0x100 + 0x60); TRACE_FLUSH(); }
  }

  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace uvlt
