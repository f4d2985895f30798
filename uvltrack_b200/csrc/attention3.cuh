// Fused multi-head attention, third generation (round 2): same contract as attention.cuh / attention2.cuh
//     O[b, q, h, :] = softmax_k( Q[b,q,h,:] . K[b,k,h,:] * scale + bias[b,k] ) @ V[b,k,h,:]
// (reference block.py:47-61 masked_fill(-1e10) and bert_backbone.py:299-325 additive -10000, both as an additive per-key
// fp32 bias).  Built on what the in-kernel timelines of the second kernel showed (profiles/r02_attention.md):
//   * one CTA per SM with two slots has the better steady state, but paid ~8k of ~29k cycles per CTA for set-up and
//     drain with nothing else resident  ->  PERSISTENT: one CTA per SM walks a static list of work items; barriers,
//     tensor memory and descriptors are set up once, the producer runs ahead into the next item's Q / K / V, the MMA
//     warp issues QK(0) of the next item before the softmax warps have finished the current item's epilogue;
//   * a softmax warp alone on its scheduler cannot overlap its own MUFU bursts with its FADD / F2FP / tcgen05.st work
//     (in-order issue), and two warps per scheduler only overlap when their phases differ  ->  TWO THREADS PER QUERY ROW:
//     each slot has eight softmax warps (two per TMEM lane quadrant, 64 of the block's 128 key columns each), i.e. four
//     softmax warps per scheduler; the two halves of a row agree on the block maximum through shared memory (one
//     64-thread named barrier per block) and keep partial row sums that are added once per item.
// 640 threads: warps 0..15 softmax (slot = warp / 8, half = (warp / 4) % 2, quadrant = warp % 4), warp 16 TMA producer,
// warps 17 / 18 MMA issuers of slot 0 / 1.  TMEM per slot (256 columns): S fp32 [128 x 128] | P bf16 [128 x 128] packed two
// per column (64 columns) | O fp32 [128 x 64]; P goes registers -> tcgen05.st -> TMEM and the PV MMA reads it as its A
// operand from tensor memory; V is consumed in place from its [key][64] tile as an MN-major B operand.
// Work items (static round-robin over the grid's CTAs, pairs first so that the shorter lone items balance the tail):
//   PAIR : two neighbouring query tiles of one (batch, head), one per slot; every K/V tile is fetched once for both.
//   LONE : the odd last tile of a sequence: both slots work on the SAME query tile (loaded into both Q buffers) and take
//          half of the key blocks each; slot 1 hands its partial (O, reference, row sums) to slot 0 through shared
//          memory and slot 0 merges the two partial softmaxes exactly.
// The key bias is read per block with one coalesced load per 32-key chunk (lane = key) and broadcast by shuffles in
// the chunks that carry a non-zero bias or a ragged tail; all other chunks take the unbiased fast path.
#pragma once
#include "attention2.cuh"

namespace uvlt {

constexpr int AT3_BQ = 128;
constexpr int AT3_BK = 128;
constexpr int AT3_SM_WARPS = 16;                        // softmax warps (both slots)
constexpr int AT3_THREADS = (AT3_SM_WARPS + 4) * 32;    // + producer + two MMA issuers + one idle warp (whole warpgroup for setmaxnreg) = 640
constexpr int AT3_STAGES = 4;
#ifndef AT3_SETMAXNREG
#define AT3_SETMAXNREG 0
#endif

struct Attn3Smem {
  static constexpr int Q_BYTES = AT3_BQ * ATT_D * 2;    // 16 KB per slot
  static constexpr int KV_BYTES = AT3_BK * ATT_D * 2;   // 16 KB each for K and V
  static constexpr int STAGE_BYTES = 2 * KV_BYTES;
  static constexpr int OFF_Q = 0;
  static constexpr int OFF_KV = OFF_Q + 2 * Q_BYTES;
  static constexpr int OFF_ST0 = OFF_KV + AT3_STAGES * STAGE_BYTES;  // bf16 output staging tile of slot 0 (TMA store source)
  static constexpr int OFF_ST1 = OFF_ST0 + Q_BYTES;                  // slot 1's staging tile; LONE items: slot 1's fp32 partial O
  static constexpr int OFF_MX = OFF_ST1 + AT3_BQ * ATT_D * 4;        // [2 parities][2 slots][2 halves][128] block maxima
  static constexpr int OFF_LS = OFF_MX + 2 * 2 * 2 * AT3_BQ * 4;     // [2 slots][2 halves][128] partial row sums
  static constexpr int OFF_MS = OFF_LS + 2 * 2 * AT3_BQ * 4;         // [128] slot 1's reference (LONE merge)
  static constexpr int OFF_BAR = OFF_MS + AT3_BQ * 4;
  static constexpr int TOTAL = OFF_BAR + 256;
};
static_assert(Attn3Smem::TOTAL <= 227 * 1024, "attention3: shared memory budget");

struct Attn3Params {
  int n;
  int H;
  int B;
  float scale_log2;
  const float* bias;
  int zero;  // always 0: an opaque branch condition that separates scheduling regions (see at2_chunk_ex2)
  int pingpong;  // 1: the two slots take turns on the MUFU (named-barrier token around the exponentials of a block)
};

struct At3Item {
  int b, h;
  int q0A, q0B;  // first query row of each slot's tile
  int jbB;       // first key block of slot 1 (slot 0 starts at block 0)
  int nsA, nsB;  // key blocks of each slot (nsA >= nsB; nsB may be 0 for a LONE item of a one-block sequence)
  bool lone;
  // selects, not arrays: an array indexed by the (runtime) slot would live in local memory
  __device__ __forceinline__ int q0(int t) const { return t ? q0B : q0A; }
  __device__ __forceinline__ int jb(int t) const { return t ? jbB : 0; }
  __device__ __forceinline__ int ns(int t) const { return t ? nsB : nsA; }
};

struct At3Geom {
  int ntiles, nblk, npairs, p_total, total;
  __device__ __forceinline__ At3Geom(int n, int H, int B) {
    ntiles = (n + AT3_BQ - 1) / AT3_BQ;
    nblk = (n + AT3_BK - 1) / AT3_BK;
    npairs = ntiles >> 1;
    p_total = B * H * npairs;
    total = p_total + ((ntiles & 1) ? B * H : 0);
  }
  __device__ __forceinline__ At3Item item(int idx, int H) const {
    At3Item it;
    int bh, tile0;
    if (idx < p_total) {
      bh = idx / npairs;
      tile0 = 2 * (idx - bh * npairs);
      it.lone = false;
    } else {
      bh = idx - p_total;
      tile0 = ntiles - 1;
      it.lone = true;
    }
    it.b = bh / H;
    it.h = bh - it.b * H;
    it.q0A = tile0 * AT3_BQ;
    it.q0B = it.lone ? it.q0A : it.q0A + AT3_BQ;
    it.jbB = it.lone ? (nblk + 1) / 2 : 0;
    it.nsA = it.lone ? (nblk + 1) / 2 : nblk;
    it.nsB = nblk - it.jbB;
    return it;
  }
  // position of (step i, slot t) in the order K/V tiles travel through the ring, relative to the item's first tile
  __device__ __forceinline__ int ring_pos(const At3Item& it, int i, int t) const {
    return it.lone ? (i < it.nsB ? 2 * i + t : it.nsB + i) : i;
  }
};

// narrow TMEM accesses for the (rare) rescale of the running output: 64 scores per thread are live at that point
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]),
               "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}

__device__ __forceinline__ void at3_bar_sync(int id, int count) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ void at3_bar_arrive(int id, int count) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory");
}

// masked / biased chunk (ragged tail or non-zero key bias): bv = this lane's key bias * log2e (lane = key inside the chunk)
__device__ __forceinline__ float at3_chunk_max_m(const uint32_t (&v)[32], float scale, float bv, int lim) {
  float m = -INFINITY;
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    const float x = fmaf(__uint_as_float(v[i]), scale, __shfl_sync(0xffffffffu, bv, i));
    m = fmaxf(m, i < lim ? x : -INFINITY);
  }
  return m;
}
__device__ __forceinline__ void at3_chunk_ex2_m(uint32_t (&v)[32], float scale, float neg_ref, float bv, int lim) {
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    // scale and reference in ONE fma exactly as in the unbiased path and the bias added afterwards: a key whose bias is 0
    // gets bit-identical probabilities whichever path its chunk takes
    const float x = fmaf(__uint_as_float(v[i]), scale, neg_ref) + __shfl_sync(0xffffffffu, bv, i);
    const float y = ex2_approx(x);
    v[i] = __float_as_uint(i < lim ? y : 0.0f);  // stale TMEM columns past the last real key must not reach P
  }
}

// ---- packed fp32 pairs (FFMA2 / FADD2, sm_100): half the FMA-pipe issue slots of the scale-and-subtract and of the row
//      sums; the 64 scores of a thread arrive from tcgen05.ld in consecutive registers, so (2i, 2i+1) are aligned pairs ----
// 2^x for a pair on the FMA pipe (same Cody-Waite split and degree-3 polynomial as ex2_poly3 in attention2.cuh)
__device__ __forceinline__ void ex2_poly3_x2(float& y0, float& y1, float x0, float x1) {
  x0 = fmaxf(x0, -125.0f);
  x1 = fmaxf(x1, -125.0f);
  float t0, t1, n0, n1, f0, f1, p0, p1;
  fadd2(t0, t1, x0, x1, 12582912.0f, 12582912.0f);
  fadd2(n0, n1, t0, t1, -12582912.0f, -12582912.0f);
  ffma2(f0, f1, n0, n1, -1.0f, -1.0f, x0, x1);
  ffma2(p0, p1, f0, f1, 0.05508868f, 0.05508868f, 0.24260405f, 0.24260405f);
  ffma2(p0, p1, p0, p1, f0, f1, 0.69327623f, 0.69327623f);
  ffma2(p0, p1, p0, p1, f0, f1, 0.99992895f, 0.99992895f);
  y0 = __int_as_float(__float_as_int(p0) + (__float_as_int(t0) << 23));
  y1 = __int_as_float(__float_as_int(p1) + (__float_as_int(t1) << 23));
}
// exponentials of a full unbiased 32-column chunk, in place; NPOLY of its 16 pairs take the FMA-pipe polynomial
template <int NPOLY>
__device__ __forceinline__ void at3_chunk_ex2_x2(uint32_t (&v)[32], float scale, float neg_ref) {
#pragma unroll
  for (int i = 0; i < 32; i += 2) {
    float x0, x1, y0, y1;
    ffma2(x0, x1, __uint_as_float(v[i]), __uint_as_float(v[i + 1]), scale, scale, neg_ref, neg_ref);
    // spread the polynomial pairs evenly over the chunk
    const bool poly = NPOLY > 0 && (((i >> 1) * NPOLY) % 16) + NPOLY > 15;
    if (poly) {
      ex2_poly3_x2(y0, y1, x0, x1);
    } else {
      y0 = ex2_approx(x0);
      y1 = ex2_approx(x1);
    }
    v[i] = __float_as_uint(y0);
    v[i + 1] = __float_as_uint(y1);
  }
}
__device__ __forceinline__ float at3_chunk_sum_pack_x2(const uint32_t (&v)[32], uint32_t (&pk)[16]) {
  float a0 = 0.f, a1 = 0.f, b0 = 0.f, b1 = 0.f;
#pragma unroll
  for (int i = 0; i < 32; i += 4) {
    fadd2(a0, a1, a0, a1, __uint_as_float(v[i]), __uint_as_float(v[i + 1]));
    fadd2(b0, b1, b0, b1, __uint_as_float(v[i + 2]), __uint_as_float(v[i + 3]));
    pk[i >> 1] = pack_bf16x2(__uint_as_float(v[i]), __uint_as_float(v[i + 1]));
    pk[(i >> 1) + 1] = pack_bf16x2(__uint_as_float(v[i + 2]), __uint_as_float(v[i + 3]));
  }
  return (a0 + a1) + (b0 + b1);
}

// VAR (UVLT_ATTN_POLY): 0 scalar FFMA / FADD, all exponentials on the MUFU; 1 the same with every fourth exponential on
// the FMA pipe; 2 packed FFMA2 / FADD2; 3 / 4 / 5 packed + 4 / 6 / 8 of every 16 pairs on the FMA pipe
template <int VAR>
static __global__ void __launch_bounds__(AT3_THREADS, 1)
attention3_kernel(const __grid_constant__ CUtensorMap tma_qkv, const __grid_constant__ CUtensorMap tma_out,
                  const Attn3Params p) {
  extern __shared__ __align__(1024) uint8_t att3_smem[];
  uint8_t* const smem = att3_smem;
  uint8_t* const sQ = smem + Attn3Smem::OFF_Q;
  uint8_t* const sKV = smem + Attn3Smem::OFF_KV;
  uint64_t* const bars = reinterpret_cast<uint64_t*>(smem + Attn3Smem::OFF_BAR);
  uint64_t* const q_full = bars;                      // [2] the slot's Q tile landed
  uint64_t* const q_empty = q_full + 2;               // [2] the item's last QK of the slot has drained: Q buffer reusable
  uint64_t* const kv_full = q_empty + 2;              // [AT3_STAGES]
  uint64_t* const kv_empty = kv_full + AT3_STAGES;    // [AT3_STAGES] two arrivals: the PV MMAs that read the stage have drained
  uint64_t* const s_full = kv_empty + AT3_STAGES;     // [2] S_t landed in TMEM
  uint64_t* const s_free = s_full + 2;                // [2] the slot's 8 softmax warps hold S_t in registers
  uint64_t* const p_full = s_free + 2;                // [2] P_t stored (and O_t rescaled), 8 warp arrivals
  uint64_t* const pv_done = p_full + 2;               // [2] PV_t drained: P_t reusable, O_t includes the block
  uint32_t* const tmem_slot = reinterpret_cast<uint32_t*>(pv_done + 2);

  const int warp = uniform_warp_idx();
  const int lane = threadIdx.x & 31;
  const int D = p.H * ATT_D;
  const At3Geom geo(p.n, p.H, p.B);
  const int nblk = geo.nblk;
  const int G = gridDim.x;
  TRACE_DECL;

  if (warp == AT3_SM_WARPS && lane == 0) {
    if (smem_u32(smem) & 1023u) __trap();
    tma_prefetch_desc(&tma_qkv);
    tma_prefetch_desc(&tma_out);
    for (int t = 0; t < 2; ++t) {
      mbar_init(&q_full[t], 1);
      mbar_init(&q_empty[t], 1);
      mbar_init(&s_full[t], 1);
      mbar_init(&s_free[t], 8);
      mbar_init(&p_full[t], 8);
      mbar_init(&pv_done[t], 1);
    }
    for (int s = 0; s < AT3_STAGES; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 2);
    }
    fence_mbar_init();
  }
  if (warp == AT3_SM_WARPS + 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();  // qkv and the bias come from earlier kernels of the chain
  pdl_trigger();

  // Register budget: 65536 / 640 -> 96 per thread at launch; the producer / MMA warpgroup gives most of its share back
  // and the four softmax warpgroups (64 scores per thread live in registers) take it.
#if AT3_SETMAXNREG
  if (warp >= AT3_SM_WARPS) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 112;");
  }
#endif
  if (warp == AT3_SM_WARPS) {
    // ---------------- TMA producer (converged warp, one elected lane issues) ----------------
    int qn[2] = {0, 0};  // Q tiles loaded so far per buffer
    int ord = 0;
#pragma unroll 1
    for (int idx = blockIdx.x; idx < geo.total; idx += G, ++ord) {
      const At3Item it = geo.item(idx, p.H);
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        if (it.ns(t) == 0) continue;
        mbar_wait_trap(&q_empty[t], (qn[t] & 1) ^ 1);
        if (elect_one_sync()) {
          mbar_expect_tx(&q_full[t], Attn3Smem::Q_BYTES);
          tma_load_3d(sQ + t * Attn3Smem::Q_BYTES, &tma_qkv, &q_full[t], it.h * ATT_D, it.q0(t), it.b);
        }
        __syncwarp();
        ++qn[t];
      }
      int c = ord * nblk;
#pragma unroll 1
      for (int i = 0; i < it.nsA; ++i) {
#pragma unroll 1
        for (int t = 0; t < (it.lone ? 2 : 1); ++t) {
          if (i >= it.ns(t)) continue;
          const int j = it.jb(t) + i;
          const int s = c % AT3_STAGES;
          mbar_wait_trap(&kv_empty[s], ((c / AT3_STAGES) & 1) ^ 1);
          if (elect_one_sync()) {
            uint8_t* const dst = sKV + s * Attn3Smem::STAGE_BYTES;
            mbar_expect_tx(&kv_full[s], Attn3Smem::STAGE_BYTES);
            tma_load_3d(dst, &tma_qkv, &kv_full[s], D + it.h * ATT_D, j * AT3_BK, it.b);
            tma_load_3d(dst + Attn3Smem::KV_BYTES, &tma_qkv, &kv_full[s], 2 * D + it.h * ATT_D, j * AT3_BK, it.b);
          }
          __syncwarp();
          ++c;
        }
      }
    }
  } else if (warp == AT3_SM_WARPS + 1 || warp == AT3_SM_WARPS + 2) {
    // ---------------- MMA issuer of slot t (one converged warp per slot, one elected lane issues) ----------------
    const int t = warp - (AT3_SM_WARPS + 1);
    const int last_valid = p.n - (nblk - 1) * AT3_BK;  // keys in the sequence's last block
    const uint32_t idesc_full = umma_idesc_bf16(AT3_BQ, AT3_BK, 0);
    const uint32_t idesc_last = umma_idesc_bf16(AT3_BQ, (last_valid + 15) & ~15, 0);  // UMMA N granularity 16
    constexpr uint32_t idesc_pv = umma_idesc_bf16(AT3_BQ, ATT_D, 1);
    const uint64_t qd = umma_smem_desc_sw128(smem_u32(sQ + t * Attn3Smem::Q_BYTES), 1024, 0);
    const uint64_t kd_base = umma_smem_desc_sw128(smem_u32(sKV), 1024, 0);
    // V: [key][64] rows of 128 B = MN-major B operand; 16 keys = 16 rows = 2048 B (+128 in the 16-byte address field)
    const uint64_t vd_base = umma_smem_desc_sw128(smem_u32(sKV + Attn3Smem::KV_BYTES), 1024, 1024);
    constexpr uint64_t STAGE_STEP = Attn3Smem::STAGE_BYTES >> 4;
    const uint32_t tS = tmem_base + t * 256;
    const uint32_t tP = tS + 128;
    const uint32_t tO = tS + 192;
    int k_qk = 0;   // QK blocks issued so far by this slot (all items)
    int k_pv = 0;   // PV blocks issued so far
    int qn = 0;     // Q tiles consumed so far
    int seen = 0;   // next ring position whose kv_full phase this warp has not observed yet (every phase is observed in
                    // order, also the other slot's tiles of LONE items: a parity wait that skips a phase can alias)
    // QK cursor: runs one block ahead of the PV loop, across item boundaries
    int c_idx = blockIdx.x, c_ord = 0, c_i = 0;
    At3Item c_it = geo.item(c_idx < geo.total ? c_idx : 0, p.H);
    auto c_seek = [&]() {
      while (c_idx < geo.total) {
        c_it = geo.item(c_idx, p.H);
        if (c_it.ns(t) > 0) break;
        c_idx += G;
        ++c_ord;
      }
      c_i = 0;
    };
    auto issue_next_qk = [&]() {
      if (c_idx >= geo.total) return;
      if (c_i == 0) {
        mbar_wait_trap(&q_full[t], qn & 1);
        ++qn;
      }
      const int c = c_ord * nblk + geo.ring_pos(c_it, c_i, t);
      while (seen <= c) {
        mbar_wait_trap(&kv_full[seen % AT3_STAGES], (seen / AT3_STAGES) & 1);
        ++seen;
      }
      if (k_qk > 0) mbar_wait_trap(&s_free[t], (k_qk - 1) & 1);  // the softmax warps hold the previous S_t in registers
      tc_fence_after();
      const uint64_t kd = kd_base + STAGE_STEP * (c % AT3_STAGES);
      const uint32_t idesc = (c_it.jb(t) + c_i == nblk - 1) ? idesc_last : idesc_full;
      const bool last_of_item = (c_i == c_it.ns(t) - 1);
      if (elect_one_sync()) {
        umma_bf16_ss(tS, qd, kd, idesc, 0u);
        umma_bf16_ss(tS, qd + 2, kd + 2, idesc, 1u);
        umma_bf16_ss(tS, qd + 4, kd + 4, idesc, 1u);
        umma_bf16_ss(tS, qd + 6, kd + 6, idesc, 1u);
        umma_commit(&s_full[t]);
        if (last_of_item) umma_commit(&q_empty[t]);  // the producer may fetch the next item's Q tile
      }
      __syncwarp();
      if (lane == 0) TRACE_PT(0x300 + t * 0x100 + 0x10 + (k_qk & 15));  // QK issued
      ++k_qk;
      if (++c_i == c_it.ns(t)) {
        c_idx += G;
        ++c_ord;
        c_seek();
      }
    };
    c_seek();
    issue_next_qk();
    int ord = 0;
#pragma unroll 1
    for (int idx = blockIdx.x; idx < geo.total; idx += G, ++ord) {
      const At3Item it = geo.item(idx, p.H);
      const int ns = it.ns(t);
#pragma unroll 1
      for (int i = 0; i < ns; ++i) {
        issue_next_qk();
        mbar_wait_trap(&p_full[t], k_pv & 1);
        tc_fence_after();
        if (lane == 0) TRACE_PT(0x300 + t * 0x100 + 0x20 + (k_pv & 15));  // p_full seen
        const int c = ord * nblk + geo.ring_pos(it, i, t);
        const uint64_t vd = vd_base + STAGE_STEP * (c % AT3_STAGES);
        const int ksteps = (it.jb(t) + i == nblk - 1) ? ((last_valid + 15) >> 4) : (AT3_BK / 16);
        if (elect_one_sync()) {
#pragma unroll
          for (int k = 0; k < AT3_BK / 16; ++k)
            if (k < ksteps) umma_bf16_ts(tO, tP + 8 * k, vd + 128 * k, idesc_pv, (i > 0 || k > 0) ? 1u : 0u);
          umma_commit(&kv_empty[c % AT3_STAGES]);
          if (it.lone) umma_commit(&kv_empty[c % AT3_STAGES]);  // LONE: the stage belongs to this slot alone
          umma_commit(&pv_done[t]);
        }
        __syncwarp();
        if (lane == 0) TRACE_PT(0x300 + t * 0x100 + 0x30 + (k_pv & 15));  // PV issued
        ++k_pv;
      }
    }
    if (lane == 0) TRACE_FLUSH();
  } else if (warp < AT3_SM_WARPS) {
    // ---------------- softmax / correction / epilogue warps ----------------
    // Register discipline: 64 scores per thread are live through a block, the cap is 96 registers (640 threads), and a
    // spill is expensive here (213 KB of shared memory leave ~28 KB of L1, so local memory lives in L2: the first
    // version spent ~8k cycles per work item reloading spilled loop-invariant values in its epilogue).  So nothing that
    // can be re-derived is kept across the block loop: thread coordinates come from a volatile %tid.x read at each
    // point of use (ptxas cannot hoist it), shared memory is addressed through 32-bit shared-space offsets, the work
    // item is decoded again in the epilogue.
    constexpr int XO_FULL = 11, XO_FREE = 12, EXP_TURN = 13;  // EXP_TURN + slot: that slot may run its exponentials
    constexpr int NPOLY = VAR == 3 ? 4 : VAR == 4 ? 6 : VAR == 5 ? 8 : 0;
    auto tid_now = []() {
      uint32_t v;
      asm volatile("mov.u32 %0, %%tid.x;" : "=r"(v));
      return v;
    };
    // tid bits: [4:0] lane, [6:5] TMEM lane quadrant, [7] half of the block's key columns, [8] slot; row = tid & 127
    auto tmem_s = [&](uint32_t tid) {  // this thread's first S column; P: + 128 - half * 32, O: + 192 - half * 32
      return tmem_base + ((tid >> 8) & 1) * 256 + ((tid & 96u) << 16) + ((tid >> 7) & 1) * 64;
    };
    const uint32_t sb = smem_u32(smem);
    const int t = warp >> 3;  // slot (warp-uniform)
    int k_blk = 0;            // key blocks this slot has processed so far (all items)
    bool first_merge = true;
    if (p.pingpong && t == 1) at3_bar_arrive(EXP_TURN, 512);  // slot 0 goes first
#pragma unroll 1
    for (int idx = blockIdx.x; idx < geo.total; idx += G) {
      int ns, jb;
      const float* bb;
      {
        const At3Item it = geo.item(idx, p.H);
        ns = it.ns(t);
        jb = it.jb(t);
        bb = p.bias ? p.bias + static_cast<long long>(it.b) * p.n : nullptr;
      }
      float m_run = -INFINITY;  // reference of the running sum / output (scaled log2 domain)
      float l_run = 0.0f;       // this half's partial row sum
#pragma unroll 1
      for (int i = 0; i < ns; ++i, ++k_blk) {
        const int hf = (warp >> 2) & 1;
        const int kv_valid = min(AT3_BK, p.n - (jb + i) * AT3_BK);
        const int hv16 = min(64, max(0, ((kv_valid + 15) & ~15) - hf * 64));  // columns of this half the PV MMA may read
        const int nchunk = (hv16 + 31) >> 5;
        const int lim0 = kv_valid - hf * 64;  // real keys in chunk 0 of this half (chunk 1: lim0 - 32); may be <= 0 or >= 32
        // key bias of this half's two chunks (lane = key): the loads are ISSUED here, ahead of the wait for S, and first
        // USED after the scores are in registers (used right away -- the first version scaled them here -- every block paid
        // the L2 round trip: 5 % of the kernel's stall samples inside the engine, profiles/r02_b32_attn_full.md)
        float bv0 = 0.0f, bv1 = 0.0f;
        if (bb) {
          const int k0 = (jb + i) * AT3_BK + hf * 64 + lane;
          if (k0 < p.n) bv0 = __ldg(bb + k0);
          if (k0 + 32 < p.n) bv1 = __ldg(bb + k0 + 32);
        }
        const float scale = p.scale_log2;
        uint32_t v[2][32], pk[16];
        mbar_wait_trap(&s_full[t], k_blk & 1);
        tc_fence_after();
        if (warp == 0 && lane == 0) TRACE_PT(0x500 + 0x10 + (k_blk & 15));  // s_full seen
        if (nchunk > 0) {
          tmem_ld64(tmem_s(tid_now()), v[0], v[1]);
          tmem_wait_ld_dep2(v[0], v[1]);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_free[t]);  // QK of the slot's next block may overwrite S
        bv0 *= ATT_LOG2E;
        bv1 *= ATT_LOG2E;
        const bool gen0 = __any_sync(0xffffffffu, bv0 != 0.0f) || lim0 < 32;  // chunk-uniform path selection
        const bool gen1 = __any_sync(0xffffffffu, bv1 != 0.0f) || lim0 < 64;
        // ---- block maximum: own half, then the row's other half through shared memory ----
        float m_half = -INFINITY;
        if (nchunk > 0) m_half = gen0 ? at3_chunk_max_m(v[0], scale, bv0, lim0) : at2_chunk_max<0>(v[0], scale, 0, 32) * scale;
        if (nchunk > 1)
          m_half = fmaxf(m_half, gen1 ? at3_chunk_max_m(v[1], scale, bv1, lim0 - 32) : at2_chunk_max<0>(v[1], scale, 0, 32) * scale);
        float m_blk;
        {
          // [2 parities][2 slots][2 halves][128 rows]; the other half of this row sits 512 bytes away (half bit = tid bit 7)
          const uint32_t tid = tid_now();
          const uint32_t own = sb + Attn3Smem::OFF_MX + (k_blk & 1) * 2048 + (tid & 511u) * 4;
          asm volatile("st.shared.f32 [%0], %1;" ::"r"(own), "f"(m_half) : "memory");
          at3_bar_sync(1 + ((tid >> 8) & 1) * 4 + ((tid >> 5) & 3), 64);  // the two warps (halves) of this (slot, quadrant)
          m_blk = fmaxf(m_half, lds_f32(own ^ 512u));
        }
        const float ref = (m_blk > m_run + 8.0f) ? m_blk : m_run;  // m_run = -inf on the first block -> m_blk
        const float alpha = (ref == m_run) ? 1.0f : ex2_approx(m_run - ref);  // 0 on the first block
        const float neg_ref = -ref;
        if (warp == 0 && lane == 0) TRACE_PT(0x500 + 0x20 + (k_blk & 15));  // reference known
        if (i > 0) {
          // P_t and O_t are free once PV_t of the previous block has drained (issued a block ago: checked here,
          // ahead of the turn, so that the wait and the rare rescale of O overlap the other slot's exponentials)
          mbar_wait_trap(&pv_done[t], (k_blk - 1) & 1);
          tc_fence_after();
          if (warp == 0 && lane == 0) TRACE_PT(0x500 + 0x30 + (k_blk & 15));  // pv_done seen
          if (__any_sync(0xffffffffu, alpha != 1.0f)) {  // bring this half's 32 output columns to the new reference (rare)
            const uint32_t tO = tmem_s(tid_now()) + 192 - hf * 32;
#pragma unroll 1
            for (int cc = 0; cc < 32; cc += 8) {
              uint32_t o[8];
              tmem_ld8(tO + cc, o);
              tmem_wait_ld();
#pragma unroll
              for (int q = 0; q < 8; ++q) o[q] = __float_as_uint(__uint_as_float(o[q]) * alpha);
              tmem_st8(tO + cc, o);
            }
          }
        }
        // The exponentials of the two slots take turns (token = named barriers EXP_TURN + slot; per-quadrant mbarrier
        // tokens measured slower, 75.8 against 70.6 us: four more polling warps per scheduler): left alone the slots
        // drift into phase, all four softmax warps of a scheduler then want the MUFU at the same time (a block's
        // exponentials take 2300 cycles instead of 1024) and leave it idle together afterwards (in-kernel timeline,
        // profiles/r02_attention.md).  In turns, one slot's exponentials run under the other slot's TMEM loads, maxima,
        // sums, packs and P stores.
        if (p.pingpong) at3_bar_sync(EXP_TURN + t, 512);
        if (warp == 0 && lane == 0) TRACE_PT(0x500 + 0x70 + (k_blk & 15));  // turn acquired
        // ---- exponentials (in place), then sums / packs / P stores; `p.zero` (always 0) gives ptxas a branch between
        //      the stages, i.e. separate scheduling regions (see at2_chunk_ex2) ----
        if (nchunk > 0) {
          if (gen0) at3_chunk_ex2_m(v[0], scale, neg_ref, bv0, lim0);
          else if (VAR >= 2) at3_chunk_ex2_x2<NPOLY>(v[0], scale, neg_ref);
          else at2_chunk_ex2<0, VAR == 1>(v[0], scale, neg_ref, 0, 32);
        }
        if (p.zero) break;
        if (nchunk > 1) {
          if (gen1) at3_chunk_ex2_m(v[1], scale, neg_ref, bv1, lim0 - 32);
          else if (VAR >= 2) at3_chunk_ex2_x2<NPOLY>(v[1], scale, neg_ref);
          else at2_chunk_ex2<0, VAR == 1>(v[1], scale, neg_ref, 0, 32);
        }
        if (p.pingpong) at3_bar_arrive(EXP_TURN + (t ^ 1), 512);
        if (warp == 0 && lane == 0) TRACE_PT(0x500 + 0x80 + (k_blk & 15));  // exponentials done, turn passed
        float l_blk = 0.0f;
        if (nchunk > 0) {
          l_blk += VAR >= 2 ? at3_chunk_sum_pack_x2(v[0], pk) : at2_chunk_sum_pack(v[0], pk);
          tmem_st16(tmem_s(tid_now()) + 128 - hf * 32, pk);
        }
        if (p.zero) break;
        if (nchunk > 1) {
          l_blk += VAR >= 2 ? at3_chunk_sum_pack_x2(v[1], pk) : at2_chunk_sum_pack(v[1], pk);
          tmem_st16(tmem_s(tid_now()) + 128 - hf * 32 + 16, pk);
        }
        l_run = l_run * alpha + l_blk;
        m_run = ref;
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[t]);
        if (warp == 0 && lane == 0) TRACE_PT(0x500 + 0x40 + (k_blk & 15));  // P stored
      }
      if (p.pingpong && t == 1) {
        // slot 1 has fewer key blocks than slot 0 in a LONE item of an odd block count: pass the turns it does not use
        const At3Item it1 = geo.item(idx, p.H);
        for (int extra = it1.nsA - it1.nsB; extra > 0; --extra) {
          at3_bar_sync(EXP_TURN + 1, 512);
          at3_bar_arrive(EXP_TURN, 512);
        }
      }
      if (ns == 0) continue;  // LONE item of a one-block sequence: slot 1 has nothing to do (and nothing to merge)
      // ---------------- item epilogue: O / l -> bf16 -> staging tile -> one TMA bulk store ----------------
      mbar_wait_trap(&pv_done[t], (k_blk - 1) & 1);
      tc_fence_after();
      if (warp == 0 && lane == 0) TRACE_PT(0x500 + 0x61);  // last PV done
      const uint32_t tid = tid_now();
      const uint32_t hf = (tid >> 7) & 1, row = tid & 127u;
      const bool store_thread = (tid & 255u) == 0;  // first thread of the slot
      const int slot_bar = 9 + t;                    // the slot's 8 softmax warps
      if (store_thread) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // staging tile free again
      const At3Item it = geo.item(idx, p.H);
      const bool merge = it.lone && it.nsB > 0;
      const uint32_t tO = tmem_s(tid) + 192 - hf * 32;
      const uint32_t ls_row = sb + Attn3Smem::OFF_LS + row * 4;       // + (slot * 2 + half) * 512
      const uint32_t xo = sb + Attn3Smem::OFF_ST1 + row * 16;         // + float4 index * 2048
      float inv_l, a_own = 1.0f, a_peer = 0.0f;
      if (!merge) {
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(ls_row + (t * 2 + hf) * 512), "f"(l_run) : "memory");
        at3_bar_sync(slot_bar, 256);
        if (warp == 0 && lane == 0) TRACE_PT(0x500 + 0x62);  // row sums exchanged
        inv_l = 1.0f / (l_run + lds_f32(ls_row + (t * 2 + (hf ^ 1)) * 512));
      } else if (t == 1) {
        // slot 1 -> slot 0: chunk-major float4s (a warp writes 512 contiguous bytes per instruction)
        at3_bar_sync(slot_bar, 256);                       // slot 1's last TMA store has read the staging tile (= xo)
        if (!first_merge) at3_bar_sync(XO_FREE, 512);      // slot 0 has read the previous partial (xo, ms, ls)
        first_merge = false;
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(ls_row + (2 + hf) * 512), "f"(l_run) : "memory");
        if (hf == 0) asm volatile("st.shared.f32 [%0], %1;" ::"r"(sb + Attn3Smem::OFF_MS + row * 4), "f"(m_run) : "memory");
#pragma unroll 1
        for (int q = 0; q < 32; q += 8) {
          uint32_t o[8];
          tmem_ld8(tO + q, o);
          tmem_wait_ld();
          sts128(xo + ((hf * 32 + q) >> 2) * 2048, __uint_as_float(o[0]), __uint_as_float(o[1]), __uint_as_float(o[2]),
                 __uint_as_float(o[3]));
          sts128(xo + (((hf * 32 + q) >> 2) + 1) * 2048, __uint_as_float(o[4]), __uint_as_float(o[5]), __uint_as_float(o[6]),
                 __uint_as_float(o[7]));
        }
        tc_fence_before();
        at3_bar_arrive(XO_FULL, 512);
        continue;  // slot 0 stores the tile
      } else {
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(ls_row + hf * 512), "f"(l_run) : "memory");
        at3_bar_sync(XO_FULL, 512);
        const float m_peer = lds_f32(sb + Attn3Smem::OFF_MS + row * 4);
        const float l_own = l_run + lds_f32(ls_row + (hf ^ 1) * 512);
        const float l_peer = lds_f32(ls_row + 2 * 512) + lds_f32(ls_row + 3 * 512);
        const float m = fmaxf(m_run, m_peer);
        a_own = ex2_approx(m_run - m);
        a_peer = ex2_approx(m_peer - m);
        inv_l = 1.0f / (l_own * a_own + l_peer * a_peer);
      }
      {
        // 8 output columns at a time (one 16-byte chunk of the bf16 row): the epilogue keeps few registers live.
        // bf16 row -> staging tile, 16-byte chunks XOR-swizzled with row % 8: that IS the 128B-swizzle TMA layout, and
        // the writes are bank-conflict free
        const uint32_t st_row = sb + (t ? Attn3Smem::OFF_ST1 : Attn3Smem::OFF_ST0) + row * 128;
        a_own *= inv_l;
        a_peer *= inv_l;
#pragma unroll
        for (int q = 0; q < 32; q += 8) {
          uint32_t o[8];
          tmem_ld8(tO + q, o);
          tmem_wait_ld();
          float f[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(o[e]) * a_own;
          if (merge) {
            const float4 x0 = lds128(xo + ((hf * 32 + q) >> 2) * 2048), x1 = lds128(xo + (((hf * 32 + q) >> 2) + 1) * 2048);
            f[0] += x0.x * a_peer; f[1] += x0.y * a_peer; f[2] += x0.z * a_peer; f[3] += x0.w * a_peer;
            f[4] += x1.x * a_peer; f[5] += x1.y * a_peer; f[6] += x1.z * a_peer; f[7] += x1.w * a_peer;
          }
          sts128(st_row + ((((hf * 32 + q) >> 3) ^ (row & 7)) << 4), __uint_as_float(pack_bf16x2(f[0], f[1])),
                 __uint_as_float(pack_bf16x2(f[2], f[3])), __uint_as_float(pack_bf16x2(f[4], f[5])),
                 __uint_as_float(pack_bf16x2(f[6], f[7])));
        }
        if (merge) at3_bar_arrive(XO_FREE, 512);  // slot 1 may overwrite its partial
      }
      if (warp == 0 && lane == 0) TRACE_PT(0x500 + 0x63);  // staging rows written
      tc_fence_before();         // this item's TMEM reads are ordered before the p_full arrive of the next item's first block
      fence_proxy_async_smem();  // generic-proxy staging writes -> visible to the TMA (async proxy)
      if (warp == 0 && lane == 0) TRACE_PT(0x500 + 0x64);  // fenced
      at3_bar_sync(slot_bar, 256);
      if (warp == 0 && lane == 0) TRACE_PT(0x500 + 0x65);  // slot barrier passed
      if (store_thread) {
        // rows >= n are clipped by the tensor map
        asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                         reinterpret_cast<uint64_t>(&tma_out)),
                     "r"(sb + (t ? Attn3Smem::OFF_ST1 : Attn3Smem::OFF_ST0)), "r"(it.h * ATT_D), "r"(it.q0(t)), "r"(it.b)
                     : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
      if (warp == 0 && lane == 0) TRACE_PT(0x500 + 0x60);  // item stored
    }
    if ((tid_now() & 255u) == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // the writes are complete before the CTA exits
    tc_fence_before();
    const bool tracer = (warp == 0 && lane == 0);
    if (tracer) TRACE_FLUSH();
  }

  __syncthreads();
  if (warp == AT3_SM_WARPS + 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace uvlt
