// Fused multi-head attention over one packed QKV activation:
//     O[b, q, h, :] = softmax_k( Q[b,q,h,:] . K[b,k,h,:] * scale + bias[b,k] ) @ V[b,k,h,:]
// for the joint [cls | template | search | text] sequence of the ViT blocks (reference block.py:47-61: masked keys are
// filled with -1e10) and for the BERT text layers (bert_backbone.py:299-325: additive -10000 mask).  Both mask
// dialects are an additive per-key fp32 bias here (adding -1e10 to an O(10) fp32 score is exactly -1e10).
//
// QKV layout in HBM: bf16 [B, n, 3*H*64] exactly as the qkv GEMM writes it (token-major, [which][head][64]).
// One CTA = one (batch, head, 128-query tile).  192 threads:
//   warp 0      TMA producer : Q tile once, then K_j / V_j tiles (128 keys x 64) through a 2-stage ring
//   warp 1      MMA issuer   : S_j = Q K_j^T   (tcgen05, M=128, N<=128, K=64)  -> TMEM cols [0,128)
//                              O  += P_j V_j   (tcgen05, M=128, N=64,  K<=128) -> TMEM cols [128,192)
//                              V is consumed straight from its [key][64] tile as an MN-major B operand.
//   warps 2..5  softmax      : thread = query row; tcgen05.ld S, online softmax in the exp2 domain (fp32 stats),
//                              P_j -> bf16 -> swizzled smem (A operand of the PV MMA), rescale O in TMEM when the
//                              running max moves, final O / l -> bf16 -> HBM.
// 256 TMEM columns and ~100 KB smem per CTA -> two CTAs per SM, so one CTA's softmax overlaps the other's MMAs.
#pragma once
#include "common.cuh"

namespace uvlt {

constexpr int ATT_BQ = 128;    // queries per CTA
constexpr int ATT_BKV = 128;   // keys per block
constexpr int ATT_D = 64;      // head dim (both UVLTrack-B and -L)
constexpr int ATT_THREADS = 192;
constexpr int ATT_STAGES = 2;
constexpr int ATT_MAX_KV = 1280;  // bias staging (n <= 1193 at 384^2/384^2)

struct AttnParams {
  int n;               // sequence length (queries == keys)
  int H;               // heads
  float scale_log2;    // head_dim^-0.5 * log2(e)
  const float* bias;   // [B, n] additive key bias (natural-log domain), or nullptr
  __nv_bfloat16* out;  // [B, n, H*64]
};

struct AttnSmem {
  static constexpr int Q_BYTES = ATT_BQ * ATT_D * 2;       // 16 KB
  static constexpr int KV_BYTES = ATT_BKV * ATT_D * 2;     // 16 KB each
  static constexpr int P_BYTES = ATT_BQ * ATT_BKV * 2;     // 32 KB (two 64-wide K halves)
  static constexpr int OFF_Q = 0;
  static constexpr int OFF_K = OFF_Q + Q_BYTES;
  static constexpr int OFF_V = OFF_K + ATT_STAGES * KV_BYTES;
  static constexpr int OFF_P = OFF_V + ATT_STAGES * KV_BYTES;
  static constexpr int OFF_BIAS = OFF_P + P_BYTES;
  static constexpr int OFF_BAR = OFF_BIAS + ATT_MAX_KV * 4;
  static constexpr int TOTAL = OFF_BAR + 256 + 1024;
};

// V_KMAJOR = true is a bring-up alternative: V^T supplied as its own tensor [B, H, 64, n_pad] (keys contiguous).
template <bool V_KMAJOR>
__global__ void __launch_bounds__(ATT_THREADS, 2)
attention_kernel(const __grid_constant__ CUtensorMap tma_qkv, const __grid_constant__ CUtensorMap tma_vt,
                 const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem + AttnSmem::OFF_Q;
  uint8_t* sK = smem + AttnSmem::OFF_K;
  uint8_t* sV = smem + AttnSmem::OFF_V;
  uint8_t* sP = smem + AttnSmem::OFF_P;
  float* sBias = reinterpret_cast<float*>(smem + AttnSmem::OFF_BIAS);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + AttnSmem::OFF_BAR);
  uint64_t* q_full = bars + 0;
  uint64_t* kv_full = bars + 1;                  // [ATT_STAGES]
  uint64_t* kv_empty = kv_full + ATT_STAGES;     // [ATT_STAGES]
  uint64_t* s_full = kv_empty + ATT_STAGES;      // S_j landed in TMEM
  uint64_t* s_empty = s_full + 1;                // softmax finished reading S_j (128 arrivals)
  uint64_t* p_full = s_empty + 1;                // P_j in smem + O rescaled (128 arrivals)
  uint64_t* pv_done = p_full + 1;                // PV_j drained: P buffer reusable, O consistent
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * ATT_BQ;
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int D = p.H * ATT_D;
  const int nblk = (p.n + ATT_BKV - 1) / ATT_BKV;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_qkv);
    if (V_KMAJOR) tma_prefetch_desc(&tma_vt);
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
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  }
  // stage the key bias (pre-multiplied by log2 e); keys >= n get -inf so padded columns vanish
  for (int i = threadIdx.x; i < nblk * ATT_BKV; i += ATT_THREADS) {
    float v = -INFINITY;
    if (i < p.n) v = p.bias ? p.bias[static_cast<long long>(b) * p.n + i] * 1.4426950408889634f : 0.0f;
    sBias[i] = v;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_S = tmem_base;
  const uint32_t tmem_O = tmem_base + ATT_BKV;

  if (warp == 0) {
    if (lane == 0) {
      // ---------------- TMA producer ----------------
      mbar_expect_tx(q_full, AttnSmem::Q_BYTES);
      tma_load_3d(sQ, &tma_qkv, q_full, h * ATT_D, q0, b);
      for (int j = 0; j < nblk; ++j) {
        const int s = j % ATT_STAGES;
        const uint32_t ph = (j / ATT_STAGES) & 1;
        mbar_wait(&kv_empty[s], ph ^ 1);
        mbar_expect_tx(&kv_full[s], 2 * AttnSmem::KV_BYTES);
        tma_load_3d(sK + s * AttnSmem::KV_BYTES, &tma_qkv, &kv_full[s], D + h * ATT_D, j * ATT_BKV, b);
        if (V_KMAJOR) {
          // V^T tile [64 d rows x 128 keys] as two 64-key halves (128 B swizzle atoms)
          tma_load_3d(sV + s * AttnSmem::KV_BYTES, &tma_vt, &kv_full[s], j * ATT_BKV, h * ATT_D, b);
          tma_load_3d(sV + s * AttnSmem::KV_BYTES + AttnSmem::KV_BYTES / 2, &tma_vt, &kv_full[s],
                      j * ATT_BKV + 64, h * ATT_D, b);
        } else {
          tma_load_3d(sV + s * AttnSmem::KV_BYTES, &tma_qkv, &kv_full[s], 2 * D + h * ATT_D, j * ATT_BKV, b);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ---------------- MMA issuer ----------------
      auto issue_qk = [&](int j) {
        const int s = j % ATT_STAGES;
        const int kv_valid = min(ATT_BKV, p.n - j * ATT_BKV);
        const int ncols = (kv_valid + 15) & ~15;  // UMMA N granularity at M=128
        const uint32_t idesc = umma_idesc_bf16(ATT_BQ, ncols, 0);
        const uint64_t adesc = umma_smem_desc_sw128(smem_u32(sQ), 1024, 0);
        const uint64_t bdesc = umma_smem_desc_sw128(smem_u32(sK + s * AttnSmem::KV_BYTES), 1024, 0);
#pragma unroll
        for (int k = 0; k < ATT_D / 16; ++k) umma_bf16_ss(tmem_S, adesc + 2 * k, bdesc + 2 * k, idesc, k > 0);
        umma_commit(s_full);
      };
      mbar_wait(q_full, 0);
      mbar_wait(&kv_full[0], 0);
      tc_fence_after();
      issue_qk(0);
      for (int j = 0; j < nblk; ++j) {
        const int s = j % ATT_STAGES;
        if (j + 1 < nblk) {
          // S_j has been consumed by the softmax warps -> overwrite with S_{j+1} while they work on P_j
          mbar_wait(s_empty, j & 1);
          mbar_wait(&kv_full[(j + 1) % ATT_STAGES], ((j + 1) / ATT_STAGES) & 1);
          tc_fence_after();
          issue_qk(j + 1);
        }
        mbar_wait(p_full, j & 1);
        tc_fence_after();
        const int kv_valid = min(ATT_BKV, p.n - j * ATT_BKV);
        const int ksteps = (kv_valid + 15) >> 4;
        constexpr uint32_t idesc_pv_mn = umma_idesc_bf16(ATT_BQ, ATT_D, 1);
        constexpr uint32_t idesc_pv_k = umma_idesc_bf16(ATT_BQ, ATT_D, 0);
        const uint32_t p_addr = smem_u32(sP);
        const uint32_t v_addr = smem_u32(sV + s * AttnSmem::KV_BYTES);
        for (int k = 0; k < ksteps; ++k) {
          // P: two [128 x 64] K-major halves; 16 keys = 32 B inside the swizzle atom
          const uint64_t adesc = umma_smem_desc_sw128(p_addr + (k >> 2) * (AttnSmem::P_BYTES / 2), 1024, 0) + 2 * (k & 3);
          uint64_t bdesc;
          if (V_KMAJOR) {
            // V^T halves: [64 d rows x 64 keys] K-major
            bdesc = umma_smem_desc_sw128(v_addr + (k >> 2) * (AttnSmem::KV_BYTES / 2), 1024, 0) + 2 * (k & 3);
            umma_bf16_ss(tmem_O, adesc, bdesc, idesc_pv_k, (j > 0 || k > 0) ? 1u : 0u);
          } else {
            // V: [key][64] rows of 128 B = MN-major B operand; 16 keys = 16 rows = 2048 B
            bdesc = umma_smem_desc_sw128(v_addr + k * 2048, 1024, 1024);
            umma_bf16_ss(tmem_O, adesc, bdesc, idesc_pv_mn, (j > 0 || k > 0) ? 1u : 0u);
          }
        }
        umma_commit(&kv_empty[s]);
        umma_commit(pv_done);
      }
    }
  } else {
    // ---------------- softmax / correction / epilogue warps ----------------
    const int lane_grp = warp & 3;
    const int row = lane_grp * 32 + lane;  // query row inside the tile == TMEM lane
    const uint32_t lane_off = static_cast<uint32_t>(lane_grp * 32) << 16;
    float m_run = -INFINITY;
    float l_run = 0.0f;
    for (int j = 0; j < nblk; ++j) {
      const int kv_valid = min(ATT_BKV, p.n - j * ATT_BKV);
      const int nchunk = (kv_valid + 31) >> 5;
      const float* bj = sBias + j * ATT_BKV;
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      // pass 1: block maximum of the biased, log2-scaled scores
      float m_blk = -INFINITY;
      for (int c = 0; c < nchunk; ++c) {
        uint32_t v[32];
        tmem_ld32(tmem_S + lane_off + c * 32, v);
        tmem_wait_ld();
        const int lim = kv_valid - c * 32;  // columns >= lim were never written by the MMA (stale TMEM)
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float x = fmaf(__uint_as_float(v[i]), p.scale_log2, bj[c * 32 + i]);
          m_blk = fmaxf(m_blk, i < lim ? x : -INFINITY);
        }
      }
      const float m_new = fmaxf(m_run, m_blk);          // finite: every block has >= 1 real, unmasked-or-finite key
      const float alpha = exp2f(m_run - m_new);          // 0 on the first block (m_run = -inf)
      // P buffer (and O) are free once PV_{j-1} has drained
      if (j > 0) mbar_wait(pv_done, (j - 1) & 1);
      tc_fence_after();
      // pass 2: probabilities -> bf16 -> swizzled smem, row sum in fp32
      float l_blk = 0.0f;
      for (int c = 0; c < nchunk; ++c) {
        uint32_t v[32];
        tmem_ld32(tmem_S + lane_off + c * 32, v);
        tmem_wait_ld();
        uint32_t pk[16];
        const int lim = kv_valid - c * 32;
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          float p0 = exp2f(fmaf(__uint_as_float(v[i]), p.scale_log2, bj[c * 32 + i]) - m_new);
          float p1 = exp2f(fmaf(__uint_as_float(v[i + 1]), p.scale_log2, bj[c * 32 + i + 1]) - m_new);
          p0 = i < lim ? p0 : 0.0f;       // stale TMEM columns past the last real key must not reach P
          p1 = i + 1 < lim ? p1 : 0.0f;
          l_blk += p0 + p1;
          pk[i >> 1] = pack_bf16x2(p0, p1);
        }
        // 32 keys = four 16 B chunks of this row; chunk index inside the 64-key half is XOR-swizzled with row%8
        uint8_t* half_base = sP + (c >> 1) * (AttnSmem::P_BYTES / 2) + row * 128;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int chunk = ((c & 1) * 4 + q) ^ (row & 7);
          *reinterpret_cast<uint4*>(half_base + chunk * 16) =
              make_uint4(pk[q * 4], pk[q * 4 + 1], pk[q * 4 + 2], pk[q * 4 + 3]);
        }
      }
      // a partially filled 16-key MMA step may read up to the next 32-key boundary: already covered (nchunk*32)
      tc_fence_before();
      mbar_arrive(s_empty);
      // rescale the running output when the maximum moved (skipped warp-wide when nobody needs it)
      if (j > 0 && !__all_sync(0xffffffffu, alpha == 1.0f)) {
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
      l_run = l_run * alpha + l_blk;
      m_run = m_new;
      fence_proxy_async_smem();  // generic-proxy P writes -> visible to the tensor core (async proxy)
      tc_fence_before();
      mbar_arrive(p_full);
    }
    // ---------------- epilogue: O / l -> bf16 ----------------
    mbar_wait(pv_done, (nblk - 1) & 1);
    tc_fence_after();
    const int q = q0 + row;
    const float inv_l = 1.0f / l_run;
#pragma unroll
    for (int c = 0; c < ATT_D; c += 32) {
      uint32_t o[32];
      tmem_ld32(tmem_O + lane_off + c, o);
      tmem_wait_ld();
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
    tc_fence_before();
  }

  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

}  // namespace uvlt
