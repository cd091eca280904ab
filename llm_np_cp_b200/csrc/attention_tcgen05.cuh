// Grouped-query causal attention for prefill on sm_100a tensor cores (tcgen05 + TMEM + TMA).
//
// Replaces, per layer, the reference's materialised chain
//   repeat_kv -> Q.K^T / sqrt(HD) -> where(tril) -> softmax -> P.V -> transpose/reshape
// (llama3.2_model_numpy.py:470-511; gemma2_model.py:536-577) with one kernel that never writes the
// [B,NH,S,T] score tensor.  Options the reference lacks but BASELINE.json asks for are in-kernel
// predicates/epilogues: sliding window and tanh soft-cap (semantics of transformers' modeling_gemma2).
//
// CTA = one (batch, q-head, 128-row q tile); 192 threads:
//   warps 0-3  softmax/correction/epilogue: thread = q row (= TMEM lane)
//   warp 4     TMA producer (Q once, K/V tiles through a 2-stage ring)
//   warp 5     tcgen05.mma issuer:  S = Q.K^T (both K-major)  ->  TMEM[0, BNKV)
//                                   O += P.V  (P K-major from smem, V MN-major) -> TMEM[BNKV, BNKV+HD)
// GQA: kv head = q head / n_rep, no K/V replication (replaces repeat_kv_np, :188-204).
// Online softmax in the exp2 domain with lazy rescaling of the TMEM-resident O accumulator.
#pragma once
#include "common.cuh"

namespace b2 {

struct AttnParams {
  int B, S, NH, NKV, Tmax;
  const int* t0_ptr;     // tokens already cached before this call (device scalar)
  float scale;           // 1/sqrt(head_dim) (reference) or query_pre_attn_scalar^-0.5 (HF flag)
  float softcap;         // 0 = off
  int causal;            // reference quirk: 0 when q_len <= 2 (llama3.2_model_numpy.py:479)
  int window;            // 0 = off; else keep kv > qpos - window
  __nv_bfloat16* out;    // [B*S, NH*HD]
  int ldo;
  int v_lbo, v_sbo;      // V (MN-major) descriptor byte offsets; 0 = defaults (BNKV*128, 1024)
};

constexpr int ATT_BM = 128;
constexpr int ATT_THREADS = 192;

template <int HD>
struct AttnCfg {
  static constexpr int BNKV = (HD <= 128) ? 128 : 64;
  static constexpr int KV_STAGES = 2;
  static constexpr int Q_BYTES = ATT_BM * HD * 2;
  static constexpr int KV_BYTES = BNKV * HD * 2;       // one K or V tile
  static constexpr int P_BYTES = ATT_BM * BNKV * 2;
  static constexpr int BAR_BYTES = 256;
  static constexpr int SMEM_BYTES = Q_BYTES + 2 * KV_STAGES * KV_BYTES + P_BYTES + BAR_BYTES + 768;
  static constexpr int TMEM_USED = BNKV + HD;
  static constexpr int TMEM_COLS = TMEM_USED <= 256 ? 256 : 512;
};


// ---- per-row softmax pieces (thread = q row).  Mask test: column cc is valid iff (unsigned)(cc - lo) < span.
template <int SCH, bool MASK>
__device__ __forceinline__ float att_row_max(uint32_t tS_row, int lo, unsigned span) {
  float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll 1
  for (int c = 0; c < SCH; ++c) {
    uint32_t sr[32];
    tmem_ld32(tS_row + c * 32, sr);
    tmem_ld_wait();
    const int rel = c * 32 - lo;
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      float v0 = __uint_as_float(sr[i]), v1 = __uint_as_float(sr[i + 1]);
      float v2 = __uint_as_float(sr[i + 2]), v3 = __uint_as_float(sr[i + 3]);
      if (MASK) {
        v0 = static_cast<unsigned>(rel + i) < span ? v0 : -INFINITY;
        v1 = static_cast<unsigned>(rel + i + 1) < span ? v1 : -INFINITY;
        v2 = static_cast<unsigned>(rel + i + 2) < span ? v2 : -INFINITY;
        v3 = static_cast<unsigned>(rel + i + 3) < span ? v3 : -INFINITY;
      }
      m0 = fmaxf(m0, v0);
      m1 = fmaxf(m1, v1);
      m2 = fmaxf(m2, v2);
      m3 = fmaxf(m3, v3);
    }
  }
  return fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
}

// P = exp2(x - m) with x = s*sl2 (or the soft-capped score in log2 units); bf16 P -> 128B-swizzled smem rows
// (the K-major A operand of P.V); returns the row sum of the fp32 probabilities.
template <int SCH, bool MASK, bool CAP>
__device__ __forceinline__ float att_row_exp(uint32_t tS_row, uint8_t* sP, int r, float sl2, float m_safe, float cap_in,
                                             float cap_out, int lo, unsigned span) {
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll 1
  for (int c = 0; c < SCH; ++c) {
    uint32_t sr[32];
    tmem_ld32(tS_row + c * 32, sr);
    tmem_ld_wait();
    const int rel = c * 32 - lo;
    float e[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const float sv = __uint_as_float(sr[i]);
      float x;
      if (CAP) x = fmaf(fast_tanh(sv * cap_in), cap_out, -m_safe);   // cap*log2e*tanh(s*scale/cap) - m
      else x = fmaf(sv, sl2, -m_safe);
      if (MASK) x = static_cast<unsigned>(rel + i) < span ? x : -INFINITY;
      e[i] = fast_exp2(x);
    }
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      s0 += e[i];
      s1 += e[i + 1];
      s2 += e[i + 2];
      s3 += e[i + 3];
    }
    uint8_t* prow = sP + (c >> 1) * (ATT_BM * 128) + r * 128;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      uint4 u;
      u.x = pack_bf16(e[8 * i + 0], e[8 * i + 1]);
      u.y = pack_bf16(e[8 * i + 2], e[8 * i + 3]);
      u.z = pack_bf16(e[8 * i + 4], e[8 * i + 5]);
      u.w = pack_bf16(e[8 * i + 6], e[8 * i + 7]);
      const int chunk16 = (c & 1) * 4 + i;
      *reinterpret_cast<uint4*>(prow + ((chunk16 ^ (r & 7)) << 4)) = u;
    }
  }
  return (s0 + s1) + (s2 + s3);
}

template <int HD>
__global__ void __launch_bounds__(ATT_THREADS, (HD == 64) ? 2 : 1)
attn_prefill_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                    const __grid_constant__ CUtensorMap tmap_v, const AttnParams p) {
  using Cfg = AttnCfg<HD>;
  constexpr int BNKV = Cfg::BNKV;
  constexpr int HB = HD / 64;        // 64-column (128-byte) blocks per head
  constexpr int SCH = BNKV / 32;     // 32-column chunks of S per row
  constexpr int OCH = HD / 32;       // 32-column chunks of O per row
  constexpr float TAU = 8.0f;        // lazy-rescale threshold (log2 units)

  extern __shared__ uint8_t smem_raw[];
  const uint32_t pad = (1024u - (smem_u32(smem_raw) & 1023u)) & 1023u;
  if (pad > 768u) __trap();
  uint8_t* smem = smem_raw + pad;
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + Cfg::Q_BYTES;
  uint8_t* sV = sK + Cfg::KV_STAGES * Cfg::KV_BYTES;
  uint8_t* sP = sV + Cfg::KV_STAGES * Cfg::KV_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + Cfg::P_BYTES);
  uint64_t* bar_q = bars;
  uint64_t* k_full = bars + 1;      // [2]
  uint64_t* v_full = bars + 3;      // [2]
  uint64_t* kv_empty = bars + 5;    // [2]
  uint64_t* bar_s = bars + 7;
  uint64_t* bar_p = bars + 8;
  uint64_t* bar_o = bars + 9;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_qt = (p.S + ATT_BM - 1) / ATT_BM;
  const int qt = n_qt - 1 - static_cast<int>(blockIdx.x);   // heaviest (latest) q tiles first
  const int h = blockIdx.y, b = blockIdx.z;
  const int kvh = h / (p.NH / p.NKV);
  const int q0 = qt * ATT_BM;
  const int t0 = *p.t0_ptr;
  const int T = t0 + p.S;                                   // kv length after this call's append
  // kv range this q tile can see
  int kv_end = T;
  if (p.causal) kv_end = min(T, t0 + q0 + ATT_BM);
  int kv_begin = 0;
  if (p.window > 0) kv_begin = max(0, t0 + q0 - p.window + 1);
  const int jt0 = kv_begin / BNKV;
  const int jt1 = (kv_end + BNKV - 1) / BNKV;               // tiles [jt0, jt1)
  const int n_tiles = jt1 - jt0;

  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_v);
    mbar_init(bar_q, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&v_full[s], 1);
      mbar_init(&kv_empty[s], 1);
    }
    mbar_init(bar_s, 1);
    mbar_init(bar_p, 128);
    mbar_init(bar_o, 1);
    fence_barrier_init();
  }
  if (warp == 5) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tS = tmem_base;            // S accumulator columns [0, BNKV)
  const uint32_t tO = tmem_base + BNKV;     // O accumulator columns [BNKV, BNKV + HD)

  if (warp == 4) {
    // ------------------------------ TMA producer ------------------------------
    if (elect_one()) {
      mbar_expect_tx(bar_q, Cfg::Q_BYTES);
#pragma unroll
      for (int hb = 0; hb < HB; ++hb)
        tma_load_2d(sQ + hb * (ATT_BM * 128), &tmap_q, bar_q, h * HD + hb * 64, b * p.S + q0, kEvictFirst);
      const int kv_row0 = (b * p.NKV + kvh) * p.Tmax;
      int stage = 0;
      uint32_t phase = 0;
      for (int j = jt0; j < jt1; ++j) {
        mbar_wait(&kv_empty[stage], phase ^ 1);
        mbar_expect_tx(&k_full[stage], Cfg::KV_BYTES);
#pragma unroll
        for (int hb = 0; hb < HB; ++hb)
          tma_load_2d(sK + stage * Cfg::KV_BYTES + hb * (BNKV * 128), &tmap_k, &k_full[stage], hb * 64,
                      kv_row0 + j * BNKV, kEvictLast);
        mbar_expect_tx(&v_full[stage], Cfg::KV_BYTES);
#pragma unroll
        for (int hb = 0; hb < HB; ++hb)
          tma_load_2d(sV + stage * Cfg::KV_BYTES + hb * (BNKV * 128), &tmap_v, &v_full[stage], hb * 64,
                      kv_row0 + j * BNKV, kEvictLast);
        if (++stage == Cfg::KV_STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 5) {
    // ------------------------------ MMA issuer ------------------------------
    if (elect_one()) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(ATT_BM, BNKV, 0, 0);   // Q (K-major) x K (K-major)
      constexpr uint32_t idesc_o = umma_idesc_bf16(ATT_BM, HD, 0, 1);     // P (K-major) x V (MN-major)
      mbar_wait(bar_q, 0);
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0; it < n_tiles; ++it) {
        const uint32_t k_addr = smem_u32(sK + stage * Cfg::KV_BYTES);
        const uint32_t v_addr = smem_u32(sV + stage * Cfg::KV_BYTES);
        // S = Q . K^T   (contraction over head_dim: HB swizzle atoms x 4 UMMA_K steps)
        mbar_wait(&k_full[stage], phase);
        tc_fence_after();
#pragma unroll
        for (int hb = 0; hb < HB; ++hb) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t ad = umma_smem_desc(smem_u32(sQ) + hb * (ATT_BM * 128) + k * 32, 16, 1024);
            const uint64_t bd = umma_smem_desc(k_addr + hb * (BNKV * 128) + k * 32, 16, 1024);
            umma_bf16_ss(tS, ad, bd, idesc_s, (hb | k) != 0 ? 1u : 0u);
          }
        }
        umma_commit(bar_s);
        // O += P . V    (contraction over the BNKV kv rows, 16 per UMMA)
        mbar_wait(bar_p, it & 1);
        mbar_wait(&v_full[stage], phase);
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < BNKV / 16; ++kk) {
          const uint64_t ad =
              umma_smem_desc(smem_u32(sP) + (kk >> 2) * (ATT_BM * 128) + (kk & 3) * 32, 16, 1024);
          // V tile: kv rows at 128 B pitch; 64-column head_dim blocks are BNKV*128 B apart (LBO)
          const uint64_t bd = umma_smem_desc(v_addr + kk * (16 * 128), p.v_lbo ? p.v_lbo : BNKV * 128,
                                             p.v_sbo ? p.v_sbo : 1024);
          umma_bf16_ss(tO, ad, bd, idesc_o, (it | kk) != 0 ? 1u : 0u);
        }
        umma_commit(&kv_empty[stage]);
        umma_commit(bar_o);
        if (++stage == Cfg::KV_STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else {
    // ------------------------------ softmax / correction / epilogue ------------------------------
    const int r = warp * 32 + lane;                 // q row inside the tile == TMEM lane
    const uint32_t lane_off = static_cast<uint32_t>(warp * 32) << 16;
    const int qpos = t0 + q0 + r;                   // absolute position of this query
    const float LOG2E = 1.4426950408889634f;
    const float sl2 = p.scale * LOG2E;
    const bool cap = p.softcap > 0.f;
    const float cap_in = cap ? p.scale / p.softcap : 0.f, cap_out = p.softcap * LOG2E;
    float m_used = -INFINITY;                        // running max (log2 domain) the P values are scaled by
    float l = 0.f;                                   // running denominator
    for (int it = 0; it < n_tiles; ++it) {
      const int kv0 = (jt0 + it) * BNKV;
      // does any row of this q tile need a mask on this kv tile?
      const bool need_mask = (p.causal && kv0 + BNKV - 1 > t0 + q0) || (kv0 + BNKV > T) ||
                             (p.window > 0 && kv0 < t0 + q0 + ATT_BM - p.window);
      int lo = 0, hi = BNKV;                        // valid kv columns of this row: [lo, hi)
      if (need_mask) {
        int last = T - 1;
        if (p.causal) last = min(last, qpos);
        hi = min(BNKV, last - kv0 + 1);
        if (p.window > 0) lo = max(0, qpos - p.window + 1 - kv0);
      }
      const unsigned span = hi > lo ? static_cast<unsigned>(hi - lo) : 0u;
      mbar_wait(bar_s, it & 1);
      tc_fence_after();
      // ---- pass 1: row max of the raw scores ----
      const float mx = need_mask ? att_row_max<SCH, true>(tS + lane_off, lo, span)
                                 : att_row_max<SCH, false>(tS + lane_off, lo, span);
      float m_tile = mx;                            // -> log2 domain, after the optional soft-cap
      if (mx != -INFINITY) {
        if (cap) m_tile = cap_out * fast_tanh(mx * cap_in);
        else m_tile = mx * sl2;
      }
      // ---- lazy rescale of O / l ----
      float alpha = 1.0f;
      if (m_tile > m_used + TAU || m_used == -INFINITY) {
        if (m_used != -INFINITY) alpha = fast_exp2(m_used - m_tile);
        if (m_tile != -INFINITY) m_used = m_tile;
      }
      if (it > 0) {
        mbar_wait(bar_o, (it - 1) & 1);             // previous P.V finished: O readable, sP reusable
        tc_fence_after();
        if (__any_sync(0xffffffffu, alpha != 1.0f)) {
          l *= alpha;
#pragma unroll 1
          for (int c = 0; c < OCH; ++c) {
            uint32_t orr[32];
            tmem_ld32(tO + lane_off + c * 32, orr);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) orr[i] = __float_as_uint(__uint_as_float(orr[i]) * alpha);
            tmem_st32(tO + lane_off + c * 32, orr);
          }
          tmem_st_wait();
        }
      }
      // ---- pass 2: P = exp2(s - m), row sum, bf16 P -> swizzled smem (K-major SW128 A operand) ----
      const float m_safe = (m_used == -INFINITY) ? 0.f : m_used;
      float lsum;
      if (cap) {
        lsum = need_mask ? att_row_exp<SCH, true, true>(tS + lane_off, sP, r, sl2, m_safe, cap_in, cap_out, lo, span)
                         : att_row_exp<SCH, false, true>(tS + lane_off, sP, r, sl2, m_safe, cap_in, cap_out, lo, span);
      } else {
        lsum = need_mask ? att_row_exp<SCH, true, false>(tS + lane_off, sP, r, sl2, m_safe, cap_in, cap_out, lo, span)
                         : att_row_exp<SCH, false, false>(tS + lane_off, sP, r, sl2, m_safe, cap_in, cap_out, lo, span);
      }
      l += lsum;
      fence_proxy_async_smem();                      // generic-proxy smem writes -> visible to the MMA
      tc_fence_before();
      mbar_arrive(bar_p);
    }
    // ---- epilogue: O / l -> bf16 -> [B*S, NH*HD] (the reference's transpose+reshape, :509-511) ----
    if (n_tiles > 0) {
      mbar_wait(bar_o, (n_tiles - 1) & 1);
      tc_fence_after();
    }
    const float inv_l = (l > 0.f) ? 1.0f / l : 0.f;
    const bool row_ok = (q0 + r) < p.S;
    __nv_bfloat16* orow = p.out + static_cast<size_t>(b * p.S + q0 + r) * p.ldo + h * HD;
#pragma unroll 1
    for (int c = 0; c < OCH; ++c) {
      __syncwarp();
      uint32_t orr[32];
      if (n_tiles > 0) {
        tmem_ld32(tO + lane_off + c * 32, orr);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) orr[i] = 0u;
      }
      if (row_ok) {
        uint4* d = reinterpret_cast<uint4*>(orow + c * 32);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint4 u;
          u.x = pack_bf16(__uint_as_float(orr[8 * i + 0]) * inv_l, __uint_as_float(orr[8 * i + 1]) * inv_l);
          u.y = pack_bf16(__uint_as_float(orr[8 * i + 2]) * inv_l, __uint_as_float(orr[8 * i + 3]) * inv_l);
          u.z = pack_bf16(__uint_as_float(orr[8 * i + 4]) * inv_l, __uint_as_float(orr[8 * i + 5]) * inv_l);
          u.w = pack_bf16(__uint_as_float(orr[8 * i + 6]) * inv_l, __uint_as_float(orr[8 * i + 7]) * inv_l);
          d[i] = u;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------
// Decode attention (q_len == 1): bandwidth-bound, CUDA cores.  One warp per (batch, q head).
// A kv position is owned by LPP = HD/64 adjacent lanes (64 head-dim elements each, 16-byte loads);
// 32/LPP positions are in flight per warp iteration, each lane group runs its own online softmax,
// and the groups are merged with warp shuffles at the end.  Reference path:
// llama3.2_model_numpy.py:470-511 with q_len = 1 (no mask branch, :479).
// ---------------------------------------------------------------------------------------------
template <int HD>
__global__ void __launch_bounds__(128)
attn_decode_kernel(const __nv_bfloat16* __restrict__ q, int ldq, const __nv_bfloat16* __restrict__ kc,
                   const __nv_bfloat16* __restrict__ vc, __nv_bfloat16* __restrict__ out, int ldo, int B, int NH,
                   int NKV, int Tmax, const int* __restrict__ t0_ptr, float scale, float softcap, int window) {
  constexpr int LPP = HD / 64;                   // lanes per kv position
  constexpr int PPI = 32 / LPP;                  // positions per warp iteration
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gw = blockIdx.x * 4 + warp;          // (b, h)
  if (gw >= B * NH) return;
  const int b = gw / NH, h = gw - b * NH;
  const int kvh = h / (NH / NKV);
  const int T = *t0_ptr + 1;                     // this step's token is already appended
  const int qpos = T - 1;
  const int kv_begin = (window > 0) ? max(0, qpos - window + 1) : 0;
  const int slice = lane % LPP;                  // which 64-element slice of the head this lane owns
  float qf[64];
  {
    const uint4* q4 = reinterpret_cast<const uint4*>(q + static_cast<size_t>(b) * ldq + h * HD + slice * 64);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const uint4 u = __ldg(q4 + i);
      const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = __bfloat1622float2(h2[e]);
        qf[8 * i + 2 * e] = f.x;
        qf[8 * i + 2 * e + 1] = f.y;
      }
    }
  }
  const __nv_bfloat16* kbase = kc + static_cast<size_t>(b * NKV + kvh) * Tmax * HD + slice * 64;
  const __nv_bfloat16* vbase = vc + static_cast<size_t>(b * NKV + kvh) * Tmax * HD + slice * 64;
  const float LOG2E = 1.4426950408889634f;
  float m = -INFINITY, l = 0.f;
  float o[64];
#pragma unroll
  for (int i = 0; i < 64; ++i) o[i] = 0.f;
  for (int tb = kv_begin; tb < T; tb += PPI) {   // warp-uniform trip count
    const int t = tb + lane / LPP;
    const bool ok = t < T;
    float s = 0.f;
    if (ok) {
      const uint4* k4 = reinterpret_cast<const uint4*>(kbase + static_cast<size_t>(t) * HD);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const uint4 u = __ldg(k4 + i);
        const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = __bfloat1622float2(h2[e]);
          s = fmaf(qf[8 * i + 2 * e], f.x, s);
          s = fmaf(qf[8 * i + 2 * e + 1], f.y, s);
        }
      }
    }
#pragma unroll
    for (int off = 1; off < LPP; off <<= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if (ok) {
      s *= scale;
      if (softcap > 0.f) s = softcap * tanhf(s / softcap);
      s *= LOG2E;
      const float m_new = fmaxf(m, s);
      const float alpha = fast_exp2(m - m_new);  // 0 on the first item (m = -inf)
      const float pexp = fast_exp2(s - m_new);
      l = l * alpha + pexp;
      const uint4* v4 = reinterpret_cast<const uint4*>(vbase + static_cast<size_t>(t) * HD);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const uint4 u = __ldg(v4 + i);
        const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = __bfloat1622float2(h2[e]);
          o[8 * i + 2 * e] = fmaf(pexp, f.x, o[8 * i + 2 * e] * alpha);
          o[8 * i + 2 * e + 1] = fmaf(pexp, f.y, o[8 * i + 2 * e + 1] * alpha);
        }
      }
      m = m_new;
    }
  }
  // merge the PPI per-group partial softmaxes (lanes with equal `slice`)
  float m_all = m;
#pragma unroll
  for (int off = 16; off >= LPP; off >>= 1) m_all = fmaxf(m_all, __shfl_xor_sync(0xffffffffu, m_all, off));
  const float w = (m == -INFINITY) ? 0.f : fast_exp2(m - m_all);
  float l_all = l * w;
#pragma unroll
  for (int off = 16; off >= LPP; off >>= 1) l_all += __shfl_xor_sync(0xffffffffu, l_all, off);
  const float inv = l_all > 0.f ? 1.0f / l_all : 0.f;
#pragma unroll
  for (int i = 0; i < 64; ++i) {
    float v = o[i] * w;
#pragma unroll
    for (int off = 16; off >= LPP; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    o[i] = v * inv;
  }
  if (lane < LPP) {
    uint4* d = reinterpret_cast<uint4*>(out + static_cast<size_t>(b) * ldo + h * HD + slice * 64);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      uint4 u;
      u.x = pack_bf16(o[8 * i + 0], o[8 * i + 1]);
      u.y = pack_bf16(o[8 * i + 2], o[8 * i + 3]);
      u.z = pack_bf16(o[8 * i + 4], o[8 * i + 5]);
      u.w = pack_bf16(o[8 * i + 6], o[8 * i + 7]);
      d[i] = u;
    }
  }
}

}  // namespace b2
