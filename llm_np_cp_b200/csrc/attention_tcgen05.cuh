// Grouped-query attention on sm_100a (tcgen05 + TMEM + TMA for prefill, mma.sync + TMA for decode).
//
// Replaces, per layer, the reference's materialised chain
//   repeat_kv -> Q.K^T / sqrt(HD) -> where(tril) -> softmax -> P.V -> transpose/reshape
// (llama3.2_model_numpy.py:470-511; gemma2_model.py:536-577) with one kernel that never writes the
// [B,NH,S,T] score tensor.  Options the reference lacks but BASELINE.json asks for are in-kernel
// predicates/epilogues: sliding window and tanh soft-cap (semantics of transformers' modeling_gemma2).
// GQA: kv head = q head / n_rep, no K/V replication (replaces repeat_kv_np, :188-204).
// Online softmax in the exp2 domain with lazy rescaling of the TMEM-resident O accumulator.
//
// Kernels in this file:
//   attn_prefill_co_kernel<HD>       prefill, head_dim 64 / 128: one 128-row q tile per CTA, TWO CTAs per SM   (default)
//   attn_prefill_kernel<HD, NP>      prefill, one CTA per SM: head_dim 256, precise (hi/lo planes) mode, S <= 128
//       CTA = one (batch, q-head, 128-row q tile); 192 threads:
//         warps 0-3  softmax/correction/epilogue: thread = q row (= TMEM lane)
//         warp 4     TMA producer (Q once, K/V tiles through a 2-stage ring)
//         warp 5     tcgen05.mma issuer:  S = Q.K^T (both K-major)  ->  TMEM[0, BNKV)
//                                         O += P.V  (P K-major from smem, V MN-major) -> TMEM[BNKV, BNKV+HD)
//   attn_decode_split_kernel<HD>     decode (q_len 1): split-KV, all query heads of a kv head per CTA, mma.sync   (default)
//   attn_decode_kernel<HD>           decode, one warp per (batch, q head) on the FMA pipe: precise mode, odd head ratios
#pragma once
#include "common.cuh"

namespace b2 {

struct AttnParams {
  int B, S, NH, NKV, Tmax;
  KvPages pages;         // block table of a paged cache (table == nullptr: dense [B][NKV][Tmax][HD] planes)
  const int* t0_ptr;     // tokens already cached before this call (device scalar)
  float scale;           // 1/sqrt(head_dim) (reference) or query_pre_attn_scalar^-0.5 (HF flag)
  float softcap;         // 0 = off
  int causal;            // reference quirk: 0 when q_len <= 2 (llama3.2_model_numpy.py:479)
  int window;            // 0 = off; else keep kv > qpos - window
  __nv_bfloat16* out;    // [B*S, NH*HD]
  int ldo;
  int v_lbo, v_sbo;      // V (MN-major) descriptor byte offsets; 0 = defaults (BNKV*128, 1024)
  // precise (hi/lo split-bf16) mode: every bf16 operand x is carried as x_hi + x_lo
  int q_lo_col;          // column offset of the lo plane inside a q row (= NH*HD), 0 = not precise
  int kv_lo_row;         // row offset of the lo plane inside the K/V tensor maps (= rows of one plane)
  int kv_plane_rows;     // rows of one K/V plane (0 = B*NKV*Tmax, the dense layout of the single-kernel test entry points)
  int out_lo_col;        // column offset of the lo plane inside an output row
  unsigned long long* trace;   // bring-up: clock64() timeline of CTA (0,0,0); nullptr = off
  unsigned long long* cta_trace;   // diagnostics: per CTA {globaltimer at start, at end, smid}; nullptr = off
};

// timeline tracing (diagnostics only): role r of CTA (0,0,0) appends (event id, clock) pairs to trace[r*1024 ...]
struct AttTrace {
  unsigned long long* buf;
  int n;
  __device__ __forceinline__ void init(unsigned long long* t, int role) {
    buf = (t && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) ? t + role * 1024 : nullptr;
    n = 0;
  }
  __device__ __forceinline__ void ev(int id) {
    if (buf && n < 510) {
      buf[2 * n] = static_cast<unsigned long long>(id);
      buf[2 * n + 1] = static_cast<unsigned long long>(clock64());
      ++n;
    }
  }
};

#ifndef ATT_POLY
#define ATT_POLY 0
#endif
#ifndef ATT_POLY_MOD
#define ATT_POLY_MOD 4
#endif
constexpr int ATT_BM = 128;
constexpr int ATT_THREADS = 192;

// NP = operand planes: 1 = plain bf16, 2 = precise mode (hi and lo planes of Q, K, V and P; three MMAs per product:
// hi.hi + hi.lo + lo.hi)
template <int HD, int NP = 1>
struct AttnCfg {
  // precise mode at head_dim 256 (Gemma-2): Q hi+lo alone is 128 KB, so kv tiles are 32 wide and K / V single-buffered
  static constexpr int BNKV = (NP == 2) ? (HD <= 128 ? 64 : 32) : ((HD <= 128) ? 128 : 64);
  static constexpr int KV_STAGES = (NP == 2 && HD > 128) ? 1 : 2;
  static constexpr int Q_BYTES = NP * ATT_BM * HD * 2;
  static constexpr int KV_BYTES = NP * BNKV * HD * 2;  // one K or V tile (all planes)
  static constexpr int P_PLANE = ATT_BM * 128 * ((BNKV + 63) / 64);   // P rows are 128-byte swizzled rows (64 columns) even when
                                                                      // only 32 of them are used
  static constexpr int P_BYTES = NP * P_PLANE;
  static constexpr int BAR_BYTES = 256;
  static constexpr int SMEM_BYTES = Q_BYTES + 2 * KV_STAGES * KV_BYTES + P_BYTES + BAR_BYTES + 768;
  static_assert(SMEM_BYTES <= 232448, "shared memory budget");
  static constexpr int TMEM_USED = BNKV + HD;
  static constexpr int TMEM_COLS = TMEM_USED <= 256 ? 256 : 512;
};


// ---- per-row softmax pieces (thread = q row).  Mask test: column cc is valid iff (unsigned)(cc - lo) < span.
// TMEM loads are software-pipelined: the load of chunk c+1 is in flight while chunk c is processed
// (tcgen05.wait::ld waits for every outstanding load, so it is placed after the compute of chunk c).
template <bool MASK>
__device__ __forceinline__ void att_max32(const uint32_t (&sr)[32], int rel, unsigned span, float (&m)[4]) {
  // FMNMX3 (three-input max): 64 instead of 128 max instructions per 128-score row
#pragma unroll
  for (int i = 0; i < 32; i += 8) {
    float v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      v[k] = __uint_as_float(sr[i + k]);
      if (MASK) v[k] = static_cast<unsigned>(rel + i + k) < span ? v[k] : -INFINITY;
    }
    m[0] = fmax3(m[0], v[0], v[1]);
    m[1] = fmax3(m[1], v[2], v[3]);
    m[2] = fmax3(m[2], v[4], v[5]);
    m[3] = fmax3(m[3], v[6], v[7]);
  }
}

// one 32-column chunk: P = exp2(x - m) with x = s*sl2 (or the soft-capped score in log2 units);
// bf16 P -> 128B-swizzled smem row (the K-major A operand of P.V); accumulates the fp32 row sum.
template <bool MASK, bool CAP>
__device__ __forceinline__ void att_exp32(const uint32_t (&sr)[32], int c, uint8_t* sP, int r, float sl2, float m_safe,
                                          float cap_in, float cap_out, int lo, unsigned span, float (&sum)[4],
                                          uint8_t* sPlo = nullptr) {
  const int rel = c * 32 - lo;
  float e[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    const float sv = __uint_as_float(sr[i]);
    float x;
    if (CAP) x = fmaf(fast_tanh(sv * cap_in), cap_out, -m_safe);   // cap*log2e*tanh(s*scale/cap) - m
    else x = fmaf(sv, sl2, -m_safe);
    if (MASK) x = static_cast<unsigned>(rel + i) < span ? x : -INFINITY;
    // MUFU.EX2 is the bottleneck of this loop (quarter-rate pipe): ATT_POLY of every ATT_POLY_MOD exponentials run on the
    // FMA pipe instead (compile-time pattern, so both paths are straight-line code)
    e[i] = ((i % ATT_POLY_MOD) < ATT_POLY) ? poly_exp2(x) : fast_exp2(x);
  }
#pragma unroll
  for (int i = 0; i < 32; i += 4) {
    sum[0] += e[i];
    sum[1] += e[i + 1];
    sum[2] += e[i + 2];
    sum[3] += e[i + 3];
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
  if (sPlo) {   // precise mode: P_lo = bf16(P - bf16(P))
    uint8_t* lrow = sPlo + (c >> 1) * (ATT_BM * 128) + r * 128;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float d[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) d[k] = e[8 * i + k] - __bfloat162float(__float2bfloat16_rn(e[8 * i + k]));
      uint4 u;
      u.x = pack_bf16(d[0], d[1]);
      u.y = pack_bf16(d[2], d[3]);
      u.z = pack_bf16(d[4], d[5]);
      u.w = pack_bf16(d[6], d[7]);
      const int chunk16 = (c & 1) * 4 + i;
      *reinterpret_cast<uint4*>(lrow + ((chunk16 ^ (r & 7)) << 4)) = u;
    }
  }
}

// One kv tile of the online softmax for one q row, SINGLE pass over TMEM: the whole S row (SCH*32 fp32) is read
// into registers once (TMEM read bandwidth, not MUFU, is the first limit of a two-pass scheme), S is handed
// back to the tensor core immediately (bar_sfree), then max -> lazy O rescale -> exp2 / row sum / bf16 P.
// S lives at tS_row, O at tO_row (TMEM, this thread's lane), P goes to sP (smem, 128B-swizzled rows).
template <int SCH, bool MASK, bool CAP>
__device__ __forceinline__ float att_exp_regs(uint32_t (&x)[SCH][32], uint8_t* sP, int r, float sl2, float m_safe,
                                              float cap_in, float cap_out, int lo, unsigned span, uint8_t* sPlo) {
  float sum[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int c = 0; c < SCH; ++c)
    att_exp32<MASK, CAP>(x[c], c, sP, r, sl2, m_safe, cap_in, cap_out, lo, span, sum, sPlo);
  return (sum[0] + sum[1]) + (sum[2] + sum[3]);
}

template <int SCH, int OCH>
__device__ __forceinline__ void att_softmax_tile(uint32_t tS_row, uint32_t tO_row, uint8_t* sP, int r, bool need_mask,
                                                 int lo, unsigned span, bool cap, float sl2, float cap_in, float cap_out,
                                                 float& m_used, float& l, bool have_o, uint64_t* bar_o, uint32_t o_parity,
                                                 uint64_t* bar_sfree, uint8_t* sPlo = nullptr, AttTrace* tr = nullptr) {
  constexpr float TAU = 8.0f;   // lazy-rescale threshold (log2 units): P stays <= 2^8
  uint32_t x[SCH][32];
#pragma unroll
  for (int c = 0; c < SCH; ++c) tmem_ld32(tS_row + c * 32, x[c]);
  tmem_ld_wait();
  if (bar_sfree) {              // scores are in registers: the tensor core may overwrite S
    tc_fence_before();
    mbar_arrive(bar_sfree);
  }
  if (tr) tr->ev(32);
  float m[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
  if (need_mask) {
#pragma unroll
    for (int c = 0; c < SCH; ++c) att_max32<true>(x[c], c * 32 - lo, span, m);
  } else {
#pragma unroll
    for (int c = 0; c < SCH; ++c) att_max32<false>(x[c], 0, 0u, m);
  }
  const float mx = fmaxf(fmaxf(m[0], m[1]), fmaxf(m[2], m[3]));
  float m_tile = mx;                               // -> log2 domain, after the optional soft-cap
  if (mx != -INFINITY) m_tile = cap ? cap_out * fast_tanh(mx * cap_in) : mx * sl2;
  float alpha = 1.0f;
  if (m_tile > m_used + TAU || m_used == -INFINITY) {
    if (m_used != -INFINITY) alpha = fast_exp2(m_used - m_tile);
    if (m_tile != -INFINITY) m_used = m_tile;
  }
  if (tr) tr->ev(33);
  if (have_o) {
    mbar_wait(bar_o, o_parity);                    // previous P.V finished: O readable, sP reusable
    tc_fence_after();
    if (tr) tr->ev(34);
    if (__any_sync(0xffffffffu, alpha != 1.0f)) {
      l *= alpha;
#pragma unroll 1
      for (int c = 0; c < OCH; ++c) {
        uint32_t orr[32];
        tmem_ld32(tO_row + c * 32, orr);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) orr[i] = __float_as_uint(__uint_as_float(orr[i]) * alpha);
        tmem_st32(tO_row + c * 32, orr);
      }
      tmem_st_wait();
    }
  }
  const float m_safe = (m_used == -INFINITY) ? 0.f : m_used;
  float lsum;
  if (cap) {
    lsum = need_mask ? att_exp_regs<SCH, true, true>(x, sP, r, sl2, m_safe, cap_in, cap_out, lo, span, sPlo)
                     : att_exp_regs<SCH, false, true>(x, sP, r, sl2, m_safe, cap_in, cap_out, lo, span, sPlo);
  } else {
    lsum = need_mask ? att_exp_regs<SCH, true, false>(x, sP, r, sl2, m_safe, cap_in, cap_out, lo, span, sPlo)
                     : att_exp_regs<SCH, false, false>(x, sP, r, sl2, m_safe, cap_in, cap_out, lo, span, sPlo);
  }
  l += lsum;
  if (tr) tr->ev(37);
}

// O / l -> bf16 -> out row (the reference's transpose+reshape, llama3.2_model_numpy.py:509-511)
template <int OCH>
__device__ __forceinline__ void att_store_out(uint32_t tO_row, bool have_o, float l, bool row_ok, __nv_bfloat16* orow,
                                              int lo_col = 0) {
  const float inv_l = (l > 0.f) ? 1.0f / l : 0.f;
#pragma unroll 1
  for (int c = 0; c < OCH; ++c) {
    __syncwarp();
    uint32_t orr[32];
    if (have_o) {
      tmem_ld32(tO_row + c * 32, orr);
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
      if (lo_col) {   // precise mode: residual plane
        uint4* dl = reinterpret_cast<uint4*>(orow + lo_col + c * 32);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float v[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const float f = __uint_as_float(orr[8 * i + k]) * inv_l;
            v[k] = f - __bfloat162float(__float2bfloat16_rn(f));
          }
          uint4 u;
          u.x = pack_bf16(v[0], v[1]);
          u.y = pack_bf16(v[2], v[3]);
          u.z = pack_bf16(v[4], v[5]);
          u.w = pack_bf16(v[6], v[7]);
          dl[i] = u;
        }
      }
    }
  }
}

template <int HD, int NP>
__global__ void __launch_bounds__(ATT_THREADS, 1)
attn_prefill_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                    const __grid_constant__ CUtensorMap tmap_v, const AttnParams p) {
  using Cfg = AttnCfg<HD, NP>;
  constexpr int QP = ATT_BM * HD * 2;                  // bytes of one Q plane
  constexpr int KP = Cfg::BNKV * HD * 2;               // bytes of one K / V plane of a stage
  constexpr int PP = Cfg::P_PLANE;                     // bytes of one P plane
  constexpr int BNKV = Cfg::BNKV;
  constexpr int HB = HD / 64;        // 64-column (128-byte) blocks per head
  constexpr int SCH = BNKV / 32;     // 32-column chunks of S per row
  constexpr int OCH = HD / 32;       // 32-column chunks of O per row

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
  pdl_trigger();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();                               // Q/K/V and the cache length come from the preceding kernels
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
  const uint32_t tS = tmem_base;            // S accumulator columns [0, BNKV)
  const uint32_t tO = tmem_base + BNKV;     // O accumulator columns [BNKV, BNKV + HD)

  if (warp == 4) {
    // ------------------------------ TMA producer ------------------------------
    if (elect_one()) {
      mbar_expect_tx(bar_q, Cfg::Q_BYTES);
#pragma unroll
      for (int pl = 0; pl < NP; ++pl)
#pragma unroll
        for (int hb = 0; hb < HB; ++hb)
          tma_load_2d(sQ + pl * QP + hb * (ATT_BM * 128), &tmap_q, bar_q, pl * p.q_lo_col + h * HD + hb * 64, b * p.S + q0,
                      kEvictFirst);
      int stage = 0;
      uint32_t phase = 0;
      for (int j = jt0; j < jt1; ++j) {
        const int trow = static_cast<int>(kv_row(p.pages, b, kvh, p.NKV, p.Tmax, j * BNKV));   // first row of kv tile j
        mbar_wait(&kv_empty[stage], phase ^ 1);
        mbar_expect_tx(&k_full[stage], Cfg::KV_BYTES);
#pragma unroll
        for (int pl = 0; pl < NP; ++pl)
#pragma unroll
          for (int hb = 0; hb < HB; ++hb)
            tma_load_2d(sK + stage * Cfg::KV_BYTES + pl * KP + hb * (BNKV * 128), &tmap_k, &k_full[stage], hb * 64,
                        pl * p.kv_lo_row + trow, kEvictLast);
        mbar_expect_tx(&v_full[stage], Cfg::KV_BYTES);
#pragma unroll
        for (int pl = 0; pl < NP; ++pl)
#pragma unroll
          for (int hb = 0; hb < HB; ++hb)
            tma_load_2d(sV + stage * Cfg::KV_BYTES + pl * KP + hb * (BNKV * 128), &tmap_v, &v_full[stage], hb * 64,
                        pl * p.kv_lo_row + trow, kEvictLast);
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
        // S = Q . K^T   (contraction over head_dim: HB swizzle atoms x 4 UMMA_K steps); precise mode sums the three
        // cross terms Qh.Kh + Qh.Kl + Ql.Kh into the same accumulator
        mbar_wait(&k_full[stage], phase);
        tc_fence_after();
#pragma unroll
        for (int term = 0; term < (NP == 2 ? 3 : 1); ++term) {
          const int qa = (term == 2) ? 1 : 0, kb = (term == 1) ? 1 : 0;   // plane of Q / of K
#pragma unroll
          for (int hb = 0; hb < HB; ++hb) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t ad = umma_smem_desc(smem_u32(sQ) + qa * QP + hb * (ATT_BM * 128) + k * 32, 16, 1024);
              const uint64_t bd = umma_smem_desc(k_addr + kb * KP + hb * (BNKV * 128) + k * 32, 16, 1024);
              umma_bf16_ss(tS, ad, bd, idesc_s, (term | hb | k) != 0 ? 1u : 0u);
            }
          }
        }
        umma_commit(bar_s);
        // O += P . V    (contraction over the BNKV kv rows, 16 per UMMA); precise: Ph.Vh + Ph.Vl + Pl.Vh
        mbar_wait(bar_p, it & 1);
        mbar_wait(&v_full[stage], phase);
        tc_fence_after();
#pragma unroll
        for (int term = 0; term < (NP == 2 ? 3 : 1); ++term) {
          const int pa = (term == 2) ? 1 : 0, vb = (term == 1) ? 1 : 0;   // plane of P / of V
#pragma unroll
          for (int kk = 0; kk < BNKV / 16; ++kk) {
            const uint64_t ad =
                umma_smem_desc(smem_u32(sP) + pa * PP + (kk >> 2) * (ATT_BM * 128) + (kk & 3) * 32, 16, 1024);
            // V tile: kv rows at 128 B pitch; 64-column head_dim blocks are BNKV*128 B apart (LBO)
            const uint64_t bd = umma_smem_desc(v_addr + vb * KP + kk * (16 * 128), p.v_lbo ? p.v_lbo : BNKV * 128,
                                               p.v_sbo ? p.v_sbo : 1024);
            umma_bf16_ss(tO, ad, bd, idesc_o, (it | term | kk) != 0 ? 1u : 0u);
          }
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
      att_softmax_tile<SCH, OCH>(tS + lane_off, tO + lane_off, sP, r, need_mask, lo, span, cap, sl2, cap_in, cap_out, m_used, l,
                                 it > 0, bar_o, (it - 1) & 1, nullptr, NP == 2 ? sP + PP : nullptr);
      fence_proxy_async_smem();                      // generic-proxy smem writes -> visible to the MMA
      tc_fence_before();
      mbar_arrive(bar_p);
    }
    // ---- epilogue: O / l -> bf16 -> [B*S, NH*HD] ----
    if (n_tiles > 0) {
      mbar_wait(bar_o, (n_tiles - 1) & 1);
      tc_fence_after();
    }
    att_store_out<OCH>(tO + lane_off, n_tiles > 0, l, (q0 + r) < p.S,
                       p.out + static_cast<size_t>(b * p.S + q0 + r) * p.ldo + h * HD, NP == 2 ? p.out_lo_col : 0);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------
// Co-resident prefill kernel (head_dim 64 / 128): ONE 128-row q tile per CTA, TWO CTAs per SM.
// The exponentials (MUFU.EX2, 8 cycles per warp instruction per SM sub-partition) bound attention at these head
// sizes.  Two independent CTAs on an SM drift out of phase on their own, so one CTA's exp phase overlaps the
// other's TMEM loads / row max / barrier waits / prologue / epilogue -- the two-q-tiles-in-one-CTA kernel above keeps
// its tiles in lockstep through the single in-order MMA issuer and leaves the XU pipe idle ~45% of the time.
// 256 threads = 2 warpgroups (setmaxnreg is per warpgroup):
//   warps 0-3   softmax / correction / epilogue (thread = q row = TMEM lane; the whole S row lives in registers)
//   warp  4     TMA producer: Q once; K through a 2-slot ring, V through a 1-slot ring, issued in the order
//               K0 K1 V0 | K(j+2) V(j+1) ... so a load never queues behind one that waits for a later event
//   warp  5     tcgen05.mma issuer:  S(j+1) = Q.K(j+1)^T as soon as S(j) is in registers, then O += P(j).V(j)
//   warps 6-7   idle (complete the second warpgroup)
// smem per CTA (head_dim 64): Q 16K + K 2x16K + V 16K + P 32K = 96 KB  -> 2 CTAs / SM;  TMEM 256 columns per CTA.
// ---------------------------------------------------------------------------------------------
template <int HD>
struct AttnCoCfg {
  static constexpr int BNKV = (HD == 64) ? 128 : 64;
  static constexpr int K_STAGES = 2;
  static constexpr int Q_BYTES = ATT_BM * HD * 2;
  static constexpr int KV_BYTES = BNKV * HD * 2;
  static constexpr int P_BYTES = ATT_BM * BNKV * 2;
  static constexpr int BAR_BYTES = 256;
  static constexpr int SMEM_BYTES = Q_BYTES + (K_STAGES + 1) * KV_BYTES + P_BYTES + BAR_BYTES + 1024;
  static constexpr int TMEM_COLS = 256;
  static_assert(BNKV + HD <= 256, "TMEM budget");
  static_assert(2 * (SMEM_BYTES + 1024) <= 233472, "two CTAs per SM");
};
constexpr int ATT_CO_THREADS = 256;

template <int HD>
__global__ void __launch_bounds__(ATT_CO_THREADS, 2)
attn_prefill_co_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                       const __grid_constant__ CUtensorMap tmap_v, const AttnParams p) {
  using Cfg = AttnCoCfg<HD>;
  constexpr int BNKV = Cfg::BNKV;
  constexpr int HB = HD / 64;
  constexpr int SCH = BNKV / 32;
  constexpr int OCH = HD / 32;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + Cfg::Q_BYTES;                      // [2]
  uint8_t* sV = sK + Cfg::K_STAGES * Cfg::KV_BYTES;     // [1]
  uint8_t* sP = sV + Cfg::KV_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + Cfg::P_BYTES);
  uint64_t* bar_q = bars;
  uint64_t* k_full = bars + 1;       // [2]
  uint64_t* k_empty = bars + 3;      // [2]
  uint64_t* v_full = bars + 5;
  uint64_t* v_empty = bars + 6;
  uint64_t* s_full = bars + 7;       // S written by the tensor core
  uint64_t* s_free = bars + 8;       // S copied to registers by all 128 softmax threads
  uint64_t* p_ready = bars + 9;      // P written to smem by all 128 softmax threads
  uint64_t* o_done = bars + 10;      // O += P.V finished
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 11);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned long long cta_t0 = 0;
  if (p.cta_trace && threadIdx.x == 0) cta_t0 = globaltimer_ns();
  const int n_qt = (p.S + ATT_BM - 1) / ATT_BM;
  const int qt = n_qt - 1 - static_cast<int>(blockIdx.x);   // heaviest (latest) q tiles first
  const int h = blockIdx.y, b = blockIdx.z;
  const int kvh = h / (p.NH / p.NKV);
  const int q0 = qt * ATT_BM;

  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_v);
    mbar_init(bar_q, 1);
    for (int s = 0; s < Cfg::K_STAGES; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], 1);
    }
    mbar_init(v_full, 1);
    mbar_init(v_empty, 1);
    mbar_init(s_full, 1);
    mbar_init(s_free, 128);
    mbar_init(p_ready, 128);
    mbar_init(o_done, 1);
    fence_barrier_init();
  }
  if (warp == 5) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  pdl_trigger();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();                               // Q/K/V and the cache length come from the preceding kernels
  const int t0 = *p.t0_ptr;
  const int T = t0 + p.S;
  int kv_end = T;
  if (p.causal) kv_end = min(T, t0 + q0 + ATT_BM);
  int kv_begin = 0;
  if (p.window > 0) kv_begin = max(0, t0 + q0 - p.window + 1);
  const int jt0 = kv_begin / BNKV;
  const int jt1 = (kv_end + BNKV - 1) / BNKV;               // kv tiles [jt0, jt1)
  const int n_tiles = jt1 - jt0;
  const uint32_t tS = tmem_base;            // S accumulator columns [0, BNKV)
  const uint32_t tO = tmem_base + BNKV;     // O accumulator columns [BNKV, BNKV + HD)

  if (warp >= 4) {
    setmaxnreg_dec<40>();
    if (warp == 4) {
      // ------------------------------ TMA producer ------------------------------
      if (elect_one()) {
        mbar_expect_tx(bar_q, Cfg::Q_BYTES);
#pragma unroll
        for (int hb = 0; hb < HB; ++hb)
          tma_load_2d(sQ + hb * (ATT_BM * 128), &tmap_q, bar_q, h * HD + hb * 64, b * p.S + q0, kEvictFirst);
        auto tile_row = [&](int it) { return static_cast<int>(kv_row(p.pages, b, kvh, p.NKV, p.Tmax, (jt0 + it) * BNKV)); };
        auto load_k = [&](int it) {               // it = tile index relative to jt0
          const int st = it & 1;
          const int trow = tile_row(it);
          mbar_wait(&k_empty[st], ((it >> 1) & 1) ^ 1);
          mbar_expect_tx(&k_full[st], Cfg::KV_BYTES);
#pragma unroll
          for (int hb = 0; hb < HB; ++hb)
            tma_load_2d(sK + st * Cfg::KV_BYTES + hb * (BNKV * 128), &tmap_k, &k_full[st], hb * 64, trow, kEvictLast);
        };
        auto load_v = [&](int it) {
          const int trow = tile_row(it);
          mbar_wait(v_empty, (it & 1) ^ 1);
          mbar_expect_tx(v_full, Cfg::KV_BYTES);
#pragma unroll
          for (int hb = 0; hb < HB; ++hb)
            tma_load_2d(sV + hb * (BNKV * 128), &tmap_v, v_full, hb * 64, trow, kEvictLast);
        };
        if (n_tiles > 0) load_k(0);
        if (n_tiles > 1) load_k(1);
        if (n_tiles > 0) load_v(0);
        for (int it = 0; it < n_tiles; ++it) {
          if (it + 2 < n_tiles) load_k(it + 2);   // slot freed by S(it)
          if (it + 1 < n_tiles) load_v(it + 1);   // slot freed by P.V(it)
        }
      }
    } else if (warp == 5) {
      // ------------------------------ MMA issuer ------------------------------
      if (elect_one()) {
        constexpr uint32_t idesc_s = umma_idesc_bf16(ATT_BM, BNKV, 0, 0);   // Q (K-major) x K (K-major)
        constexpr uint32_t idesc_o = umma_idesc_bf16(ATT_BM, HD, 0, 1);     // P (K-major) x V (MN-major)
        auto issue_s = [&](int it) {
          const int st = it & 1;
          mbar_wait(&k_full[st], (it >> 1) & 1);
          tc_fence_after();
          const uint32_t k_addr = smem_u32(sK + st * Cfg::KV_BYTES);
#pragma unroll
          for (int hb = 0; hb < HB; ++hb)
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16_ss(tS, umma_smem_desc(smem_u32(sQ) + hb * (ATT_BM * 128) + k * 32, 16, 1024),
                           umma_smem_desc(k_addr + hb * (BNKV * 128) + k * 32, 16, 1024), idesc_s, (hb | k) != 0 ? 1u : 0u);
          umma_commit(&k_empty[st]);
          umma_commit(s_full);
        };
        AttTrace tr;
        tr.init(p.trace, 0);
        tr.ev(0);
        mbar_wait(bar_q, 0);
        tr.ev(1);
        if (n_tiles > 0) issue_s(0);
        for (int it = 0; it < n_tiles; ++it) {
          if (it + 1 < n_tiles) {
            mbar_wait(s_free, it & 1);              // S(it) is in registers
            tr.ev(10);
            issue_s(it + 1);
            tr.ev(12);
          }
          mbar_wait(p_ready, it & 1);
          tr.ev(20);
          mbar_wait(v_full, it & 1);
          tr.ev(21);
          tc_fence_after();
          const uint32_t v_addr = smem_u32(sV);
#pragma unroll
          for (int kk = 0; kk < BNKV / 16; ++kk)
            umma_bf16_ss(tO, umma_smem_desc(smem_u32(sP) + (kk >> 2) * (ATT_BM * 128) + (kk & 3) * 32, 16, 1024),
                         umma_smem_desc(v_addr + kk * (16 * 128), BNKV * 128, 1024), idesc_o, (it | kk) != 0 ? 1u : 0u);
          umma_commit(v_empty);
          umma_commit(o_done);
          tr.ev(22);
        }
      }
    }
  } else {
    // ------------------------------ softmax / correction / epilogue ------------------------------
    setmaxnreg_inc<216>();
    const int r = warp * 32 + lane;                 // q row inside the tile == TMEM lane
    const uint32_t lane_off = static_cast<uint32_t>(warp * 32) << 16;
    const int qpos = t0 + q0 + r;
    const float LOG2E = 1.4426950408889634f;
    const float sl2 = p.scale * LOG2E;
    const bool cap = p.softcap > 0.f;
    const float cap_in = cap ? p.scale / p.softcap : 0.f, cap_out = p.softcap * LOG2E;
    float m_used = -INFINITY, l = 0.f;
    AttTrace tr;
    tr.init((warp == 0 && lane == 0) ? p.trace : nullptr, 1);
    for (int it = 0; it < n_tiles; ++it) {
      const int kv0 = (jt0 + it) * BNKV;
      const bool need_mask = (p.causal && kv0 + BNKV - 1 > t0 + q0) || (kv0 + BNKV > T) ||
                             (p.window > 0 && kv0 < t0 + q0 + ATT_BM - p.window);
      int lo = 0, hi = BNKV;
      if (need_mask) {
        int last = T - 1;
        if (p.causal) last = min(last, qpos);
        hi = min(BNKV, last - kv0 + 1);
        if (p.window > 0) lo = max(0, qpos - p.window + 1 - kv0);
      }
      const unsigned span = hi > lo ? static_cast<unsigned>(hi - lo) : 0u;
      tr.ev(30);
      mbar_wait(s_full, it & 1);
      tr.ev(31);
      tc_fence_after();
      att_softmax_tile<SCH, OCH>(tS + lane_off, tO + lane_off, sP, r, need_mask, lo, span, cap, sl2, cap_in, cap_out, m_used, l,
                                 it > 0, o_done, (it - 1) & 1, s_free, nullptr, &tr);
      fence_proxy_async_smem();
      tr.ev(38);
      tc_fence_before();
      mbar_arrive(p_ready);
      tr.ev(35);
    }
    if (n_tiles > 0) {
      mbar_wait(o_done, (n_tiles - 1) & 1);
      tc_fence_after();
    }
    att_store_out<OCH>(tO + lane_off, n_tiles > 0, l, (q0 + r) < p.S,
                       p.out + static_cast<size_t>(b * p.S + q0 + r) * p.ldo + h * HD);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
  if (p.cta_trace && threadIdx.x == 0) {
    unsigned long long* t = p.cta_trace + 4 * ((static_cast<size_t>(blockIdx.z) * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x);
    t[0] = cta_t0;
    t[1] = globaltimer_ns();
    t[2] = smid();
    t[3] = static_cast<unsigned long long>(clock64());
  }
}

// ---------------------------------------------------------------------------------------------
// Persistent variant of the co-resident kernel: grid = 2 CTAs per SM, each CTA walks a static, heaviest-first list of
// (q tile, head, batch) work items -- item t, t + grid, t + 2 grid, ... with items ordered by their kv length (longest
// first, LPT) and the q heads of one kv head adjacent (their K/V tiles hit in L2).
// What it removes: the fixed cost per 128-row tile of the one-tile-per-CTA kernel (measured 5.4 us of prologue -- launch,
// barrier init, TMEM alloc, Q + first K load, first S -- and epilogue against ~1.9 us per kv iteration): barriers and
// TMEM live for the whole kernel, the producer runs ahead across items (the next item's Q is requested as soon as the
// last S = Q.K^T of the current one has retired, its K tiles as the ring slots free up), the MMA thread issues the next
// item's first S while the softmax warps are still in the last kv tile / epilogue of the current one, and the epilogue
// hands O back (o_free) after its TMEM loads, before the global stores.  All barrier parities come from running
// counters (items, kv iterations, K / V tiles) that the three roles advance in lock step.
// ---------------------------------------------------------------------------------------------
template <int HD>
__global__ void __launch_bounds__(ATT_CO_THREADS, 2)
attn_prefill_persist_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                            const __grid_constant__ CUtensorMap tmap_v, const AttnParams p, const int total_items) {
  using Cfg = AttnCoCfg<HD>;
  constexpr int BNKV = Cfg::BNKV;
  constexpr int HB = HD / 64;
  constexpr int SCH = BNKV / 32;
  constexpr int OCH = HD / 32;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + Cfg::Q_BYTES;                      // [2]
  uint8_t* sV = sK + Cfg::K_STAGES * Cfg::KV_BYTES;     // [1]
  uint8_t* sP = sV + Cfg::KV_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + Cfg::P_BYTES);
  uint64_t* q_full = bars;
  uint64_t* q_empty = bars + 1;      // every S = Q.K^T of the item has retired: sQ may be overwritten
  uint64_t* k_full = bars + 2;       // [2]
  uint64_t* k_empty = bars + 4;      // [2]
  uint64_t* v_full = bars + 6;
  uint64_t* v_empty = bars + 7;
  uint64_t* s_full = bars + 8;       // S written by the tensor core
  uint64_t* s_free = bars + 9;       // S copied to registers by all 128 softmax threads
  uint64_t* p_ready = bars + 10;     // P written to smem by all 128 softmax threads
  uint64_t* o_done = bars + 11;      // O += P.V finished
  uint64_t* o_free = bars + 12;      // the epilogue has read O out of TMEM
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_v);
    mbar_init(q_full, 1);
    mbar_init(q_empty, 1);
    for (int s = 0; s < Cfg::K_STAGES; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], 1);
    }
    mbar_init(v_full, 1);
    mbar_init(v_empty, 1);
    mbar_init(s_full, 1);
    mbar_init(s_free, 128);
    mbar_init(p_ready, 128);
    mbar_init(o_done, 1);
    mbar_init(o_free, 128);
    fence_barrier_init();
  }
  if (warp == 5) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  pdl_trigger();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();                               // Q/K/V and the cache length come from the preceding kernels
  const int t0 = *p.t0_ptr;
  const int T = t0 + p.S;
  const int n_qt = (p.S + ATT_BM - 1) / ATT_BM;
  const int BH = p.B * p.NH;
  const uint32_t tS = tmem_base;            // S accumulator columns [0, BNKV)
  const uint32_t tO = tmem_base + BNKV;     // O accumulator columns [BNKV, BNKV + HD)

  // work item -> (q tile, head, batch) and its kv tile range; items are ordered heaviest (latest q tile) first
  struct Item { int q0, h, b, jt0, n; };
  auto item = [&](int t) {
    Item w;
    const int qt = n_qt - 1 - t / BH, bh = t - (t / BH) * BH;
    w.h = bh % p.NH;
    w.b = bh / p.NH;
    w.q0 = qt * ATT_BM;
    int kv_end = T;
    if (p.causal) kv_end = min(T, t0 + w.q0 + ATT_BM);
    int kv_begin = 0;
    if (p.window > 0) kv_begin = max(0, t0 + w.q0 - p.window + 1);
    w.jt0 = kv_begin / BNKV;
    w.n = (kv_end + BNKV - 1) / BNKV - w.jt0;
    return w;
  };
  const int first = blockIdx.x, stride = gridDim.x;

  if (warp >= 4) {
    setmaxnreg_dec<40>();
    if (warp == 4) {
      // ------------------------------ TMA producer ------------------------------
      if (elect_one()) {
        uint32_t kc = 0, vc = 0, ic = 0;        // K tiles, V tiles, items loaded so far
        for (int t = first; t < total_items; t += stride, ++ic) {
          const Item w = item(t);
          const int kvh = w.h / (p.NH / p.NKV);
          mbar_wait(q_empty, (ic & 1) ^ 1);     // previous item's S products are done with sQ (fresh barrier: passes)
          mbar_expect_tx(q_full, Cfg::Q_BYTES);
#pragma unroll
          for (int hb = 0; hb < HB; ++hb)
            tma_load_2d(sQ + hb * (ATT_BM * 128), &tmap_q, q_full, w.h * HD + hb * 64, w.b * p.S + w.q0, kEvictFirst);
          auto tile_row = [&](int it) { return static_cast<int>(kv_row(p.pages, w.b, kvh, p.NKV, p.Tmax, (w.jt0 + it) * BNKV)); };
          auto load_k = [&](int it) {
            const int st = kc & 1;
            const int trow = tile_row(it);          // (block-table lookup for a paged cache) before the slot wait
            mbar_wait(&k_empty[st], ((kc >> 1) & 1) ^ 1);
            mbar_expect_tx(&k_full[st], Cfg::KV_BYTES);
#pragma unroll
            for (int hb = 0; hb < HB; ++hb)
              tma_load_2d(sK + st * Cfg::KV_BYTES + hb * (BNKV * 128), &tmap_k, &k_full[st], hb * 64, trow, kEvictLast);
            ++kc;
          };
          auto load_v = [&](int it) {
            const int trow = tile_row(it);
            mbar_wait(v_empty, (vc & 1) ^ 1);
            mbar_expect_tx(v_full, Cfg::KV_BYTES);
#pragma unroll
            for (int hb = 0; hb < HB; ++hb)
              tma_load_2d(sV + hb * (BNKV * 128), &tmap_v, v_full, hb * 64, trow, kEvictLast);
            ++vc;
          };
          if (w.n > 0) load_k(0);
          if (w.n > 1) load_k(1);
          if (w.n > 0) load_v(0);
          for (int it = 0; it < w.n; ++it) {
            if (it + 2 < w.n) load_k(it + 2);   // slot freed by S(it)
            if (it + 1 < w.n) load_v(it + 1);   // slot freed by P.V(it)
          }
        }
      }
    } else if (warp == 5) {
      // ------------------------------ MMA issuer ------------------------------
      if (elect_one()) {
        constexpr uint32_t idesc_s = umma_idesc_bf16(ATT_BM, BNKV, 0, 0);   // Q (K-major) x K (K-major)
        constexpr uint32_t idesc_o = umma_idesc_bf16(ATT_BM, HD, 0, 1);     // P (K-major) x V (MN-major)
        uint32_t kc = 0, vc = 0, gi = 0, ic = 0;   // K tiles consumed, V tiles consumed, kv iterations, items
        auto issue_s = [&](bool last_of_item) {    // S = Q . K(kc)^T -> TMEM; k_full of the slot must be complete
          const int st = kc & 1;
          mbar_wait(&k_full[st], (kc >> 1) & 1);
          tc_fence_after();
          const uint32_t k_addr = smem_u32(sK + st * Cfg::KV_BYTES);
#pragma unroll
          for (int hb = 0; hb < HB; ++hb)
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16_ss(tS, umma_smem_desc(smem_u32(sQ) + hb * (ATT_BM * 128) + k * 32, 16, 1024),
                           umma_smem_desc(k_addr + hb * (BNKV * 128) + k * 32, 16, 1024), idesc_s, (hb | k) != 0 ? 1u : 0u);
          umma_commit(&k_empty[st]);
          if (last_of_item) umma_commit(q_empty);
          umma_commit(s_full);
          ++kc;
        };
        bool s0_issued = false;                    // the current item's S(0) was issued ahead, during the previous item
        for (int t = first; t < total_items; t += stride, ++ic) {
          const Item w = item(t);                  // w.n >= 1 always: the diagonal kv tile exists for every q tile
          const int tn = t + stride;
          const bool has_next = tn < total_items;
          if (!s0_issued) {
            mbar_wait(q_full, ic & 1);
            issue_s(w.n == 1);
          }
          s0_issued = false;
          for (int it = 0; it < w.n; ++it) {
            const bool last = it + 1 == w.n;
            if (!last) {
              mbar_wait(s_free, gi & 1);           // S(it) is in registers
              issue_s(it + 2 == w.n);
            }
            if (last && has_next) {
              // overlap: while the softmax warps work on the last kv tile, issue the next item's S(0) as soon as S is
              // free and its Q / K tiles have landed -- but never delay this item's last P.V for them
              for (;;) {
                if (!s0_issued && mbar_test_wait(s_free, gi & 1) && mbar_test_wait(q_full, (ic + 1) & 1) &&
                    mbar_test_wait(&k_full[kc & 1], (kc >> 1) & 1)) {
                  issue_s(item(tn).n == 1);
                  s0_issued = true;
                }
                if (mbar_test_wait(p_ready, gi & 1)) break;
                __nanosleep(40);
              }
            } else {
              mbar_wait(p_ready, gi & 1);
            }
            mbar_wait(v_full, vc & 1);
            if (it == 0) mbar_wait(o_free, (ic & 1) ^ 1);   // previous item's epilogue has read O (fresh barrier: passes)
            tc_fence_after();
            const uint32_t v_addr = smem_u32(sV);
#pragma unroll
            for (int kk = 0; kk < BNKV / 16; ++kk)
              umma_bf16_ss(tO, umma_smem_desc(smem_u32(sP) + (kk >> 2) * (ATT_BM * 128) + (kk & 3) * 32, 16, 1024),
                           umma_smem_desc(v_addr + kk * (16 * 128), BNKV * 128, 1024), idesc_o, (it | kk) != 0 ? 1u : 0u);
            umma_commit(v_empty);
            umma_commit(o_done);
            ++vc;
            ++gi;
          }
        }
      }
    }
  } else {
    // ------------------------------ softmax / correction / epilogue ------------------------------
    setmaxnreg_inc<216>();
    const int r = warp * 32 + lane;                 // q row inside the tile == TMEM lane
    const uint32_t lane_off = static_cast<uint32_t>(warp * 32) << 16;
    const float LOG2E = 1.4426950408889634f;
    const float sl2 = p.scale * LOG2E;
    const bool cap = p.softcap > 0.f;
    const float cap_in = cap ? p.scale / p.softcap : 0.f, cap_out = p.softcap * LOG2E;
    uint32_t gi = 0;
    for (int t = first; t < total_items; t += stride) {
      const Item w = item(t);
      const int qpos = t0 + w.q0 + r;
      float m_used = -INFINITY, l = 0.f;
      for (int it = 0; it < w.n; ++it) {
        const int kv0 = (w.jt0 + it) * BNKV;
        const bool need_mask = (p.causal && kv0 + BNKV - 1 > t0 + w.q0) || (kv0 + BNKV > T) ||
                               (p.window > 0 && kv0 < t0 + w.q0 + ATT_BM - p.window);
        int lo = 0, hi = BNKV;
        if (need_mask) {
          int last = T - 1;
          if (p.causal) last = min(last, qpos);
          hi = min(BNKV, last - kv0 + 1);
          if (p.window > 0) lo = max(0, qpos - p.window + 1 - kv0);
        }
        const unsigned span = hi > lo ? static_cast<unsigned>(hi - lo) : 0u;
        mbar_wait(s_full, gi & 1);
        tc_fence_after();
        att_softmax_tile<SCH, OCH>(tS + lane_off, tO + lane_off, sP, r, need_mask, lo, span, cap, sl2, cap_in, cap_out, m_used, l,
                                   it > 0, o_done, (gi - 1) & 1, s_free);
        fence_proxy_async_smem();
        tc_fence_before();
        mbar_arrive(p_ready);
        ++gi;
      }
      // epilogue: O (TMEM) -> registers, hand O back to the tensor core, then scale / convert / store
      mbar_wait(o_done, (gi - 1) & 1);
      tc_fence_after();
      uint32_t o[OCH][32];
#pragma unroll
      for (int c = 0; c < OCH; ++c) tmem_ld32(tO + lane_off + c * 32, o[c]);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(o_free);
      if (w.q0 + r < p.S) {
        const float inv_l = (l > 0.f) ? 1.0f / l : 0.f;
        __nv_bfloat16* orow = p.out + static_cast<size_t>(w.b * p.S + w.q0 + r) * p.ldo + w.h * HD;
#pragma unroll
        for (int c = 0; c < OCH; ++c) {
          uint4* d = reinterpret_cast<uint4*>(orow + c * 32);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            uint4 u;
            u.x = pack_bf16(__uint_as_float(o[c][8 * i + 0]) * inv_l, __uint_as_float(o[c][8 * i + 1]) * inv_l);
            u.y = pack_bf16(__uint_as_float(o[c][8 * i + 2]) * inv_l, __uint_as_float(o[c][8 * i + 3]) * inv_l);
            u.z = pack_bf16(__uint_as_float(o[c][8 * i + 4]) * inv_l, __uint_as_float(o[c][8 * i + 5]) * inv_l);
            u.w = pack_bf16(__uint_as_float(o[c][8 * i + 6]) * inv_l, __uint_as_float(o[c][8 * i + 7]) * inv_l);
            d[i] = u;
          }
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
// Decode attention (q_len == 1): bandwidth/latency-bound, CUDA cores.  One warp per (batch, q head).
// A kv position is owned by LPP = HD/16 adjacent lanes (16 head-dim elements = two 16-byte loads each), so
// 32/LPP positions are in flight per warp iteration and a K/V row is read as one contiguous 128..512-byte
// segment by neighbouring lanes.  Each lane group runs its own online softmax; the groups are merged with
// warp shuffles at the end.  Reference path: llama3.2_model_numpy.py:470-511 with q_len = 1 (no mask, :479).
// ---------------------------------------------------------------------------------------------
template <int HD>
__global__ void __launch_bounds__(128)
attn_decode_kernel(const __nv_bfloat16* __restrict__ q, int ldq, const __nv_bfloat16* __restrict__ kc,
                   const __nv_bfloat16* __restrict__ vc, __nv_bfloat16* __restrict__ out, int ldo, int B, int NH,
                   int NKV, int Tmax, const int* __restrict__ t0_ptr, float scale, float softcap, int window,
                   int q_lo_col, size_t kv_lo_off, int out_lo_col,     // precise mode: lo planes (0 = plain bf16)
                   const KvPages pages) {
  constexpr int LPP = HD / 16;                   // lanes per kv position (4 / 8 / 16)
  constexpr int PPI = 32 / LPP;                  // positions per warp iteration
  pdl_trigger();
  pdl_wait();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gw = blockIdx.x * 4 + warp;          // (b, h)
  if (gw >= B * NH) return;
  const int b = gw / NH, h = gw - b * NH;
  const int kvh = h / (NH / NKV);
  const int T = *t0_ptr + 1;                     // this step's token is already appended
  const int qpos = T - 1;
  const int kv_begin = (window > 0) ? max(0, qpos - window + 1) : 0;
  const int slice = lane % LPP;                  // which 16-element slice of the head this lane owns
  const int grp = lane / LPP;                    // which of the PPI positions of an iteration
  float qf[16];
  {
    const uint4* q4 = reinterpret_cast<const uint4*>(q + static_cast<size_t>(b) * ldq + h * HD + slice * 16);
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const uint4 u = __ldg(q4 + i);
      const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = __bfloat1622float2(h2[e]);
        qf[8 * i + 2 * e] = f.x;
        qf[8 * i + 2 * e + 1] = f.y;
      }
      if (q_lo_col) {
        const uint4 ul = __ldg(q4 + i + q_lo_col / 8);
        const __nv_bfloat162* l2 = reinterpret_cast<const __nv_bfloat162*>(&ul);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = __bfloat1622float2(l2[e]);
          qf[8 * i + 2 * e] += f.x;
          qf[8 * i + 2 * e + 1] += f.y;
        }
      }
    }
  }
  const __nv_bfloat16* kbase = kc + slice * 16;
  const __nv_bfloat16* vbase = vc + slice * 16;
  const float LOG2E = 1.4426950408889634f;
  float m = -INFINITY, l = 0.f;
  float o[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) o[i] = 0.f;
  for (int tb = kv_begin; tb < T; tb += PPI) {   // warp-uniform trip count
    const int t = tb + grp;
    const bool ok = t < T;
    float s = 0.f;
    uint4 kv[2], vv[2], vlo[2] = {};
    if (ok) {
      const size_t trow = kv_row(pages, b, kvh, NKV, Tmax, t);
      const uint4* k4 = reinterpret_cast<const uint4*>(kbase + trow * HD);
      const uint4* v4 = reinterpret_cast<const uint4*>(vbase + trow * HD);
      kv[0] = __ldg(k4);
      kv[1] = __ldg(k4 + 1);
      vv[0] = __ldg(v4);          // V issued together with K: one memory round trip per iteration
      vv[1] = __ldg(v4 + 1);
      uint4 kl[2] = {}, vl[2] = {};
      if (kv_lo_off) {
        kl[0] = __ldg(k4 + kv_lo_off / 8);
        kl[1] = __ldg(k4 + kv_lo_off / 8 + 1);
        vl[0] = __ldg(v4 + kv_lo_off / 8);
        vl[1] = __ldg(v4 + kv_lo_off / 8 + 1);
      }
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&kv[i]);
        const __nv_bfloat162* l2 = reinterpret_cast<const __nv_bfloat162*>(&kl[i]);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float2 f = __bfloat1622float2(h2[e]);
          if (kv_lo_off) {
            const float2 g = __bfloat1622float2(l2[e]);
            f.x += g.x;
            f.y += g.y;
          }
          s = fmaf(qf[8 * i + 2 * e], f.x, s);
          s = fmaf(qf[8 * i + 2 * e + 1], f.y, s);
        }
      }
      if (kv_lo_off) {   // fold V_lo into the registers holding V (as fp32 pairs below)
#pragma unroll
        for (int i = 0; i < 2; ++i) vlo[i] = vl[i];
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
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&vv[i]);
        const __nv_bfloat162* l2 = reinterpret_cast<const __nv_bfloat162*>(&vlo[i]);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float2 f = __bfloat1622float2(h2[e]);
          if (kv_lo_off) {
            const float2 g = __bfloat1622float2(l2[e]);
            f.x += g.x;
            f.y += g.y;
          }
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
  for (int i = 0; i < 16; ++i) {
    float v = o[i] * w;
#pragma unroll
    for (int off = 16; off >= LPP; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    o[i] = v * inv;
  }
  if (lane < LPP) {
    uint4* d = reinterpret_cast<uint4*>(out + static_cast<size_t>(b) * ldo + h * HD + slice * 16);
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      uint4 u;
      u.x = pack_bf16(o[8 * i + 0], o[8 * i + 1]);
      u.y = pack_bf16(o[8 * i + 2], o[8 * i + 3]);
      u.z = pack_bf16(o[8 * i + 4], o[8 * i + 5]);
      u.w = pack_bf16(o[8 * i + 6], o[8 * i + 7]);
      d[i] = u;
      if (out_lo_col) {
        float v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = o[8 * i + k] - __bfloat162float(__float2bfloat16_rn(o[8 * i + k]));
        uint4 ul;
        ul.x = pack_bf16(v[0], v[1]);
        ul.y = pack_bf16(v[2], v[3]);
        ul.z = pack_bf16(v[4], v[5]);
        ul.w = pack_bf16(v[6], v[7]);
        d[i + out_lo_col / 8] = ul;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Split-KV GQA decode attention (q_len == 1, long caches).  The per-(batch, q head) kernel above walks the cache
// serially with one memory round trip per 8 positions and reads K/V once per Q head; at T ~ 2K that is ~6x off the
// HBM roofline.  Here a CTA owns (T-slice, kv head, batch) and serves ALL n_rep = NH/NKV query heads of the kv head
// from one pass over K/V (the GQA reuse the reference materialises with repeat_kv, llama3.2_model_numpy.py:188-204,
// :470-471):
//   warp 4      producer: 64-position K and V tiles through a shared-memory ring (TMA, 128B swizzle)
//   warps 0-3   consumers: warp w owns positions [16w, 16w+16) of every tile.  The n_rep query heads are the rows of
//               a 16-row MMA tile: S = Q.K^T and O += P.V run on the tensor cores with mma.sync.m16n8k16 (ldmatrix
//               fragments straight from the swizzled tiles) -- ~4 instructions per position instead of ~35 on the
//               FMA pipe -- with a per-row online softmax on the accumulator fragments.
// Partial (m, l, o[HD]) per (head, slice) go to a scratch buffer; the LAST CTA to finish a (batch, kv head) -- found
// with a counter -- merges the slices in fixed order (deterministic) and writes bf16.  Grid = (slices, NKV, B).
// ---------------------------------------------------------------------------------------------
template <int HD>
struct DecSplitCfg {
  static constexpr int TP = 64;                  // positions per tile
  static constexpr int STAGES = (HD == 64) ? 4 : (HD == 128 ? 3 : 2);
  static constexpr int HB = HD / 64;             // 64-column (128-byte) blocks per row
  static constexpr int TILE_BYTES = TP * HD * 2; // one K or V tile
  static constexpr int SMEM_BYTES = 2 * STAGES * TILE_BYTES + 128 + 1024;
};
constexpr int DEC_SPLIT_THREADS = 160;
constexpr int DEC_MAX_SPLITS = 16;
constexpr int DEC_MAX_NREP = 8;                  // query heads per kv head served by rows 0..7 of the MMA tile

template <int HD>
__global__ void __launch_bounds__(DEC_SPLIT_THREADS)
attn_decode_split_kernel(const __grid_constant__ CUtensorMap tmap_k, const __grid_constant__ CUtensorMap tmap_v,
                         const __nv_bfloat16* __restrict__ q, int ldq, __nv_bfloat16* __restrict__ out, int ldo, int NKV,
                         int nrep, int Tmax, const int* __restrict__ t0_ptr, float scale, float softcap, int window,
                         float* __restrict__ part, int* __restrict__ counters, const KvPages pages) {
  using Cfg = DecSplitCfg<HD>;
  constexpr int TP = Cfg::TP, STAGES = Cfg::STAGES, HB = Cfg::HB, TILE_BYTES = Cfg::TILE_BYTES;
  constexpr int KS = HD / 16;                    // k-steps of S = Q.K^T
  constexpr int NT = HD / 8;                     // 8-column n-tiles of O
  constexpr int PW = HD + 2;                     // floats per partial: o[HD], m, l
  static_assert(4 * DEC_MAX_NREP * PW * 4 <= 2 * STAGES * TILE_BYTES, "merge area aliases the K/V ring");
  extern __shared__ uint8_t dsm_raw[];
  uint8_t* dsm = dsm_raw + ((1024u - (smem_u32(dsm_raw) & 1023u)) & 1023u);
  uint8_t* sK = dsm;
  uint8_t* sV = dsm + STAGES * TILE_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(dsm + 2 * STAGES * TILE_BYTES);
  uint64_t* empty = full + STAGES;
  int* s_flag = reinterpret_cast<int*>(empty + STAGES);
  float* s_merge = reinterpret_cast<float*>(dsm);          // [4 warps][DEC_MAX_NREP][PW], valid after the main loop only

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nsplit = gridDim.x, sp = blockIdx.x, kvh = blockIdx.y, b = blockIdx.z;
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_v);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 4);
    }
    fence_barrier_init();
  }
  pdl_trigger();
  __syncthreads();
  pdl_wait();
  const int T = *t0_ptr + 1;                     // this step's token is already appended
  const int qpos = T - 1;
  const int kv_begin = (window > 0) ? max(0, qpos - window + 1) : 0;
  const int t_base = kv_begin & ~(TP - 1);       // tiles start on multiples of TP: they never straddle a cache page
  const int len = T - t_base;
  const int per = (((len + nsplit - 1) / nsplit) + TP - 1) / TP * TP;     // positions per slice, whole tiles
  const int t_lo = t_base + sp * per;
  const int t_hi = min(T, t_lo + per);
  const int n_tiles = t_hi > t_lo ? (t_hi - t_lo + TP - 1) / TP : 0;

  const int g = lane >> 2, tq = lane & 3;        // MMA fragment coordinates: row g (= query head), column pair tq
  float m_row = -INFINITY, l_row = 0.f;          // row g's running max (log2 units) / this thread's share of the row sum
  float o[NT][4];
  if (warp == 4) {
    if (lane == 0) {
      for (int i = 0; i < n_tiles; ++i) {
        const int st = i % STAGES;
        mbar_wait(&empty[st], ((i / STAGES) & 1) ^ 1);
        mbar_expect_tx(&full[st], 2 * TILE_BYTES);
        const int trow = static_cast<int>(kv_row(pages, b, kvh, NKV, Tmax, t_lo + i * TP));
#pragma unroll
        for (int hb = 0; hb < HB; ++hb) {
          tma_load_2d(sK + st * TILE_BYTES + hb * (TP * 128), &tmap_k, &full[st], hb * 64, trow, kEvictFirst);
          tma_load_2d(sV + st * TILE_BYTES + hb * (TP * 128), &tmap_v, &full[st], hb * 64, trow, kEvictFirst);
        }
      }
    }
  } else {
    // A fragments of Q: row g = head kvh*nrep + g (zero rows beyond nrep), k = head dims
    uint32_t qa[KS][2];
    {
      const __nv_bfloat16* qrow = q + static_cast<size_t>(b) * ldq + (kvh * nrep + g) * HD;
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        qa[ks][0] = (g < nrep) ? __ldg(reinterpret_cast<const uint32_t*>(qrow + ks * 16 + 2 * tq)) : 0u;
        qa[ks][1] = (g < nrep) ? __ldg(reinterpret_cast<const uint32_t*>(qrow + ks * 16 + 8 + 2 * tq)) : 0u;
      }
    }
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) o[nt][0] = o[nt][1] = o[nt][2] = o[nt][3] = 0.f;
    const float LOG2E = 1.4426950408889634f;
    const float sl2 = scale * LOG2E;
    const int p0 = warp * 16;                    // this warp's positions inside a tile
    // ldmatrix row addresses (swizzled): K fragments: lane -> (row p0 + nt*8 + lane%8, chunk 4*half + lane/8)
    const int lrow = lane & 7, lsel = lane >> 3;
    for (int i = 0; i < n_tiles; ++i) {
      const int st = i % STAGES;
      mbar_wait(&full[st], (i / STAGES) & 1);
      const int tpos = t_lo + i * TP + p0;       // absolute position of this warp's first row
      if (tpos < t_hi && tpos + 16 > kv_begin) {   // this warp's 16 positions intersect [kv_begin, t_hi)
        const uint32_t kb = smem_u32(sK + st * TILE_BYTES), vb = smem_u32(sV + st * TILE_BYTES);
        float sc[2][4];
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
          sc[nt][0] = sc[nt][1] = sc[nt][2] = sc[nt][3] = 0.f;
          const int r = p0 + nt * 8 + lrow;
#pragma unroll
          for (int hb = 0; hb < HB; ++hb)
#pragma unroll
            for (int half = 0; half < 2; ++half) {
              uint32_t b0, b1, b2, b3;
              ldmatrix_x4(kb + hb * (TP * 128) + r * 128 + (((4 * half + lsel) ^ (r & 7)) << 4), b0, b1, b2, b3);
              const int ks = hb * 4 + half * 2;
              mma_bf16_16816(sc[nt], qa[ks][0], 0u, qa[ks][1], 0u, b0, b1);
              mma_bf16_16816(sc[nt], qa[ks + 1][0], 0u, qa[ks + 1][1], 0u, b2, b3);
            }
        }
        // row g's four scores: positions tpos + nt*8 + 2*tq + {0, 1}
        float x[4];
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            float v = sc[nt][e];
            if (softcap > 0.f) v = softcap * LOG2E * tanhf(v * scale / softcap);
            else v *= sl2;
            const int pos = tpos + nt * 8 + 2 * tq + e;
            x[nt * 2 + e] = (pos < t_hi && pos >= kv_begin) ? v : -INFINITY;
          }
        float mx = fmaxf(fmaxf(x[0], x[1]), fmaxf(x[2], x[3]));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
        const float m_new = fmaxf(m_row, mx);    // finite: at least one of the 16 positions is valid, for every row alike
        const float alpha = fast_exp2(m_row - m_new);
        float pe[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) pe[e] = fast_exp2(x[e] - m_new);
        l_row = l_row * alpha + (pe[0] + pe[1]) + (pe[2] + pe[3]);
        m_row = m_new;
        if (alpha != 1.0f) {
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) o[nt][0] *= alpha, o[nt][1] *= alpha;
        }
        // P as the A fragment of O += P.V (k = this warp's 16 positions); rows 8..15 are unused -> 0
        const uint32_t pa0 = pack_bf16(pe[0], pe[1]), pa2 = pack_bf16(pe[2], pe[3]);
#pragma unroll
        for (int np = 0; np < NT / 2; ++np) {    // two 8-column n-tiles (16 head dims) per ldmatrix.x4.trans
          const int chunk = (2 * np) & 7, hb = (2 * np) >> 3;
          const int r = p0 + lrow + 8 * (lsel & 1);
          uint32_t b0, b1, b2, b3;
          ldmatrix_x4_trans(vb + hb * (TP * 128) + r * 128 + (((chunk + (lsel >> 1)) ^ (r & 7)) << 4), b0, b1, b2, b3);
          mma_bf16_16816(o[2 * np], pa0, 0u, pa2, 0u, b0, b1);
          mma_bf16_16816(o[2 * np + 1], pa0, 0u, pa2, 0u, b2, b3);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[st]);
    }
    l_row += __shfl_xor_sync(0xffffffffu, l_row, 1);
    l_row += __shfl_xor_sync(0xffffffffu, l_row, 2);
  }
  __syncthreads();                               // every warp is done with the K/V ring: reuse it for the merge
  if (warp < 4 && g < nrep) {
    float* d = s_merge + (warp * DEC_MAX_NREP + g) * PW;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      d[nt * 8 + 2 * tq] = o[nt][0];
      d[nt * 8 + 2 * tq + 1] = o[nt][1];
    }
    if (tq == 0) d[HD] = m_row, d[HD + 1] = l_row;
  }
  __syncthreads();
  const size_t pbase = (static_cast<size_t>(b * NKV + kvh) * nrep) * nsplit;
  for (int idx = threadIdx.x; idx < nrep * HD; idx += DEC_SPLIT_THREADS) {
    const int j = idx / HD, d = idx - j * HD;
    float m_all = -INFINITY;
#pragma unroll
    for (int w = 0; w < 4; ++w) m_all = fmaxf(m_all, s_merge[(w * DEC_MAX_NREP + j) * PW + HD]);
    float acc = 0.f, ls = 0.f;
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      const float mw = s_merge[(w * DEC_MAX_NREP + j) * PW + HD];
      const float wgt = (mw == -INFINITY) ? 0.f : fast_exp2(mw - m_all);
      acc = fmaf(wgt, s_merge[(w * DEC_MAX_NREP + j) * PW + d], acc);
      ls = fmaf(wgt, s_merge[(w * DEC_MAX_NREP + j) * PW + HD + 1], ls);
    }
    if (nsplit == 1) {   // short cache, one slice per (batch, kv head): this CTA holds the whole row -- no partials, no counter
      out[static_cast<size_t>(b) * ldo + (kvh * nrep + j) * HD + d] = __float2bfloat16_rn(ls > 0.f ? acc / ls : 0.f);
      continue;
    }
    float* pr = part + ((pbase + static_cast<size_t>(j) * nsplit) + sp) * PW;
    pr[d] = acc;
    if (d == 0) pr[HD] = m_all, pr[HD + 1] = ls;
  }
  if (nsplit == 1) return;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const int prev = atomicAdd(&counters[b * NKV + kvh], 1);
    *s_flag = (prev == nsplit - 1);
  }
  __syncthreads();
  if (*s_flag) {                                 // last slice of this (batch, kv head): merge all slices, fixed order
    __threadfence();
    for (int idx = threadIdx.x; idx < nrep * HD; idx += DEC_SPLIT_THREADS) {
      const int j = idx / HD, d = idx - j * HD;
      const float* pj = part + (pbase + static_cast<size_t>(j) * nsplit) * PW;
      float m_all = -INFINITY;
      for (int s = 0; s < nsplit; ++s) m_all = fmaxf(m_all, __ldcg(pj + s * PW + HD));
      float acc = 0.f, ls = 0.f;
      for (int s = 0; s < nsplit; ++s) {
        const float ms = __ldcg(pj + s * PW + HD);
        const float wgt = (ms == -INFINITY) ? 0.f : fast_exp2(ms - m_all);
        acc = fmaf(wgt, __ldcg(pj + s * PW + d), acc);
        ls = fmaf(wgt, __ldcg(pj + s * PW + HD + 1), ls);
      }
      out[static_cast<size_t>(b) * ldo + (kvh * nrep + j) * HD + d] = __float2bfloat16_rn(ls > 0.f ? acc / ls : 0.f);
    }
    if (threadIdx.x == 0) counters[b * NKV + kvh] = 0;     // self-cleaning: ready for the next launch
  }
}

}  // namespace b2
