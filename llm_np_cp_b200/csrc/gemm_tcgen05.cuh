// Persistent warp-specialised bf16 GEMM for sm_100a:  C[M,N] = A[M,K] . W[N,K]^T
//
//   A  activations, bf16 row-major (K contiguous)         -> UMMA A operand, K-major
//   W  weights in HF layout [out_features, in_features]   -> UMMA B operand, K-major
//   (this is the reference's Linear_np: y = x . W^T, llama3.2_model_numpy.py:124-144)
//
// One CTA per SM, 320 threads:
//   warp 0    TMA producer: cp.async.bulk.tensor 128B-swizzled {64 x 128} A and {64 x BN} W boxes
//             into a STAGES-deep shared-memory ring, completion on "full" mbarriers
//   warp 1    MMA issuer: one elected thread issues tcgen05.mma (M=128, N=BN, K=16) into one of two
//             TMEM accumulators; tcgen05.commit releases ring slots / publishes the accumulator
//   warps 2-9 epilogue (two warps per TMEM lane quarter, each taking every other column group):
//             tcgen05.ld (lane = output row) -> fused epilogue -> global memory, overlapped with the
//             next tile's main loop through the double-buffered accumulator.  fp32 outputs are staged
//             in 128B-swizzled shared memory and leave through TMA (bulk tensor store, or bulk tensor
//             reduce-add for the residual stream), so HBM sees full 128-byte lines and the residual
//             add needs no load.
//
// Fused epilogues: plain bf16 / fp32 (+tanh soft-cap, gemma2_model.py:867-870),
// residual add (llama3.2_model_numpy.py:569,578), RoPE + KV-cache append for the fused QKV weight
// (:459, :464-465), and act(gate)*up for the interleaved gate/up weight (:176-182).
#pragma once
#include "common.cuh"

namespace b2 {

enum GemmEpi : int { EPI_BF16 = 0, EPI_F32 = 1, EPI_RESID = 2, EPI_QKV_ROPE = 3, EPI_GLU = 4, EPI_RESID_NORM = 5 };

struct GemmParams {
  int M, N, K;   // K = contraction length seen by A (2*k_wrap in hi/lo "precise" mode)
  int k_wrap;    // W k-coordinate = k % k_wrap
  int band_m;    // rasterisation: m-blocks per band (operand tiles stay hot in L2 inside a band)
  int split_k;   // >1: K is split across CTAs (skinny GEMMs): EPI_RESID reduce-adds, EPI_F32 stores split s at rows
                 // [s*split_rows, s*split_rows + M) of `out` (deterministic: the consumer sums the partials in order)
  int split_rows;
  // EPI_BF16 / EPI_GLU (direct stores); EPI_F32 / EPI_RESID go through the output tensor map
  void* out;
  int ldo;         // row stride of out, in elements
  float softcap;   // EPI_F32: cap * tanh(x / cap) when > 0
  int act;         // EPI_GLU: 0 = SiLU, 1 = tanh-GELU
  int lo_off;      // EPI_GLU/EPI_BF16 precise mode: column offset of the bf16 "lo" plane (0 = none)
  // EPI_QKV_ROPE
  int q_cols, kv_cols, head_dim, S, Tmax, nkv;
  __nv_bfloat16* q_out;
  int ldq;
  KvPages pages;           // block table of a paged cache (table == nullptr: dense)
  __nv_bfloat16* k_cache;  // [B][NKV][Tmax][HD] (this layer), or the layer's paged plane [n_pages][NKV][page][HD]
  __nv_bfloat16* v_cache;
  const int* t0_ptr;       // device scalar: tokens already in the cache
  const int* pos_ids;      // optional [S] caller position_ids for RoPE (llama3.2_model_numpy.py:657); nullptr = t0 + s.
                           // The cache slot is always t0 + s.
  const float* rope_cos;   // [max_pos][HD/2]
  const float* rope_sin;
  int kv_lo_off;           // precise mode: element offset of the lo plane of the caches (0 = none)
  int stream_out;          // EPI_F32 whose output is much larger than L2 (lm_head logits): A loads evict-last, W loads and C stores
                           // evict-first, so the write-once C stream and the read-once W stream do not displace A
  int a_box_rows;          // 1-CTA kernel: rows of A one TMA box brings in (32/64 when M is that small, else 0 = 128);
                           // the remaining rows of the 128-row MMA tile are zeroed once instead of fetched every k-step
  // ---- RMSNorm split across two GEMMs (decode step; LlamaRMSNorm_np, llama3.2_model_numpy.py:256-278):
  //   producer (EPI_RESID_NORM): x += acc (fp32, in place); a_out = bf16(x * gamma_next); ss_out[row][32-column block]
  //                              = sum of x^2 over that block -- "x * gamma" now, "* 1/sqrt(mean(x^2)+eps)" later
  //   consumer (ss_in != nullptr, EPI_QKV_ROPE / EPI_GLU / EPI_F32): its A operand is such an a_out; every accumulator
  //                              row is scaled by rs = 1/sqrt(sum(ss_in[row][0..ss_n)) / norm_dim + eps) first (the GEMM is
  //                              linear in its rows, so scaling after the contraction equals normalising before it)
  __nv_bfloat16* a_out;    // [M, N] bf16
  const float* gamma_next; // [N]
  float* ss_out;           // [M, N/32]
  const float* ss_in;      // [M, ss_n]
  int ss_n, norm_dim;
  float norm_eps;
};

// 1/sqrt(mean(x^2) + eps) of a row whose partial sums of squares were left by an EPI_RESID_NORM producer (or by
// embed_norm_row_kernel); fixed summation order => deterministic and independent of the batch neighbours
// All loads are issued before the first add (one L2 round trip instead of ss_n / 4 dependent ones): the callers run
// this once per kernel, while the main loop of their first tile is still streaming weights.
constexpr int SS_MAX4 = 24;   // float4 per row: hidden_size <= 3072
__device__ __forceinline__ float row_rs(const GemmParams& p, int row) {
  const float4* s = reinterpret_cast<const float4*>(p.ss_in + static_cast<size_t>(row) * p.ss_n);
  const int n4 = p.ss_n >> 2;
  float4 v[SS_MAX4];
#pragma unroll
  for (int i = 0; i < SS_MAX4; ++i)
    if (i < n4) v[i] = __ldcg(s + i);
  float t = 0.f;
#pragma unroll
  for (int i = 0; i < SS_MAX4; ++i)
    if (i < n4) t += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  return 1.0f / sqrtf(t / static_cast<float>(p.norm_dim) + p.norm_eps);
}

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;
constexpr int GEMM_EPI_WARPS = 8;
constexpr int GEMM_THREADS = 64 + 32 * GEMM_EPI_WARPS;

template <int BN>
struct GemmCfg {
  static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;
  static constexpr int B_BYTES = BN * GEMM_BK * 2;
  static constexpr int STAGES = (BN >= 256) ? 4 : (BN >= 128 ? 6 : 8);
  static constexpr int STAGING_BYTES = GEMM_EPI_WARPS * 4096;   // one 32x32 fp32 slab per epilogue warp
  static constexpr int BAR_BYTES = 256;
  static constexpr int SMEM_BYTES = STAGES * (A_BYTES + B_BYTES) + STAGING_BYTES + BAR_BYTES + 1024;  // +1024: alignment
  static constexpr int TMEM_COLS = (2 * BN < 32) ? 32 : 2 * BN;
};

__device__ __forceinline__ float act_silu(float g) { return __fdividef(g, 1.0f + __expf(-g)); }
__device__ __forceinline__ float act_gelu_tanh(float g) {
  const float k = 0.7978845608028654f;
  return 0.5f * g * (1.0f + tanhf(k * (g + 0.044715f * g * g * g)));
}

__device__ __forceinline__ void store_bf16x32(__nv_bfloat16* dst, const float (&v)[32]) {
  uint4* d = reinterpret_cast<uint4*>(dst);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    uint4 u;
    u.x = pack_bf16(v[8 * i + 0], v[8 * i + 1]);
    u.y = pack_bf16(v[8 * i + 2], v[8 * i + 3]);
    u.z = pack_bf16(v[8 * i + 4], v[8 * i + 5]);
    u.w = pack_bf16(v[8 * i + 6], v[8 * i + 7]);
    d[i] = u;
  }
}
// residual of a bf16 rounding, itself rounded to bf16 (the "lo" plane of the precise mode)
__device__ __forceinline__ void split_lo(const float (&v)[32], float (&lo)[32]) {
#pragma unroll
  for (int i = 0; i < 32; ++i) lo[i] = v[i] - __bfloat162float(__float2bfloat16_rn(v[i]));
}

// Epilogue of one tile for one warp.  `quarter` = TMEM lane quarter (rows), `half` = which of the two
// warps sharing that quarter (takes column groups of its parity), `slab` = this warp's 4 KB staging.
template <int BN, int EPI>
// `rs_in`: 1/rms of this thread's row when the A operand came out of a split RMSNorm (p.ss_in != nullptr), computed once
// per kernel by the caller (row_rs); ignored otherwise.
__device__ __forceinline__ void gemm_epilogue(const GemmParams& p, const CUtensorMap* tmap_c, uint8_t* slab,
                                              uint32_t tacc, int mb, int nb, int ks, int quarter, int half, int lane,
                                              float rs_in = 1.0f) {
  const int row0 = mb * GEMM_BM + quarter * 32;
  const int row = row0 + lane;
  const bool row_ok = row < p.M;
  const int col0 = nb * BN;
  const uint32_t trow = tacc + (static_cast<uint32_t>(quarter * 32) << 16);

  if constexpr (EPI == EPI_F32 || EPI == EPI_RESID) {
    // registers -> swizzled smem slab (row = lane, 16-byte chunk j at j ^ (lane & 7)) -> TMA store / reduce-add
    const float inv_cap = p.softcap > 0.f ? 1.0f / p.softcap : 0.f;
    const float rs = (EPI == EPI_F32 && p.ss_in) ? rs_in : 1.0f;
#pragma unroll 1
    for (int c = half; c < BN / 32; c += 2) {
      const int col = col0 + c * 32;
      if (col >= p.N) break;
      uint32_t r[32];
      tmem_ld32(trow + c * 32, r);
      tmem_ld_wait();
      if (EPI == EPI_F32 && p.ss_in) {
#pragma unroll
        for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * rs);
      }
      if (EPI == EPI_F32 && p.softcap > 0.f) {
#pragma unroll
        for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(p.softcap * tanhf(__uint_as_float(r[i]) * inv_cap));
      }
      if (lane == 0) tma_store_wait_read<0>();   // the previous bulk store has finished reading the slab
      __syncwarp();
      uint8_t* srow = slab + lane * 128;
#pragma unroll
      for (int j = 0; j < 8; ++j)
        *reinterpret_cast<uint4*>(srow + ((j ^ (lane & 7)) << 4)) = make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0 && row0 < p.M) {
        if (EPI == EPI_F32) {
          if (p.stream_out) tma_store_2d_hint(tmap_c, slab, col, ks * p.split_rows + row0, kEvictFirst);
          else tma_store_2d(tmap_c, slab, col, ks * p.split_rows + row0);
        }
        else tma_reduce_add_2d(tmap_c, slab, col, row0);   // x += acc  (fp32 residual stream)
        tma_store_commit();
      }
    }
  } else if constexpr (EPI == EPI_BF16) {
#pragma unroll 1
    for (int c = half; c < BN / 32; c += 2) {
      const int col = col0 + c * 32;
      if (col >= p.N) break;
      __syncwarp();  // tcgen05.ld is .sync.aligned: reconverge after the predicated stores
      uint32_t r[32];
      tmem_ld32(trow + c * 32, r);
      tmem_ld_wait();
      if (!row_ok) continue;
      float v[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
      __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + static_cast<size_t>(row) * p.ldo + col;
      store_bf16x32(o, v);
      if (p.lo_off) {
        float lo[32];
        split_lo(v, lo);
        store_bf16x32(o + p.lo_off, lo);
      }
    }
  } else if constexpr (EPI == EPI_GLU) {
    // W rows are interleaved in 32-row groups: [gate 32j..32j+31 | up 32j..32j+31] -> column chunk 2j is
    // gate, chunk 2j+1 is up, output column = N/2-space (col0/2 + 32j).
    const float rs = p.ss_in ? rs_in : 1.0f;
#pragma unroll 1
    for (int j = half; j < BN / 64; j += 2) {
      const int col = col0 + j * 64;
      if (col >= p.N) break;
      __syncwarp();
      uint32_t g[32], u[32];
      tmem_ld32(trow + j * 64, g);
      tmem_ld32(trow + j * 64 + 32, u);
      tmem_ld_wait();
      if (!row_ok) continue;
      float v[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float gv = __uint_as_float(g[i]) * rs;
        const float a = p.act == 0 ? act_silu(gv) : act_gelu_tanh(gv);
        v[i] = a * (__uint_as_float(u[i]) * rs);
      }
      __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + static_cast<size_t>(row) * p.ldo + (col >> 1);
      store_bf16x32(o, v);
      if (p.lo_off) {
        float lo[32];
        split_lo(v, lo);
        store_bf16x32(o + p.lo_off, lo);
      }
    }
  } else if constexpr (EPI == EPI_RESID_NORM) {
    // x += acc; a' = bf16(x * gamma_next); ss[row][column block] = sum x^2   (row = lane: each thread owns 32 columns of its row)
    float* xout = reinterpret_cast<float*>(p.out);
#pragma unroll 1
    for (int c = half; c < BN / 32; c += 2) {
      const int col = col0 + c * 32;
      if (col >= p.N) break;
      __syncwarp();
      uint32_t r[32];
      tmem_ld32(trow + c * 32, r);
      tmem_ld_wait();
      if (!row_ok) continue;
      float4* xr = reinterpret_cast<float4*>(xout + static_cast<size_t>(row) * p.ldo + col);
      const float4* g4 = reinterpret_cast<const float4*>(p.gamma_next + col);
      float v[32], ss = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float4 xv = xr[j];
        xv.x += __uint_as_float(r[4 * j]);
        xv.y += __uint_as_float(r[4 * j + 1]);
        xv.z += __uint_as_float(r[4 * j + 2]);
        xv.w += __uint_as_float(r[4 * j + 3]);
        xr[j] = xv;
        ss += (xv.x * xv.x + xv.y * xv.y) + (xv.z * xv.z + xv.w * xv.w);
        const float4 gg = __ldg(g4 + j);
        v[4 * j] = xv.x * gg.x;
        v[4 * j + 1] = xv.y * gg.y;
        v[4 * j + 2] = xv.z * gg.z;
        v[4 * j + 3] = xv.w * gg.w;
      }
      store_bf16x32(p.a_out + static_cast<size_t>(row) * p.N + col, v);
      p.ss_out[static_cast<size_t>(row) * (p.N / 32) + (col >> 5)] = ss;
    }
  } else {  // EPI_QKV_ROPE
    const int HD = p.head_dim;
    const int half_chunks = HD >> 6;  // 32-column chunks per half head
    const int t0 = *p.t0_ptr;
    const int b = row / p.S, s = row - b * p.S;
    const int pos = t0 + s;                                    // cache slot of this token
    const int rpos = (p.pos_ids && row_ok) ? p.pos_ids[s] : pos;   // position the rotary table is read at
    const float* cosr = p.rope_cos + static_cast<size_t>(rpos) * (HD >> 1);
    const float* sinr = p.rope_sin + static_cast<size_t>(rpos) * (HD >> 1);
    // work items: (head hh, half-chunk i); this warp takes the items of its parity
    float4 cs4[8], sn4[8];   // cos/sin of this row's position for half-chunk tab_i (reused across the heads of a tile)
    int tab_i = -1;
    // row scale of a split RMSNorm (decode step only: few-row tiles; the 256-wide prefill instantiations skip it)
    constexpr bool SCALE = BN <= 128;
    const float rs = (SCALE && p.ss_in) ? rs_in : 1.0f;
#pragma unroll 1
    for (int t = half; t < BN / 64; t += 2) {
      const int hh = t / half_chunks, i = t - hh * half_chunks;
      const int n0 = col0 + hh * HD;  // first fused-QKV column of this head
      if (n0 >= p.N) break;
      __nv_bfloat16* dst;
      bool rope = true;
      int lo_off = 0;
      if (n0 < p.q_cols) {
        dst = p.q_out + static_cast<size_t>(row) * p.ldq + n0;
        lo_off = p.lo_off;
      } else {
        int n1 = n0 - p.q_cols;
        __nv_bfloat16* base = p.k_cache;
        if (n1 >= p.kv_cols) {
          n1 -= p.kv_cols;
          base = p.v_cache;
          rope = false;
        }
        const int kvh = n1 / HD;
        // rows past M (the tile's padding) have no batch row: no block-table lookup for them (they store nothing)
        dst = base + (row_ok ? kv_row(p.pages, b, kvh, p.nkv, p.Tmax, pos) : 0) * HD;
        lo_off = p.kv_lo_off;
      }
      __syncwarp();
      uint32_t a[32], c[32];
      const int cl = hh * (HD >> 5) + i;  // chunk holding x1 (first half of the head)
      tmem_ld32(trow + cl * 32, a);
      tmem_ld32(trow + (cl + half_chunks) * 32, c);  // matching x2 chunk (second half)
      if (rope && row_ok && i != tab_i) {   // table loads overlap the TMEM load latency; head_dim 64: one load per tile
        const float4* c4 = reinterpret_cast<const float4*>(cosr + i * 32);
        const float4* s4 = reinterpret_cast<const float4*>(sinr + i * 32);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          cs4[j] = __ldg(c4 + j);
          sn4[j] = __ldg(s4 + j);
          if (SCALE) {   // fold 1/rms into the rotation: (p * rs) * cos == p * (rs * cos)
            cs4[j].x *= rs; cs4[j].y *= rs; cs4[j].z *= rs; cs4[j].w *= rs;
            sn4[j].x *= rs; sn4[j].y *= rs; sn4[j].z *= rs; sn4[j].w *= rs;
          }
        }
        tab_i = i;
      }
      __syncwarp();
      tmem_ld_wait();
      if (!row_ok) continue;
      float x1[32], x2[32];
      if (rope) {
        // q*cos + rotate_half(q)*sin with rotate_half = [-x2, x1]  (llama3.2_model_numpy.py:69-90)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float cs[4] = {cs4[j].x, cs4[j].y, cs4[j].z, cs4[j].w}, sn[4] = {sn4[j].x, sn4[j].y, sn4[j].z, sn4[j].w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float p1 = __uint_as_float(a[4 * j + e]), p2 = __uint_as_float(c[4 * j + e]);
            x1[4 * j + e] = p1 * cs[e] - p2 * sn[e];
            x2[4 * j + e] = p2 * cs[e] + p1 * sn[e];
          }
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          x1[j] = SCALE ? __uint_as_float(a[j]) * rs : __uint_as_float(a[j]);
          x2[j] = SCALE ? __uint_as_float(c[j]) * rs : __uint_as_float(c[j]);
        }
      }
      store_bf16x32(dst + i * 32, x1);
      store_bf16x32(dst + (HD >> 1) + i * 32, x2);
      if (lo_off) {
        float lo[32];
        split_lo(x1, lo);
        store_bf16x32(dst + lo_off + i * 32, lo);
        split_lo(x2, lo);
        store_bf16x32(dst + lo_off + (HD >> 1) + i * 32, lo);
      }
    }
  }
}

template <int BN, int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                 const __grid_constant__ CUtensorMap tmap_c, const GemmParams p) {
  using Cfg = GemmCfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * Cfg::A_BYTES;
  uint8_t* sStage = smem + STAGES * (Cfg::A_BYTES + Cfg::B_BYTES);     // 1024-byte aligned
  uint64_t* full = reinterpret_cast<uint64_t*>(sStage + Cfg::STAGING_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_blocks = (p.M + GEMM_BM - 1) / GEMM_BM;
  const int n_blocks = (p.N + BN - 1) / BN;
  const int total = m_blocks * n_blocks * p.split_k;
  const int KB = (p.K / GEMM_BK) / p.split_k;   // k-blocks per tile
  const int per_band = p.band_m * n_blocks;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    if (EPI == EPI_F32 || EPI == EPI_RESID) tma_prefetch_desc(&tmap_c);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull[a], 1);
      mbar_init(&tempty[a], GEMM_EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  const int a_rows_box = p.a_box_rows > 0 ? p.a_box_rows : GEMM_BM;
  const uint32_t a_tx = static_cast<uint32_t>(a_rows_box) * 128u;       // bytes of A per ring slot actually loaded
  if (a_rows_box < GEMM_BM) {                  // few-row GEMM (decode): rows the box does not cover stay zero
    const int tail16 = (GEMM_BM - a_rows_box) * 8;                       // 16-byte chunks per slot
    for (int i = threadIdx.x; i < STAGES * tail16; i += GEMM_THREADS) {
      const int s = i / tail16, c = i - s * tail16;
      *reinterpret_cast<uint4*>(sA + s * Cfg::A_BYTES + a_rows_box * 128 + c * 16) = make_uint4(0u, 0u, 0u, 0u);
    }
    fence_proxy_async_smem();
  }
  pdl_trigger();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // Everything above overlapped the previous kernel's tail (PDL).  The wait for the previous kernel sits inside the
  // roles: the producer first streams the W tiles of its first k-steps (weights are constants, nothing in the stream
  // writes them) and only then waits to load activations; the epilogue warps wait before they touch global memory;
  // the MMA issuer only ever sees shared memory through the barriers.

  auto tile_coords = [&](int tile, int& mb, int& nb, int& kb0) {
    kb0 = (tile % p.split_k) * KB;
    tile /= p.split_k;
    const int band = tile / per_band, r = tile - band * per_band;
    const int mrem = m_blocks - band * p.band_m;
    const int mcnt = mrem < p.band_m ? mrem : p.band_m;
    nb = r / mcnt;
    mb = band * p.band_m + (r - nb * mcnt);
  };

  if (warp == 0) {
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      int pre = 0;                               // ring slots whose W tile is already in flight (first tile only)
      if (static_cast<int>(blockIdx.x) < total) {
        int mb, nb, kb0;
        tile_coords(blockIdx.x, mb, nb, kb0);
        pre = KB < STAGES ? KB : STAGES;
        for (int kb = 0; kb < pre; ++kb) {
          mbar_expect_tx(&full[kb], a_tx + Cfg::B_BYTES);
          const int ka = (kb0 + kb) * GEMM_BK;
          const int kw = ka >= p.k_wrap ? ka - p.k_wrap : ka;
          tma_load_2d(sB + kb * Cfg::B_BYTES, &tmap_b, &full[kb], kw, nb * BN, kEvictNormal);
        }
      }
      pdl_wait();                                // activations come from the preceding kernels
      for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
        int mb, nb, kb0;
        tile_coords(tile, mb, nb, kb0);
        for (int kb = 0; kb < KB; ++kb) {
          const int ka = (kb0 + kb) * GEMM_BK;
          const int kw = ka >= p.k_wrap ? ka - p.k_wrap : ka;
          if (pre > 0) {                         // slot armed and W requested above: only the activations are missing
            --pre;
            tma_load_2d(sA + stage * Cfg::A_BYTES, &tmap_a, &full[stage], ka, mb * GEMM_BM, kEvictNormal);
          } else {
            mbar_wait(&empty[stage], phase ^ 1);
            mbar_expect_tx(&full[stage], a_tx + Cfg::B_BYTES);
            tma_load_2d(sA + stage * Cfg::A_BYTES, &tmap_a, &full[stage], ka, mb * GEMM_BM, kEvictNormal);
            tma_load_2d(sB + stage * Cfg::B_BYTES, &tmap_b, &full[stage], kw, nb * BN, kEvictNormal);
          }
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(GEMM_BM, BN, 0, 0);
      int stage = 0;
      uint32_t phase = 0, acc = 0, acc_phase = 0;
      for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
        mbar_wait(&tempty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d = tmem_base + acc * BN;
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(sA + stage * Cfg::A_BYTES);
          const uint32_t b_addr = smem_u32(sB + stage * Cfg::B_BYTES);
#pragma unroll
          for (int k = 0; k < GEMM_BK / 16; ++k) {
            const uint64_t ad = umma_smem_desc(a_addr + k * 32, 16, 1024);
            const uint64_t bd = umma_smem_desc(b_addr + k * 32, 16, 1024);
            umma_bf16_ss(d, ad, bd, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty[stage]);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&tfull[acc]);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else {
    const int ew = warp - 2;                 // 0..7
    const int quarter = warp & 3;            // TMEM lanes [32*quarter, +32) are the only ones this warp may touch
    const int half = ew >> 2;                // two warps share a quarter; each takes every other column group
    uint8_t* slab = sStage + ew * 4096;
    uint32_t acc = 0, acc_phase = 0;
    pdl_wait();                                // residual / tables are read and outputs written only after this
    // split RMSNorm (decode step, one m-block): this thread's row scale, fetched while the first tile's weights stream
    float rs_row = 1.0f;
    if (p.ss_in && quarter * 32 + lane < p.M) rs_row = row_rs(p, quarter * 32 + lane);
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
      int mb, nb, kb0;
      tile_coords(tile, mb, nb, kb0);
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      gemm_epilogue<BN, EPI>(p, &tmap_c, slab, tmem_base + acc * BN, mb, nb, kb0 / KB, quarter, half, lane, rs_row);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if ((EPI == EPI_F32 || EPI == EPI_RESID) && lane == 0) tma_store_wait_all();   // bulk stores drained before exit
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

}  // namespace b2
