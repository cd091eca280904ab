// Few-row ("skinny", M <= 64) GEMM for the decode step:  C[M,N] = A[M,K] . W[N,K]^T with the K range split over the CTAs
// of a thread-block CLUSTER and the partial accumulators reduced through distributed shared memory.
//
// Why: at M = batch = 64 every weight matrix is read once per token step and the GEMMs are pure HBM streaming.  A
// 64-column tile per CTA gives only N/64 CTAs (48 for the fused QKV weight, 32 for o_proj / down_proj), far fewer
// than the 148 SMs whose combined TMA issue rate is needed to saturate HBM.  gemm_bf16_kernel's answer (split-K with
// fp32 partial planes in global memory, folded by the following RMSNorm kernel) costs an extra kernel between every
// pair of GEMMs and cannot be combined with epilogues that need the finished sum (RoPE + cache append).  Here the CTAs
// that share an output tile form a cluster (one k-range each); ranks 1.. ship their TMEM accumulators to rank 0's
// shared memory (st.shared::cluster), rank 0 adds them in rank order (deterministic) and runs the usual fused epilogue
// on the finished sum:
//   EPI_QKV_ROPE    q/k/v projection + RoPE + KV-cache append (llama3.2_model_numpy.py:419-465), on 2x as many CTAs
//   EPI_RESID_NORM  o_proj / down_proj + residual add (:514 + :569, :176-182 + :578) AND the first half of the
//                   RMSNorm that follows (:256-278): writes x (fp32), a' = bf16(x * gamma_next) -- the next GEMM's A
//                   operand -- and per-(row, column block) sums of squares; the consumer GEMM's epilogue scales its
//                   rows by 1/sqrt(mean(x^2) + eps).  The separate norm kernels disappear from the decode step.
//
// One (tile, k-range) per CTA, no persistence: a decode GEMM is a single wave by construction.
#pragma once
#include "gemm_tcgen05.cuh"

namespace b2 {

constexpr int SK_BN = 64;
constexpr int SK_STAGES = 8;
constexpr int SK_MAX_ROWS = 64;     // rows actually present (TMA box rows); the MMA tile is 128 rows, rows >= 64 are never read back
constexpr int SK_MAX_CLUSTER = 4;   // CTAs per output tile (the peers' partial tiles live in rank 0's shared memory)

struct GemmSkCfg {
  static constexpr int A_BYTES = SK_MAX_ROWS * GEMM_BK * 2;          // 8 KB: the 64 rows that exist (the 128-row UMMA reads past it
                                                                     // into the next slot; those accumulator lanes are never read)
  static constexpr int B_BYTES = SK_BN * GEMM_BK * 2;                // 8 KB
  static constexpr int RING_BYTES = SK_STAGES * (A_BYTES + B_BYTES) + 8192;   // + slack for the last A slot's over-read
  static constexpr int RED_BYTES = (SK_MAX_CLUSTER - 1) * SK_MAX_ROWS * SK_BN * 4;   // 48 KB: peers' partial tiles
  static constexpr int STAGING_BYTES = 0;                            // the epilogues used here store straight from registers
  static constexpr int BAR_BYTES = 256;
  static constexpr int SMEM_BYTES = RING_BYTES + RED_BYTES + STAGING_BYTES + BAR_BYTES + 1024;
  static constexpr int TMEM_COLS = 64;
};
static_assert(GemmSkCfg::SMEM_BYTES <= 232448, "shared memory budget");

__device__ __forceinline__ void st_cluster_f4(uint32_t cluster_addr, float a, float b, float c, float d) {
  asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(cluster_addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

template <int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_sk_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
               const __grid_constant__ CUtensorMap tmap_c, const GemmParams p) {
  using Cfg = GemmSkCfg;
  constexpr int BN = SK_BN, STAGES = SK_STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * Cfg::A_BYTES;
  uint8_t* sRed = smem + Cfg::RING_BYTES;                                   // [peer][row 0..63][64 fp32], 16-byte chunks swizzled
  uint8_t* sStage = sRed + Cfg::RED_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(sStage + Cfg::STAGING_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int CK = p.split_k;                                 // cluster size = CTAs per output tile
  const uint32_t rank = CK > 1 ? cluster_ctarank() : 0u;
  const int tile = blockIdx.x / CK;                         // clusters are consecutive blocks
  const int n_blocks = (p.N + BN - 1) / BN;
  const int nb = tile % n_blocks, mb = tile / n_blocks;     // mb is always 0 for M <= 64; kept for clarity
  const int KB = (p.K / GEMM_BK) / CK;                      // k-blocks of this CTA
  const int kb0 = static_cast<int>(rank) * KB;
  const int a_rows = p.a_box_rows;                          // 32 or 64
  const uint32_t a_tx = static_cast<uint32_t>(a_rows) * 128u;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(tfull, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  pdl_trigger();
  tc_fence_before();
  // Cluster barrier #0 ("every CTA of the cluster is running, its shared memory may be written remotely") is split:
  // arrive here, wait only right before the first remote store -- the weight stream starts without it.
  if (CK > 1) asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      // weights are constants: the first ring slots' W tiles are requested before waiting for the previous kernel
      const int pre = KB < STAGES ? KB : STAGES;
      for (int kb = 0; kb < pre; ++kb) {
        mbar_expect_tx(&full[kb], a_tx + Cfg::B_BYTES);
        tma_load_2d(sB + kb * Cfg::B_BYTES, &tmap_b, &full[kb], (kb0 + kb) * GEMM_BK, nb * BN, kEvictFirst);
      }
      pdl_wait();                        // activations come from the preceding kernels
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = 0; kb < KB; ++kb) {
        const int ka = (kb0 + kb) * GEMM_BK;
        if (kb < pre) {
          tma_load_2d(sA + stage * Cfg::A_BYTES, &tmap_a, &full[stage], ka, mb * GEMM_BM, kEvictLast);
        } else {
          mbar_wait(&empty[stage], phase ^ 1);
          mbar_expect_tx(&full[stage], a_tx + Cfg::B_BYTES);
          tma_load_2d(sA + stage * Cfg::A_BYTES, &tmap_a, &full[stage], ka, mb * GEMM_BM, kEvictLast);
          tma_load_2d(sB + stage * Cfg::B_BYTES, &tmap_b, &full[stage], ka, nb * BN, kEvictFirst);
        }
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(GEMM_BM, BN, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = 0; kb < KB; ++kb) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(sA + stage * Cfg::A_BYTES);
        const uint32_t b_addr = smem_u32(sB + stage * Cfg::B_BYTES);
#pragma unroll
        for (int k = 0; k < GEMM_BK / 16; ++k)
          umma_bf16_ss(tmem_base, umma_smem_desc(a_addr + k * 32, 16, 1024), umma_smem_desc(b_addr + k * 32, 16, 1024), idesc,
                       (kb | k) != 0 ? 1u : 0u);
        umma_commit(&empty[stage]);
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
      umma_commit(tfull);
    }
  }

  // ---- epilogue warps (2..9): two per TMEM lane quarter; only the quarters that hold real rows (< a_rows) work ----
  const int ew = warp - 2, quarter = warp & 3, half = ew >> 2;
  const bool epi_warp = warp >= 2 && quarter * 32 < a_rows;
  const int row_in_tile = quarter * 32 + lane;
  const int row = mb * GEMM_BM + row_in_tile;
  const bool row_ok = row < p.M;
  const uint32_t trow = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
  const int col = nb * BN + half * 32;                       // this warp's 32 columns of the tile
  uint32_t r[32];                                            // this thread's 32 accumulators (row = lane of the quarter)
  float4 xv[8];                                              // EPI_RESID_NORM: the residual stream values they are added to
  float rs_row = 1.0f;                                       // EPI_QKV_ROPE: 1/rms of the row (split RMSNorm)
  if (epi_warp) {
    if (rank == 0) {   // everything the epilogue needs from earlier kernels is fetched while the weights stream
      pdl_wait();
      if (EPI == EPI_RESID_NORM && row_ok) {
        const float4* xr = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.out) + static_cast<size_t>(row) * p.ldo + col);
#pragma unroll
        for (int j = 0; j < 8; ++j) xv[j] = __ldcg(xr + j);
      }
      if (EPI == EPI_QKV_ROPE && p.ss_in && row_ok) rs_row = row_rs(p, row);
    }
    mbar_wait(tfull, 0);
    tc_fence_after();
    tmem_ld32(trow + half * 32, r);
    tmem_ld_wait();
  }
  if (CK > 1) {
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");     // barrier #0: the peers are running
    if (epi_warp && rank != 0) {   // ship this CTA's partial tile to rank 0: 32 columns of this thread's row
      const uint32_t base = mapa_u32(smem_u32(sRed), 0) + (rank - 1) * (SK_MAX_ROWS * BN * 4) + row_in_tile * (BN * 4);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int chunk = half * 8 + j;                                     // 16-byte chunk inside the 256-byte row
        st_cluster_f4(base + ((chunk ^ (row_in_tile & 15)) << 4), __uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]),
                      __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
      }
    }
    tc_fence_before();
    cluster_sync_all();                  // barrier #1 (release / acquire at cluster scope): the partial tiles are in rank 0's smem
    tc_fence_after();
  }
  if (epi_warp && rank == 0) {
    for (int pr = 0; pr < CK - 1; ++pr) {   // add the peers' tiles in rank order: deterministic
      const uint8_t* src = sRed + pr * (SK_MAX_ROWS * BN * 4) + row_in_tile * (BN * 4);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int chunk = half * 8 + j;
        const float4 v = *reinterpret_cast<const float4*>(src + ((chunk ^ (row_in_tile & 15)) << 4));
        r[4 * j] = __float_as_uint(__uint_as_float(r[4 * j]) + v.x);
        r[4 * j + 1] = __float_as_uint(__uint_as_float(r[4 * j + 1]) + v.y);
        r[4 * j + 2] = __float_as_uint(__uint_as_float(r[4 * j + 2]) + v.z);
        r[4 * j + 3] = __float_as_uint(__uint_as_float(r[4 * j + 3]) + v.w);
      }
    }
    if constexpr (EPI == EPI_RESID_NORM) {
      // x += acc; a' = bf16(x * gamma_next); ss[row][32-column block] = sum x^2  (same arithmetic as gemm_epilogue's branch)
      if (row_ok) {
        float4* xw = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + static_cast<size_t>(row) * p.ldo + col);
        const float4* g4 = reinterpret_cast<const float4*>(p.gamma_next + col);
        float v[32], ss = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float4 x4 = xv[j];
          x4.x += __uint_as_float(r[4 * j]);
          x4.y += __uint_as_float(r[4 * j + 1]);
          x4.z += __uint_as_float(r[4 * j + 2]);
          x4.w += __uint_as_float(r[4 * j + 3]);
          xw[j] = x4;
          ss += (x4.x * x4.x + x4.y * x4.y) + (x4.z * x4.z + x4.w * x4.w);
          const float4 gg = __ldg(g4 + j);
          v[4 * j] = x4.x * gg.x;
          v[4 * j + 1] = x4.y * gg.y;
          v[4 * j + 2] = x4.z * gg.z;
          v[4 * j + 3] = x4.w * gg.w;
        }
        store_bf16x32(p.a_out + static_cast<size_t>(row) * p.N + col, v);
        p.ss_out[static_cast<size_t>(row) * (p.N / 32) + (col >> 5)] = ss;
      }
    } else {
      // QKV: the RoPE epilogue pairs the two 32-column chunks of a head (x1 | x2) that the two warps of this quarter hold:
      // put the finished sums back into TMEM and run the ordinary epilogue
      if (CK > 1) {
        tmem_st32(trow + half * 32, r);
        tmem_st_wait();
      }
      tc_fence_before();
      named_bar_sync(1 + quarter, 64);
      tc_fence_after();
      gemm_epilogue<BN, EPI>(p, &tmap_c, nullptr, tmem_base, mb, nb, 0, quarter, half, lane, rs_row);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

}  // namespace b2
