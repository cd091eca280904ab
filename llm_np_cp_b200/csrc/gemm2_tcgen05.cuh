// 2-CTA variant of the persistent GEMM (gemm_tcgen05.cuh): a cluster of two CTAs on one TPC computes a 256 x 256
// output tile with ONE tcgen05.mma.cta_group::2 stream (M = 256).  Each CTA stages its own 128 rows of A and only HALF
// of the W tile (128 of the 256 rows), so the shared-memory fill traffic per SM drops from 48 KB to 32 KB per k-block
// and the ring can be 6 deep.  The leader CTA's elected thread issues the MMAs for the pair; completion is multicast
// to both CTAs' barriers; each CTA runs the epilogue (same fused epilogues, gemm_epilogue<>) for its 128 rows.
#pragma once
#include "gemm_tcgen05.cuh"

namespace b2 {

struct Gemm2Cfg {
  static constexpr int BN = 256;
  static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;        // this CTA's 128 rows of A
  static constexpr int B_BYTES = (BN / 2) * GEMM_BK * 2;       // this CTA's half of the W tile
  static constexpr int STAGES = 6;
  static constexpr int STAGING_BYTES = GEMM_EPI_WARPS * 4096;
  static constexpr int BAR_BYTES = 256;
  static constexpr int SMEM_BYTES = STAGES * (A_BYTES + B_BYTES) + STAGING_BYTES + BAR_BYTES + 1024;
  static constexpr int TMEM_COLS = 512;
};

template <int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
gemm2_bf16_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                  const __grid_constant__ CUtensorMap tmap_c, const GemmParams p) {
  using Cfg = Gemm2Cfg;
  constexpr int STAGES = Cfg::STAGES, BN = Cfg::BN;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * Cfg::A_BYTES;
  uint8_t* sStage = smem + STAGES * (Cfg::A_BYTES + Cfg::B_BYTES);
  uint64_t* full = reinterpret_cast<uint64_t*>(sStage + Cfg::STAGING_BYTES);   // used in the leader CTA only
  uint64_t* empty = full + STAGES;                                             // per CTA
  uint64_t* tfull = empty + STAGES;                                            // per CTA
  uint64_t* tempty = tfull + 2;                                                // used in the leader CTA only
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int m_blocks = (p.M + 2 * GEMM_BM - 1) / (2 * GEMM_BM);               // 256-row tiles
  const int n_blocks = (p.N + BN - 1) / BN;
  const int total = m_blocks * n_blocks;
  const int KB = p.K / GEMM_BK;
  const int per_band = p.band_m * n_blocks;
  const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;

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
      mbar_init(&tempty[a], 2 * GEMM_EPI_WARPS);     // epilogue warps of BOTH CTAs arrive on the leader's barrier
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_2sm(tmem_slot, Cfg::TMEM_COLS);
  pdl_trigger();
  tc_fence_before();
  cluster_sync_all();                                 // peer barriers are initialised before any remote arrive
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  auto tile_coords = [&](int tile, int& mb, int& nb) {   // mb in units of 256 rows; band_m in the same units
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
      for (int tile = cluster_id; tile < total; tile += n_clusters) {
        int mb, nb;
        tile_coords(tile, mb, nb);
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          const uint32_t full_leader = mapa_u32(smem_u32(&full[stage]), 0);
          if (leader) mbar_expect_tx(&full[stage], 2 * (Cfg::A_BYTES + Cfg::B_BYTES));   // both CTAs' bytes
          const int ka = kb * GEMM_BK;
          const int kw = ka >= p.k_wrap ? ka - p.k_wrap : ka;
          tma_load_2d_2sm(sA + stage * Cfg::A_BYTES, &tmap_a, full_leader, ka, mb * 2 * GEMM_BM + rank * GEMM_BM,
                          p.stream_out ? kEvictLast : kEvictNormal);
          tma_load_2d_2sm(sB + stage * Cfg::B_BYTES, &tmap_b, full_leader, kw, nb * BN + rank * (BN / 2),
                          p.stream_out ? kEvictFirst : kEvictNormal);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (leader && elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(2 * GEMM_BM, BN, 0, 0);   // M = 256 across the pair
      int stage = 0;
      uint32_t phase = 0, acc = 0, acc_phase = 0;
      for (int tile = cluster_id; tile < total; tile += n_clusters) {
        mbar_wait(&tempty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d = tmem_base + acc * BN;
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(sA + stage * Cfg::A_BYTES);
          const uint32_t b_addr = smem_u32(sB + stage * Cfg::B_BYTES);
#pragma unroll
          for (int k = 0; k < GEMM_BK / 16; ++k)
            umma_bf16_ss_2sm(d, umma_smem_desc(a_addr + k * 32, 16, 1024), umma_smem_desc(b_addr + k * 32, 16, 1024), idesc,
                             (kb | k) != 0 ? 1u : 0u);
          umma_commit_2sm(&empty[stage], 3);           // frees the ring slot in both CTAs
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit_2sm(&tfull[acc], 3);               // accumulator ready in both CTAs
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else {
    const int ew = warp - 2;
    const int quarter = warp & 3;
    const int half = ew >> 2;
    uint8_t* slab = sStage + ew * 4096;
    uint32_t acc = 0, acc_phase = 0;
    for (int tile = cluster_id; tile < total; tile += n_clusters) {
      int mb, nb;
      tile_coords(tile, mb, nb);
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      gemm_epilogue<BN, EPI>(p, &tmap_c, slab, tmem_base + acc * BN, 2 * mb + static_cast<int>(rank), nb, 0, quarter, half, lane);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&tempty[acc]), 0));   // the leader's barrier
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if ((EPI == EPI_F32 || EPI == EPI_RESID) && lane == 0) tma_store_wait_all();
  }
  tc_fence_before();
  cluster_sync_all();                                 // nobody exits (or frees TMEM) while the peer may still touch it
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, Cfg::TMEM_COLS);
  }
}

}  // namespace b2
