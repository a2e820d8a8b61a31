// tcgen05 implicit-GEMM convolution on CTA PAIRS (cta_group::2): the wide-N variant of tc_conv.cu.
//
// Why: tc_conv_kernel<128> is bound by SHARED-MEMORY bandwidth, not by the tensor pipe.  Per K block
// (64 channels of one tap) a CTA writes 64 KB (TMA: A hi/lo + B hi/lo) and the twelve exact-mode MMAs read
// 12 x (4 KB A + 4 KB B) = 96 KB, i.e. 160 KB per 768 MMA cycles = 208 B/clk against the SM's 128 B/clk
// (ncu on the layer-4 conv: 79.7 MB of shared-memory traffic per SM / 128 B/clk = 622 k cycles of the 627 k
// elapsed; tensor pipe ~50 % active).  Here two CTAs of a cluster share one M = 256 x N = 256 tile:
//   * each CTA stages its own 128 pixel rows of A and only HALF of the B rows (the tensor cores of the pair
//     exchange the halves over the SM-to-SM fabric), and one tcgen05.mma.cta_group::2 drives both SMs;
//   * per CTA and K block: 64 KB written + 12 x (4 KB A + 4 KB half-B) = 96 KB read for 1536 MMA cycles
//     = 104 B/clk -- under the shared-memory limit, so the kernel becomes MMA-bound.
// Protocol (both CTAs run every role; only the leader = even CTA issues MMAs):
//   warp 0  TMA producer: own A box + own half of B per stage; every load signals the LEADER's `full` barrier
//           (the leader expects the bytes of both CTAs); waits on its own `empty` barrier.
//   warp 1  leader: waits `full`, issues 12 MMAs (M256 x N x K16), commits with a 2-CTA multicast to the
//           `empty` barriers of both CTAs and, per finished chunk, to both `tmem_full` barriers.
//   warps 2-9  epilogue on the CTA's own TMEM half (rows of its M tile), same chunked fp32 register
//           accumulation / fused scale-bias-residual-activation-split store as tc_conv.cu; a drained
//           accumulator is released on the leader's `tmem_empty` barrier (one arrival per warp).
// Arithmetic is identical to tc_conv.cu (same products, same chunking), so results are bit-identical to it.
#include "tc_common.cuh"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <cuda.h>

namespace tdn {

template <int BLOCK_N>
struct PairCfg {
  static constexpr int B_HALF_PLANE = (BLOCK_N / 2) * TC_BLOCK_K * 2;          // this CTA's B rows, one plane
  static constexpr int STAGE_BYTES = 2 * TC_A_PLANE + 2 * B_HALF_PLANE;        // 64 KiB (N=256) / 48 KiB (N=128)
  static constexpr int STAGES = (200 * 1024) / STAGE_BYTES;                    // 3 / 4
  static constexpr int NUM_ACC = 512 / BLOCK_N;                                // chunk accumulators in TMEM
  static constexpr int TMEM_COLS = 512;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 512;
};

// QUAD (TDN_TC_PAIR_QUAD, an experiment kept as an explicit variant): clusters of FOUR CTAs = two pairs that work on the same
// output-channel tile of two consecutive rows of pair tiles and share its weights: each CTA fetches ONE plane (pair 0: hi,
// pair 1: lo) of its half of the B rows and multicasts it to the CTA of the same parity in the other pair.  The stage ring
// couples the two pairs: a stage is free when the MMAs of BOTH pairs that read it have retired (two multicast commits per
// `empty` barrier).  Hypothesis tested: the pair kernel is paced by L2 reads (the weights are half of its bytes and identical
// for every cluster).  Result: half the L2 reads of B move the tile time by 1.6 % (and only 33 clusters of four are
// resident against 74 of two) -- it is not; DESIGN.md section 10, third session.
template <int BLOCK_N, bool QUAD>
__global__ void __cluster_dims__(QUAD ? 4 : 2, 1, 1) __launch_bounds__(TC_THREADS, 1)
tc_conv_pair_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                    const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
                    const TcParams p) {
  using Cfg = PairCfg<BLOCK_N>;
  constexpr int STAGES = Cfg::STAGES;
  constexpr int NUM_ACC = Cfg::NUM_ACC;
  extern __shared__ uint8_t smem_raw[];
  // identical layout in both CTAs: descriptors and barrier offsets are shared by the pair
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + NUM_ACC;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + NUM_ACC);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int rank = (int)(blockIdx.x & 1);            // parity inside the pair (== %cluster_ctarank & 1)
  const bool leader = rank == 0;
  const int pair_in_cluster = QUAD ? (int)((blockIdx.x >> 1) & 1) : 0;
  const int cluster_id = (int)(blockIdx.x >> 1);
  const int num_clusters = (int)(gridDim.x >> 1);
  const int kc_per_tap = p.Cin / TC_BLOCK_K;
  const int num_kb = p.taps_h * p.taps_w * kc_per_tap;

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tmA_hi);
    prefetch_tensormap(&tmA_lo);
    prefetch_tensormap(&tmB_hi);
    prefetch_tensormap(&tmB_lo);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);                    // used on the leader only: its producer's expect_tx arrival
      mbar_init(&empty_bar[s], QUAD ? 2 : 1);        // multicast commit of the MMA thread of every pair of the cluster
    }
    for (int s = 0; s < NUM_ACC; ++s) {
      mbar_init(&tmem_full[s], 1);                   // multicast commit
      mbar_init(&tmem_empty[s], 2 * TC_EPI_WARPS);   // leader only: one arrival per epilogue warp of both CTAs
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc_pair(tmem_ptr, Cfg::TMEM_COLS);
    tmem_relinquish_pair();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                // the peer's barriers exist before anything signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  tc_pdl_sync();

  if (warp == 0) {
    // ======================= TMA producer (both CTAs) =======================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const uint16_t mc_mask = (uint16_t)(5u << rank);   // QUAD: this CTA and the CTA of the same parity in the other pair
      for (int tile = p.tile_begin + cluster_id; tile < p.num_tiles; tile += num_clusters) {
        int nt, mt;
        pair_tile_coords(p, tile, nt, mt);
        mt = 2 * mt + rank;                          // this CTA's 128-pixel M tile (may lie past the last image:
        const int tx = mt % p.tiles_w;               //  TMA then zero-fills and the epilogue stores nothing)
        mt /= p.tiles_w;
        const int ty = mt % p.tiles_h;
        const int img = mt / p.tiles_h;
        for (int kb = 0; kb < num_kb; ++kb) {
          const int tap = kb / kc_per_tap;
          const int kc = kb - tap * kc_per_tap;
          const int ky = tap / p.taps_w;
          const int kx = tap - ky * p.taps_w;
          const int x0 = tx * p.BW * p.conv_stride + (kx - (p.taps_w - 1) / 2) * p.dil;
          const int y0 = ty * p.BH * p.conv_stride + (ky - (p.taps_h - 1) / 2) * p.dil;
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
          uint8_t* sb = sa + 2 * TC_A_PLANE;
          if (leader) mbar_expect_tx(&full_bar[stage], 2 * Cfg::STAGE_BYTES);
          tma_load_4d_pair(sa, &tmA_hi, &full_bar[stage], kc * TC_BLOCK_K, x0, y0, img);
          tma_load_4d_pair(sa + TC_A_PLANE, &tmA_lo, &full_bar[stage], kc * TC_BLOCK_K, x0, y0, img);
          const int kcol = tap * p.Cin + kc * TC_BLOCK_K;
          const int brow = nt * BLOCK_N + rank * (BLOCK_N / 2);
          if (!QUAD) {
            tma_load_3d_pair(sb, &tmB_hi, &full_bar[stage], kcol, brow, 0);
            tma_load_3d_pair(sb + Cfg::B_HALF_PLANE, &tmB_lo, &full_bar[stage], kcol, brow, 0);
          } else if (pair_in_cluster == 0) {
            tma_load_3d_pair_mc(sb, &tmB_hi, &full_bar[stage], mc_mask, kcol, brow, 0);
          } else {
            tma_load_3d_pair_mc(sb + Cfg::B_HALF_PLANE, &tmB_lo, &full_bar[stage], mc_mask, kcol, brow, 0);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ======================= MMA issuer (leader CTA only) =======================
    if (leader) {
      constexpr uint32_t idesc = umma_idesc_f16(2 * TC_BLOCK_M, BLOCK_N);
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int tile = p.tile_begin + cluster_id; tile < p.num_tiles; tile += num_clusters) {
        for (int kb0 = 0; kb0 < num_kb; kb0 += p.chunk_kb) {
          const int kb1 = min(kb0 + p.chunk_kb, num_kb);
          mbar_wait(&tmem_empty[as], aphase ^ 1);
          const uint32_t d_tmem = tmem_base + as * BLOCK_N;
          for (int kb = kb0; kb < kb1; ++kb) {
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
            const uint32_t sb = sa + 2 * TC_A_PLANE;
            if (elect_one()) {
#pragma unroll
              for (int k = 0; k < TC_BLOCK_K / 16; ++k) {
                const uint64_t a_hi = umma_desc_k_sw128(sa + k * 32);
                const uint64_t a_lo = umma_desc_k_sw128(sa + TC_A_PLANE + k * 32);
                const uint64_t b_hi = umma_desc_k_sw128(sb + k * 32);
                const uint64_t b_lo = umma_desc_k_sw128(sb + Cfg::B_HALF_PLANE + k * 32);
                if (p.fast) {
                  umma_f16_pair(d_tmem, a_hi, b_hi, idesc, ((kb - kb0) | k) != 0);
                } else {
                  umma_f16_pair(d_tmem, a_hi, b_lo, idesc, ((kb - kb0) | k) != 0);
                  umma_f16_pair(d_tmem, a_lo, b_hi, idesc, 1);
                  umma_f16_pair(d_tmem, a_hi, b_hi, idesc, 1);
                }
              }
              umma_commit_pair(&empty_bar[stage], QUAD ? 0xF : 3);
              if (kb == kb1 - 1) umma_commit_pair(&tmem_full[as], (uint16_t)(3u << (2 * pair_in_cluster)));
            }
            __syncwarp();
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
          if (++as == NUM_ACC) { as = 0; aphase ^= 1; }
        }
      }
    }
  } else {
    tc_epilogue_role<BLOCK_N, NUM_ACC, true>(p, tmem_base, tmem_full, tmem_empty, warp, lane, num_kb);
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                // nobody leaves (or frees TMEM) while the peer still works
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, Cfg::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------
int encode_map_f16(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                   const cuuint32_t* box, const char* what, const cuuint32_t* elem_strides, int swizzle128 = 1);

// Resident clusters of CSIZE CTAs (2: one pair, 4: two pairs) the device can hold.
template <int BLOCK_N, bool QUAD>
static int pair_clusters(int num_sms, int* clusters) {
  using Cfg = PairCfg<BLOCK_N>;
  constexpr int CSIZE = QUAD ? 4 : 2;
  static PerDeviceInt cache;
  const int slot = current_device_slot();
  int max_clusters = cache.get(slot);
  if (max_clusters == 0) {
    TDN_CUDA_OK(cudaFuncSetAttribute(tc_conv_pair_kernel<BLOCK_N, QUAD>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     Cfg::SMEM_BYTES));
    // how many clusters the device can keep resident (GPCs whose usable SM count is not a multiple of the cluster size
    // leave SMs unused): a persistent kernel must not launch more, or the surplus runs as a second wave
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(num_sms / CSIZE * CSIZE, 1, 1);
    cfg.blockDim = dim3(TC_THREADS, 1, 1);
    cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = CSIZE; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, tc_conv_pair_kernel<BLOCK_N, QUAD>, &cfg) != cudaSuccess || n <= 0) {
      cudaGetLastError();
      n = QUAD ? 0 : num_sms / CSIZE;                // quad: unknown -> not used
    }
    max_clusters = n < num_sms / CSIZE ? n : num_sms / CSIZE;
    if (const char* e = getenv("TDNET_TC_PAIR_VERBOSE"))
      if (atoi(e)) fprintf(stderr, "[tdnet_b200] tc_conv_pair_kernel<%d>: %d resident clusters of %d CTAs\n", BLOCK_N, max_clusters, CSIZE);
    cache.set(slot, max_clusters > 0 ? max_clusters : -1);
  }
  *clusters = max_clusters > 0 ? max_clusters : 0;
  return TDN_OK;
}

int conv2d_tc_pair_clusters(int block_n, int num_sms, int* clusters) {
  return block_n == 256 ? pair_clusters<256, false>(num_sms, clusters) : pair_clusters<128, false>(num_sms, clusters);
}
int conv2d_tc_quad_clusters(int num_sms, int* clusters) { return pair_clusters<256, true>(num_sms, clusters); }

template <int BLOCK_N, bool QUAD>
static int launch_pair(const CUtensorMap& a_hi, const CUtensorMap& a_lo, const CUtensorMap& b_hi, const CUtensorMap& b_lo,
                       const TcParams& p, int num_sms, cudaStream_t stream) {
  using Cfg = PairCfg<BLOCK_N>;
  constexpr int PAIRS = QUAD ? 2 : 1;                // pairs per cluster; p.num_tiles counts pair tiles
  int max_clusters = 0, rc;
  if ((rc = pair_clusters<BLOCK_N, QUAD>(num_sms, &max_clusters))) return rc;
  TDN_REQUIRE(max_clusters > 0, TDN_ERR_UNSUPPORTED, "conv2d_tc_pair: no resident clusters for this variant");
  const int todo = (p.num_tiles - p.tile_begin + PAIRS - 1) / PAIRS;
  const int clusters = todo < max_clusters ? todo : max_clusters;
  TDN_CUDA_OK(tc_launch(tc_conv_pair_kernel<BLOCK_N, QUAD>, 2 * PAIRS * clusters, TC_THREADS, Cfg::SMEM_BYTES, stream,
                        todo <= 2 * clusters, a_hi, a_lo, b_hi, b_lo, p));
  TDN_LAUNCH_OK();
  return TDN_OK;
}

// Called by conv2d_tc() with the epilogue / output / residual fields of `p` filled in and the 128-pixel tile
// shape chosen; re-tiles into pair tiles of 2 x 128 pixels x block_n output channels.  max_pair_tiles > 0 limits
// the launch to the first pair tiles; first_pair_row > 0 starts it at that row of pair tiles (256 pixels each, all
// output-channel tiles) -- together they let the caller run the full rounds and the ragged last round as two launches.
// quad: clusters of two pairs that share the weight tile (block_n 256 only; tile indices then follow pair_tile_coords).
int conv2d_tc_pair(const tdn_tc_conv_desc* d, TcParams p, int block_n, int num_sms, int max_pair_tiles,
                   cudaStream_t stream, int first_pair_row, bool quad) {
  const tdn_tensor& in = d->in;
  const int cs = p.conv_stride;
  const int taps = d->kh * d->kw;
  TDN_REQUIRE(!d->weight_batched, TDN_ERR_UNSUPPORTED, "conv2d_tc_pair: per-image weights are not supported");
  p.n_tiles_n = ceil_div(d->cout, block_n);
  const long long m_tiles = (long long)in.n * p.tiles_h * p.tiles_w;
  const long long num_tiles = ((m_tiles + 1) / 2) * p.n_tiles_n;
  TDN_REQUIRE(num_tiles < (1ll << 30), TDN_ERR_UNSUPPORTED, "conv2d_tc_pair: too many tiles");
  p.num_tiles = (int)num_tiles;
  if (max_pair_tiles > 0 && max_pair_tiles < p.num_tiles) p.num_tiles = max_pair_tiles;
  p.tile_begin = first_pair_row * p.n_tiles_n;
  p.quad = 0;
  if (quad) {
    TDN_REQUIRE(block_n == 256 && max_pair_tiles <= 0 && first_pair_row == 0, TDN_ERR_INVALID,
                "conv2d_tc_pair: the quad variant runs whole launches of N = 256 tiles");
    p.quad = 1;
    p.num_tiles = (int)(((m_tiles + 3) / 4) * 2 * p.n_tiles_n);   // both pairs of a cluster walk the same number of tiles
  }

  CUtensorMap a_hi, a_lo, b_hi, b_lo;
  int rc;
  {
    cuuint64_t dims[4] = {(cuuint64_t)in.c, (cuuint64_t)in.w, (cuuint64_t)in.h, (cuuint64_t)in.n};
    cuuint64_t str[3] = {(cuuint64_t)in.stride_w * 2, (cuuint64_t)in.stride_h * 2, (cuuint64_t)in.stride_n * 2};
    cuuint32_t box[4] = {(cuuint32_t)TC_BLOCK_K, (cuuint32_t)(p.BW * cs), (cuuint32_t)(p.BH * cs), 1};
    cuuint32_t est[4] = {1, (cuuint32_t)cs, (cuuint32_t)cs, 1};
    if ((rc = encode_map_f16(&a_hi, in.data, 4, dims, str, box, "A.hi(pair)", est, 1))) return rc;
    if ((rc = encode_map_f16(&a_lo, in.data_lo, 4, dims, str, box, "A.lo(pair)", est, 1))) return rc;
  }
  {
    cuuint64_t dims[3] = {(cuuint64_t)taps * in.c, (cuuint64_t)d->cout, 1};
    cuuint64_t str[2] = {(cuuint64_t)d->weight_ld * 2, (cuuint64_t)d->weight_ld * 2 * (cuuint64_t)d->cout};
    cuuint32_t box[3] = {(cuuint32_t)TC_BLOCK_K, (cuuint32_t)(block_n / 2), 1};   // each CTA loads half of the rows
    if ((rc = encode_map_f16(&b_hi, d->weight_hi, 3, dims, str, box, "B.hi(pair)", nullptr, 1))) return rc;
    if ((rc = encode_map_f16(&b_lo, d->weight_lo, 3, dims, str, box, "B.lo(pair)", nullptr, 1))) return rc;
  }
  if (quad) return launch_pair<256, true>(a_hi, a_lo, b_hi, b_lo, p, num_sms, stream);
  if (block_n == 256) return launch_pair<256, false>(a_hi, a_lo, b_hi, b_lo, p, num_sms, stream);
  return launch_pair<128, false>(a_hi, a_lo, b_hi, b_lo, p, num_sms, stream);
}

}  // namespace tdn
