// tcgen05 3x3 convolution on CTA pairs with ONE activation load per (channel block, filter ROW): the "band" variant of
// tc_conv_pair.cu for the wide dilated layers (ResNet layer 3 / 4: 3x3, stride 1, dilation <= 4, cout % 256 == 0).
//
// Hypothesis tested (TDN_TC_PAIR_BAND, an experiment kept as an explicit variant): tc_conv_pair.cu is paced by the bytes that
// enter the SM -- it needs 64 KB per K block and CTA (32 KB of activations + 32 KB of weights) for 1 536 cycles of MMAs and
// takes ~2 250.  The three taps of one filter row read the same pixels shifted by the dilation, so here a CTA loads, per
// 64-channel block and filter row ky, ONE band of 16 rows x 16 pixels around its 16 x 8 pixel tile (the halo in x only) and
// the taps kx = 0, 1, 2 read it in place.  Result: correct on the first run, 17 % fewer bytes into the SM, and the same time
// as tc_conv_pair.cu on every layer-3 / layer-4 shape (336.7 vs 333.4 us at 512 -> 512, dilation 4) -- it is not; DESIGN.md
// section 10, third session.
//   * the band lives in shared memory like every other operand: rows of 128 bytes (64 channels of a pixel), SWIZZLE_128B,
//     one TMA box per plane; it is 16 pixels wide, so a pixel's swizzle phase is its column & 7 in every band row;
//   * with BW = 8 an output-tile row is one 8-row swizzle group: the A descriptor of tap kx starts kx * d rows (128 B each)
//     into the band, SBO = 2048 B; the start is not 1024-byte aligned, which needs nothing (the hardware swizzles by
//     absolute address: tc_conv_halo_sw.cu);
//   * activation bytes per K block: 64 KB / 3 instead of 32 KB -- 53 KB per K block and CTA with the weights.
// A full halo region (all nine taps) would need 16 x 24 pixels x 2 planes = 96 KB per channel block at dilation 4, which does
// not fit twice next to a weight ring; bands of one filter row do: 2 x 64 KB of bands + 3 x 32 KB of weights.
// K blocks run in (channel block, ky, kx) order -- the order of tc_conv_halo.cu, bit-identical to it where both apply; against
// tc_conv.cu / tc_conv_pair.cu (tap-major order) the fp32 sums differ in the last bits.
// Protocol as in tc_conv_pair.cu (both CTAs run every role, the leader issues the MMAs, TMA bytes of both CTAs are counted on
// the leader's `full` barriers, commits are multicast to both CTAs), with separate rings for bands and weights.
#include "tc_common.cuh"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <cuda.h>

namespace tdn {

constexpr int PB_BW = 8, PB_BH = 16;                 // output tile per CTA: 16 rows x 8 pixels = 128 GEMM rows
constexpr int PB_PITCH = 16;                         // band width in pixels (8 + 2 d <= 16)
constexpr int PB_BAND_PLANE = PB_PITCH * PB_BH * 128;   // 32 KB: one plane of a band
constexpr int PB_A_STAGE = 2 * PB_BAND_PLANE;        // hi | lo
constexpr int PB_A_STAGES = 2;
constexpr int PB_N = 256;                            // pair tile: M 256 x N 256
constexpr int PB_B_HALF_PLANE = (PB_N / 2) * TC_BLOCK_K * 2;   // this CTA's 128 weight rows, one plane: 16 KB
constexpr int PB_B_STAGE = 2 * PB_B_HALF_PLANE;      // hi | lo
constexpr int PB_B_STAGES = 3;
constexpr int PB_NUM_ACC = 512 / PB_N;
constexpr int PB_SMEM_BYTES = PB_B_STAGES * PB_B_STAGE + PB_A_STAGES * PB_A_STAGE + 1024 + 512;
static_assert(PB_SMEM_BYTES <= 232448, "band kernel exceeds the 227 KB shared-memory limit");

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC_THREADS, 1)
tc_conv_pair_band_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                         const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
                         const TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  // identical layout in both CTAs: descriptors and barrier offsets are shared by the pair
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sB = smem;
  uint8_t* sA = sB + PB_B_STAGES * PB_B_STAGE;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(sA + PB_A_STAGES * PB_A_STAGE);
  uint64_t* a_empty = a_full + PB_A_STAGES;
  uint64_t* b_full = a_empty + PB_A_STAGES;
  uint64_t* b_empty = b_full + PB_B_STAGES;
  uint64_t* tmem_full = b_empty + PB_B_STAGES;
  uint64_t* tmem_empty = tmem_full + PB_NUM_ACC;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + PB_NUM_ACC);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int rank = (int)(blockIdx.x & 1);            // == %cluster_ctarank for cluster dims (2,1,1)
  const bool leader = rank == 0;
  const int pair_id = (int)(blockIdx.x >> 1);
  const int num_pairs = (int)(gridDim.x >> 1);
  const int cblocks = p.Cin / TC_BLOCK_K;
  const int num_kb = 9 * cblocks;

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tmA_hi);
    prefetch_tensormap(&tmA_lo);
    prefetch_tensormap(&tmB_hi);
    prefetch_tensormap(&tmB_lo);
    for (int s = 0; s < PB_A_STAGES; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < PB_B_STAGES; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
    for (int s = 0; s < PB_NUM_ACC; ++s) {
      mbar_init(&tmem_full[s], 1);                   // multicast commit
      mbar_init(&tmem_empty[s], 2 * TC_EPI_WARPS);   // leader only: one arrival per epilogue warp of both CTAs
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc_pair(tmem_ptr, 512);
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
      int as_ = 0, bs = 0;
      uint32_t aph = 0, bph = 0;
      for (int tile = p.tile_begin + pair_id; tile < p.num_tiles; tile += num_pairs) {
        const int nt = tile % p.n_tiles_n;
        int mt = 2 * (tile / p.n_tiles_n) + rank;    // this CTA's 128-pixel M tile (may lie past the last image:
        const int tx = mt % p.tiles_w;               //  TMA then zero-fills and the epilogue stores nothing)
        mt /= p.tiles_w;
        const int ty = mt % p.tiles_h;
        const int img = mt / p.tiles_h;
        const int x0 = tx * PB_BW - p.dil;           // left pixel of the band (may be < 0)
        for (int cb = 0; cb < cblocks; ++cb) {
          for (int ky = 0; ky < 3; ++ky) {
            const int y0 = ty * PB_BH + (ky - 1) * p.dil;
            mbar_wait(&a_empty[as_], aph ^ 1);
            uint8_t* da = sA + as_ * PB_A_STAGE;
            if (leader) mbar_expect_tx(&a_full[as_], 2 * PB_A_STAGE);
            tma_load_4d_pair(da, &tmA_hi, &a_full[as_], cb * TC_BLOCK_K, x0, y0, img);
            tma_load_4d_pair(da + PB_BAND_PLANE, &tmA_lo, &a_full[as_], cb * TC_BLOCK_K, x0, y0, img);
            if (++as_ == PB_A_STAGES) { as_ = 0; aph ^= 1; }
            for (int kx = 0; kx < 3; ++kx) {
              mbar_wait(&b_empty[bs], bph ^ 1);
              uint8_t* db = sB + bs * PB_B_STAGE;
              if (leader) mbar_expect_tx(&b_full[bs], 2 * PB_B_STAGE);
              const int kcol = (ky * 3 + kx) * p.Cin + cb * TC_BLOCK_K;
              const int brow = nt * PB_N + rank * (PB_N / 2);
              tma_load_3d_pair(db, &tmB_hi, &b_full[bs], kcol, brow, 0);
              tma_load_3d_pair(db + PB_B_HALF_PLANE, &tmB_lo, &b_full[bs], kcol, brow, 0);
              if (++bs == PB_B_STAGES) { bs = 0; bph ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ======================= MMA issuer (leader CTA only) =======================
    if (leader) {
      constexpr uint32_t idesc = umma_idesc_f16(2 * TC_BLOCK_M, PB_N);
      constexpr uint32_t sbo = (uint32_t)(PB_PITCH * 128) >> 4;      // next tile row = next 8-row group, 16-byte units
      int as_ = 0, bs = 0, acc = 0;
      uint32_t aph = 0, bph = 0, accph = 0;
      for (int tile = p.tile_begin + pair_id; tile < p.num_tiles; tile += num_pairs) {
        int kb = 0;                                                    // K block counter within the tile
        for (int cb = 0; cb < cblocks; ++cb) {
          for (int ky = 0; ky < 3; ++ky) {
            mbar_wait(&a_full[as_], aph);
            const uint32_t a_base = smem_u32(sA + as_ * PB_A_STAGE);
            for (int kx = 0; kx < 3; ++kx, ++kb) {
              const int in_chunk = kb % p.chunk_kb;
              if (in_chunk == 0) mbar_wait(&tmem_empty[acc], accph ^ 1);
              mbar_wait(&b_full[bs], bph);
              tc_fence_after();
              const uint32_t d_tmem = tmem_base + acc * PB_N;
              const uint32_t sb = smem_u32(sB + bs * PB_B_STAGE);
              const uint32_t shift = (uint32_t)(kx * p.dil) * 128u;
              const bool last_in_chunk = (in_chunk == p.chunk_kb - 1) || (kb == num_kb - 1);
              if (elect_one()) {
#pragma unroll
                for (int k = 0; k < TC_BLOCK_K / 16; ++k) {
                  const uint64_t a_hi = umma_desc_k_sw128_at(a_base + shift + k * 32, sbo);
                  const uint64_t a_lo = umma_desc_k_sw128_at(a_base + PB_BAND_PLANE + shift + k * 32, sbo);
                  const uint64_t b_hi = umma_desc_k_sw128(sb + k * 32);
                  const uint64_t b_lo = umma_desc_k_sw128(sb + PB_B_HALF_PLANE + k * 32);
                  if (p.fast) {
                    umma_f16_pair(d_tmem, a_hi, b_hi, idesc, (in_chunk | k) != 0);
                  } else {
                    umma_f16_pair(d_tmem, a_hi, b_lo, idesc, (in_chunk | k) != 0);
                    umma_f16_pair(d_tmem, a_lo, b_hi, idesc, 1);
                    umma_f16_pair(d_tmem, a_hi, b_hi, idesc, 1);
                  }
                }
                umma_commit_pair(&b_empty[bs], 3);
                if (kx == 2) umma_commit_pair(&a_empty[as_], 3);
                if (last_in_chunk) umma_commit_pair(&tmem_full[acc], 3);
              }
              __syncwarp();
              if (++bs == PB_B_STAGES) { bs = 0; bph ^= 1; }
              if (last_in_chunk) {
                if (++acc == PB_NUM_ACC) { acc = 0; accph ^= 1; }
              }
            }
            if (++as_ == PB_A_STAGES) { as_ = 0; aph ^= 1; }
          }
        }
      }
    }
  } else {
    tc_epilogue_role<PB_N, PB_NUM_ACC, true>(p, tmem_base, tmem_full, tmem_empty, warp, lane, num_kb);
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                // nobody leaves (or frees TMEM) while the peer still works
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------
int encode_map_f16(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                   const cuuint32_t* box, const char* what, const cuuint32_t* elem_strides, int swizzle128 = 1);
int conv2d_tc_pair_clusters(int block_n, int num_sms, int* clusters);

// Called by conv2d_tc() for 3x3 / stride 1 / dilation <= 4 / cout % 256 == 0 / shared weights; `p` arrives with the epilogue,
// output and residual fields filled in and is re-tiled here (16 x 8 pixel tiles per CTA).
int conv2d_tc_pair_band(const tdn_tc_conv_desc* d, TcParams p, int num_sms, cudaStream_t stream) {
  const tdn_tensor& in = d->in;
  TDN_REQUIRE(d->kh == 3 && d->kw == 3 && p.conv_stride == 1 && d->dilation >= 1 && d->dilation <= 4 && d->cout % PB_N == 0 &&
                  !d->weight_batched && in.c % TC_BLOCK_K == 0,
              TDN_ERR_UNSUPPORTED, "conv2d_tc_pair_band: needs a 3x3 stride-1 convolution, dilation <= 4, cout %% 256 == 0");
  p.BW = PB_BW; p.BH = PB_BH;
  p.tiles_h = ceil_div(in.h, PB_BH);
  p.tiles_w = ceil_div(in.w, PB_BW);
  p.n_tiles_n = d->cout / PB_N;
  const long long m_tiles = (long long)in.n * p.tiles_h * p.tiles_w;
  const long long num_tiles = ((m_tiles + 1) / 2) * p.n_tiles_n;
  TDN_REQUIRE(num_tiles < (1ll << 30), TDN_ERR_UNSUPPORTED, "conv2d_tc_pair_band: too many tiles");
  p.num_tiles = (int)num_tiles;
  p.tile_begin = 0;
  p.quad = 0;

  CUtensorMap a_hi, a_lo, b_hi, b_lo;
  int rc;
  {
    cuuint64_t dims[4] = {(cuuint64_t)in.c, (cuuint64_t)in.w, (cuuint64_t)in.h, (cuuint64_t)in.n};
    cuuint64_t str[3] = {(cuuint64_t)in.stride_w * 2, (cuuint64_t)in.stride_h * 2, (cuuint64_t)in.stride_n * 2};
    cuuint32_t box[4] = {(cuuint32_t)TC_BLOCK_K, (cuuint32_t)PB_PITCH, (cuuint32_t)PB_BH, 1};
    if ((rc = encode_map_f16(&a_hi, in.data, 4, dims, str, box, "A.hi(band)", nullptr, 1))) return rc;
    if ((rc = encode_map_f16(&a_lo, in.data_lo, 4, dims, str, box, "A.lo(band)", nullptr, 1))) return rc;
  }
  {
    cuuint64_t dims[3] = {(cuuint64_t)9 * in.c, (cuuint64_t)d->cout, 1};
    cuuint64_t str[2] = {(cuuint64_t)d->weight_ld * 2, (cuuint64_t)d->weight_ld * 2 * (cuuint64_t)d->cout};
    cuuint32_t box[3] = {(cuuint32_t)TC_BLOCK_K, (cuuint32_t)(PB_N / 2), 1};   // each CTA loads half of the rows
    if ((rc = encode_map_f16(&b_hi, d->weight_hi, 3, dims, str, box, "B.hi(band)", nullptr, 1))) return rc;
    if ((rc = encode_map_f16(&b_lo, d->weight_lo, 3, dims, str, box, "B.lo(band)", nullptr, 1))) return rc;
  }
  static PerDeviceFlag attr_set;
  const int slot = current_device_slot();
  if (!attr_set.is_set(slot)) {
    TDN_CUDA_OK(cudaFuncSetAttribute(tc_conv_pair_band_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PB_SMEM_BYTES));
    attr_set.set(slot);
  }
  int max_clusters = 0;
  if ((rc = conv2d_tc_pair_clusters(PB_N, num_sms, &max_clusters))) return rc;   // same cluster shape and block size
  const int clusters = p.num_tiles < max_clusters ? p.num_tiles : max_clusters;
  TDN_CUDA_OK(tc_launch(tc_conv_pair_band_kernel, 2 * clusters, TC_THREADS, PB_SMEM_BYTES, stream, p.num_tiles <= 2 * clusters,
                        a_hi, a_lo, b_hi, b_lo, p));
  TDN_LAUNCH_OK();
  return TDN_OK;
}

}  // namespace tdn
