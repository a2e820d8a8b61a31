// tcgen05 3x3 convolution with ONE activation load per 64-channel block, 128-byte-swizzled: the second "halo" variant.
//
// tc_conv_halo.cu keeps the (BH + 2d) x (BW + 2d) pixel region around a 16 x 8 pixel output tile in the UMMA NO-SWIZZLE
// layout so that a filter tap is a byte offset into it.  That saves the nine-fold activation traffic of tc_conv.cu, but its
// MMAs retire slowly (ncu, layer 2: 2.4 K cycles per K block for 768 cycles of tensor work, the MMA warp never waiting on a
// barrier): the no-swizzle operand -- 8-pixel core-matrix rows at a 160-byte pitch, shifted by 16 bytes per tap -- is never
// 128-byte aligned.  Here the region is stored the way every other operand of this library is: K-major rows of 128 bytes
// (64 channels of one pixel), SWIZZLE_128B, written by ONE TMA box per plane:
//   * the region is 16 pixels wide in shared memory (pitch 16 px = 2048 B; only 8 + 2d of them are used) and 16 + 2d rows
//     high: pixel (ry, rx) is row ry * 16 + rx, so its swizzle phase is rx & 7 in EVERY region row;
//   * with BW = 8 an output-tile row is exactly one 8-row swizzle group: the A descriptor of tap (ky, kx) starts at row
//     (ky d) * 16 + kx d, SBO = 2048 B (next tile row); the start is 128-byte but not 1024-byte aligned.  Measured on
//     B200: the hardware takes the swizzle phase of every row from its ABSOLUTE shared-memory address, so the descriptor's
//     base-offset field must stay 0 (setting it to (address >> 7) & 7, as one reading of CuTe's comment suggests, gives
//     wrong results; with 0 the output is bit-identical to tc_conv_halo.cu);
//   * B (weights) keeps the per-tap 128B-swizzled tiles and its own ring.
// Activation traffic per output tile and channel block: 16 x (16 + 2d) = 288 / 320 px instead of 9 x 128.
// Result (tools/halo_probe.py, profiles/r02_halo_probe.txt): NOT faster than the no-swizzle region -- layer 2 39.2 vs 39.7 us,
// dilation 2 39.2 vs 42.2 us, layer 1 54.2 vs 52.5 us (per-tap kernel: 46.5) -- so the operand layout was not what slows the
// halo kernel's K blocks (they take 2-3x their tensor time in every N <= 128 kernel of this library, whatever the layout or
// the operand source: DESIGN.md section 10).  Kept as an explicit variant (TDN_TC_HALO_SW), never picked by TDN_TC_AUTO.  Arithmetic (exact mode, chunked fp32 accumulation, epilogue) is identical to
// tc_conv.cu: same products in the same order, bit-identical output.
#include "tc_common.cuh"

#include <stdlib.h>
#include <string.h>
#include <cuda.h>

namespace tdn {

constexpr int HS_BW = 8, HS_BH = 16;                 // output tile: 16 rows x 8 pixels = 128 GEMM rows
constexpr int HS_PITCH = 16;                         // region pitch in pixels (one swizzle phase per region column)
constexpr int HS_A_STAGES = 2;

template <int BLOCK_N>
struct HaloSwCfg {
  static constexpr int B_PLANE = BLOCK_N * TC_BLOCK_K * 2;
  static constexpr int B_STAGE = 2 * B_PLANE;
  static constexpr int B_STAGES = BLOCK_N == 64 ? 4 : 2;            // 64 KB either way: the two A stages take 144-160 KB
  static constexpr int NUM_ACC = 512 / BLOCK_N;
  static constexpr int TMEM_COLS = 512;
};

struct HaloSwGeom {
  int RH;                // region height in pixels (16 + 2 d)
  int plane;             // bytes of one plane of the region: RH * 16 px * 128 B
  int a_stage;           // bytes of one A stage: hi plane | lo plane
};

template <int BLOCK_N>
__global__ void __launch_bounds__(TC_THREADS, 1)
tc_conv_halo_sw_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                    const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
                    const TcParams p, const HaloSwGeom g) {
  using Cfg = HaloSwCfg<BLOCK_N>;
  constexpr int NUM_ACC = Cfg::NUM_ACC;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sB = smem;                                              // B ring first: keeps its tiles 1024-aligned
  uint8_t* sA = sB + Cfg::B_STAGES * Cfg::B_STAGE;                 // A ring (planes are multiples of 2048 B: 1024-aligned)
  uint64_t* a_full = reinterpret_cast<uint64_t*>(sA + HS_A_STAGES * g.a_stage);
  uint64_t* a_empty = a_full + HS_A_STAGES;
  uint64_t* b_full = a_empty + HS_A_STAGES;
  uint64_t* b_empty = b_full + Cfg::B_STAGES;
  uint64_t* tmem_full = b_empty + Cfg::B_STAGES;
  uint64_t* tmem_empty = tmem_full + NUM_ACC;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + NUM_ACC);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int cblocks = p.Cin / TC_BLOCK_K;
  const int num_kb = 9 * cblocks;                                  // (channel block, tap) pairs per tile

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tmA_hi); prefetch_tensormap(&tmA_lo);
    prefetch_tensormap(&tmB_hi); prefetch_tensormap(&tmB_lo);
    for (int s = 0; s < HS_A_STAGES; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < Cfg::B_STAGES; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
    for (int s = 0; s < NUM_ACC; ++s) { mbar_init(&tmem_full[s], 1); mbar_init(&tmem_empty[s], 32 * TC_EPI_WARPS); }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  tc_pdl_sync();

  if (warp == 0) {
    // ======================= TMA producer =======================
    if (lane == 0) {
      int as_ = 0, bs = 0;
      uint32_t aph = 0, bph = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        const int nt = tile % p.n_tiles_n;
        int mt = tile / p.n_tiles_n;
        const int tx = mt % p.tiles_w;
        mt /= p.tiles_w;
        const int ty = mt % p.tiles_h;
        const int img = mt / p.tiles_h;
        const int x0 = tx * HS_BW - p.dil, y0 = ty * HS_BH - p.dil;   // top-left pixel of the region (may be < 0)
        for (int cb = 0; cb < cblocks; ++cb) {
          mbar_wait(&a_empty[as_], aph ^ 1);
          uint8_t* dst = sA + as_ * g.a_stage;
          mbar_expect_tx(&a_full[as_], 2 * g.plane);
          tma_load_4d(dst, &tmA_hi, &a_full[as_], cb * TC_BLOCK_K, x0, y0, img);
          tma_load_4d(dst + g.plane, &tmA_lo, &a_full[as_], cb * TC_BLOCK_K, x0, y0, img);
          if (++as_ == HS_A_STAGES) { as_ = 0; aph ^= 1; }
          for (int tap = 0; tap < 9; ++tap) {
            mbar_wait(&b_empty[bs], bph ^ 1);
            uint8_t* db = sB + bs * Cfg::B_STAGE;
            mbar_expect_tx(&b_full[bs], Cfg::B_STAGE);
            const int kcol = tap * p.Cin + cb * TC_BLOCK_K;
            tma_load_3d(db, &tmB_hi, &b_full[bs], kcol, nt * BLOCK_N, 0);
            tma_load_3d(db + Cfg::B_PLANE, &tmB_lo, &b_full[bs], kcol, nt * BLOCK_N, 0);
            if (++bs == Cfg::B_STAGES) { bs = 0; bph ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ======================= MMA issuer (warp-uniform loop, elected lane issues) =======================
    constexpr uint32_t idesc = umma_idesc_f16(TC_BLOCK_M, BLOCK_N);
    constexpr uint32_t sbo = (uint32_t)(HS_PITCH * 128) >> 4;      // next tile row = next 8-row group, 16-byte units
    int as_ = 0, bs = 0, acc = 0;
    uint32_t aph = 0, bph = 0, accph = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      int kb = 0;                                                    // (cb, tap) counter within the tile
      for (int cb = 0; cb < cblocks; ++cb) {
        mbar_wait(&a_full[as_], aph);
        const uint32_t a_base = smem_u32(sA + as_ * g.a_stage);
        for (int tap = 0; tap < 9; ++tap, ++kb) {
          const int in_chunk = kb % p.chunk_kb;
          if (in_chunk == 0) mbar_wait(&tmem_empty[acc], accph ^ 1);
          mbar_wait(&b_full[bs], bph);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
          const uint32_t sb = smem_u32(sB + bs * Cfg::B_STAGE);
          const int ky = tap / 3, kx = tap - ky * 3;
          const uint32_t shift = (uint32_t)((ky * p.dil) * HS_PITCH + kx * p.dil) * 128u;
          const bool last_in_chunk = (in_chunk == p.chunk_kb - 1) || (kb == num_kb - 1);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < TC_BLOCK_K / 16; ++k) {
              const uint64_t a_hi = umma_desc_k_sw128_at(a_base + shift + k * 32, sbo);
              const uint64_t a_lo = umma_desc_k_sw128_at(a_base + g.plane + shift + k * 32, sbo);
              const uint64_t b_hi = umma_desc_k_sw128(sb + k * 32);
              const uint64_t b_lo = umma_desc_k_sw128(sb + Cfg::B_PLANE + k * 32);
              if (p.fast) {
                umma_f16(d_tmem, a_hi, b_hi, idesc, (in_chunk | k) != 0);
              } else {
                umma_f16(d_tmem, a_hi, b_lo, idesc, (in_chunk | k) != 0);
                umma_f16(d_tmem, a_lo, b_hi, idesc, 1);
                umma_f16(d_tmem, a_hi, b_hi, idesc, 1);
              }
            }
            umma_commit(&b_empty[bs]);
            if (tap == 8) umma_commit(&a_empty[as_]);
            if (last_in_chunk) umma_commit(&tmem_full[acc]);
          }
          __syncwarp();
          if (++bs == Cfg::B_STAGES) { bs = 0; bph ^= 1; }
          if (last_in_chunk) {
            if (++acc == NUM_ACC) { acc = 0; accph ^= 1; }
          }
        }
        if (++as_ == HS_A_STAGES) { as_ = 0; aph ^= 1; }
      }
    }
  } else {
    tc_epilogue_role<BLOCK_N, NUM_ACC>(p, tmem_base, tmem_full, tmem_empty, warp, lane, num_kb);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------
int encode_map_f16(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                   const cuuint32_t* box, const char* what, const cuuint32_t* elem_strides, int swizzle128 = 1);

template <int BLOCK_N>
static int launch_halo_sw(const CUtensorMap& a_hi, const CUtensorMap& a_lo, const CUtensorMap& b_hi, const CUtensorMap& b_lo,
                       const TcParams& p, const HaloSwGeom& g, int num_sms, cudaStream_t stream) {
  using Cfg = HaloSwCfg<BLOCK_N>;
  const int smem = Cfg::B_STAGES * Cfg::B_STAGE + HS_A_STAGES * g.a_stage + 1024 + 512;
  TDN_REQUIRE(smem <= 232448, TDN_ERR_UNSUPPORTED, "conv2d_tc_halo_sw: %d bytes of shared memory", smem);
  static PerDeviceFlag attr_set;
  const int slot = current_device_slot();
  if (!attr_set.is_set(slot)) {
    TDN_CUDA_OK(cudaFuncSetAttribute(tc_conv_halo_sw_kernel<BLOCK_N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    attr_set.set(slot);
  }
  int grid = p.num_tiles < num_sms ? p.num_tiles : num_sms;
  TDN_CUDA_OK(tc_launch(tc_conv_halo_sw_kernel<BLOCK_N>, grid, TC_THREADS, smem, stream, p.num_tiles <= 2 * grid, a_hi, a_lo, b_hi, b_lo, p, g));
  TDN_LAUNCH_OK();
  return TDN_OK;
}

// Called by conv2d_tc() for 3x3 / stride 1 / dilation <= 2 problems with cin % 64 == 0; `p` arrives with the epilogue,
// output and residual fields filled in and is re-tiled here.
int conv2d_tc_halo_sw(const tdn_tc_conv_desc* d, TcParams p, int num_sms, int chunk_kb, cudaStream_t stream) {
  const tdn_tensor& in = d->in;
  p.BW = HS_BW; p.BH = HS_BH;
  p.tiles_h = ceil_div(in.h, HS_BH);
  p.tiles_w = ceil_div(in.w, HS_BW);
  const int block_n = d->cout <= 64 ? 64 : 128;
  p.n_tiles_n = ceil_div(d->cout, block_n);
  long long num_tiles = (long long)in.n * p.tiles_h * p.tiles_w * p.n_tiles_n;
  TDN_REQUIRE(num_tiles < (1ll << 31), TDN_ERR_UNSUPPORTED, "conv2d_tc_halo_sw: too many tiles");
  p.num_tiles = (int)num_tiles;
  p.chunk_kb = chunk_kb;
  HaloSwGeom g;
  g.RH = HS_BH + 2 * d->dilation;
  g.plane = g.RH * HS_PITCH * 128;
  g.a_stage = 2 * g.plane;

  CUtensorMap a_hi, a_lo, b_hi, b_lo;
  int rc;
  {
    cuuint64_t dims[4] = {(cuuint64_t)in.c, (cuuint64_t)in.w, (cuuint64_t)in.h, (cuuint64_t)in.n};
    cuuint64_t str[3] = {(cuuint64_t)in.stride_w * 2, (cuuint64_t)in.stride_h * 2, (cuuint64_t)in.stride_n * 2};
    cuuint32_t box[4] = {(cuuint32_t)TC_BLOCK_K, (cuuint32_t)HS_PITCH, (cuuint32_t)g.RH, 1};   // 128-byte rows, swizzled
    if ((rc = encode_map_f16(&a_hi, in.data, 4, dims, str, box, "A.hi(halo sw)", nullptr, 1))) return rc;
    if ((rc = encode_map_f16(&a_lo, in.data_lo, 4, dims, str, box, "A.lo(halo sw)", nullptr, 1))) return rc;
  }
  {
    cuuint64_t dims[3] = {(cuuint64_t)9 * in.c, (cuuint64_t)d->cout, 1};
    cuuint64_t str[2] = {(cuuint64_t)d->weight_ld * 2, (cuuint64_t)d->weight_ld * 2 * (cuuint64_t)d->cout};
    cuuint32_t box[3] = {(cuuint32_t)TC_BLOCK_K, (cuuint32_t)block_n, 1};
    if ((rc = encode_map_f16(&b_hi, d->weight_hi, 3, dims, str, box, "B.hi", nullptr, 1))) return rc;
    if ((rc = encode_map_f16(&b_lo, d->weight_lo, 3, dims, str, box, "B.lo", nullptr, 1))) return rc;
  }
  if (block_n == 64) return launch_halo_sw<64>(a_hi, a_lo, b_hi, b_lo, p, g, num_sms, stream);
  return launch_halo_sw<128>(a_hi, a_lo, b_hi, b_lo, p, g, num_sms, stream);
}

}  // namespace tdn
