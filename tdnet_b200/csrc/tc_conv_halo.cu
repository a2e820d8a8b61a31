// tcgen05 3x3 convolution with ONE activation load per 64-channel block: the "halo" variant of tc_conv.cu
// for stride-1 3x3 convolutions with dilation 1 or 2 (ResNet layer1 / layer2 / layer3).
//
// tc_conv.cu fetches a fresh [128 pixel] x [64 channel] A tile for every filter tap, i.e. every input
// pixel crosses the L2->SM fabric nine times per output tile.  That makes the short-K, narrow-N layers
// L2-bandwidth-bound (measured: layer 1 moves 442 MB per 9.7 GFLOP launch and reaches only ~40 % of the
// tensor rate).  Here the CTA loads the (BH + 2d) x (BW + 2d) pixel REGION around its 16 x 8 pixel tile once
// per 64-channel block and all nine taps read it in place:
//   * the region lives in shared memory in the UMMA *no-swizzle* K-major layout: per 8-channel chunk a
//     dense [region pixels] x 16 B array (one TMA box with a 16-byte inner extent per chunk and plane);
//   * with BW = 8 a row of the output tile is exactly one 8-row core matrix (8 pixels x 16 B contiguous), so the
//     A descriptor of tap (ky, kx) is the same descriptor shifted by ((ky*d)*RW + kx*d) * 16 bytes:
//     SBO = RW * 16 B (next tile row), LBO = the chunk pitch (next 8 channels);
//   * B (weights) keeps the 128B-swizzled per-tap tiles and its own deeper ring.
// Activation traffic per output tile drops from 9 x 128 px to 180 px (d=1) / 240 px (d=2) per channel block.
// Arithmetic (exact mode, chunked fp32 accumulation, epilogue) is identical to tc_conv.cu.
#include "tc_common.cuh"

#include <stdlib.h>
#include <string.h>
#include <cuda.h>

namespace tdn {

constexpr int HL_BW = 8, HL_BH = 16;                 // output tile: 16 rows x 8 pixels = 128 GEMM rows
constexpr int HL_A_STAGES = 2;

template <int BLOCK_N>
struct HaloCfg {
  static constexpr int B_PLANE = BLOCK_N * TC_BLOCK_K * 2;
  static constexpr int B_STAGE = 2 * B_PLANE;
  static constexpr int B_STAGES = BLOCK_N == 64 ? 4 : 3;
  static constexpr int NUM_ACC = 512 / BLOCK_N;
  static constexpr int TMEM_COLS = 512;
};

struct HaloGeom {
  int RW, RH;            // region width / height in pixels
  int chunk_pitch;       // bytes between the 8-channel chunks of one plane (region bytes rounded up to 128)
  int a_stage;           // bytes of one A stage: 8 chunks x 2 planes
  int region_bytes;      // RW * RH * 16
};

template <int BLOCK_N>
__global__ void __launch_bounds__(TC_THREADS, 1)
tc_conv_halo_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                    const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
                    const TcParams p, const HaloGeom g) {
  using Cfg = HaloCfg<BLOCK_N>;
  constexpr int NUM_ACC = Cfg::NUM_ACC;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sB = smem;                                              // B ring first: keeps its tiles 1024-aligned
  uint8_t* sA = sB + Cfg::B_STAGES * Cfg::B_STAGE;                 // A ring (128-byte aligned chunks)
  uint64_t* a_full = reinterpret_cast<uint64_t*>(sA + HL_A_STAGES * g.a_stage);
  uint64_t* a_empty = a_full + HL_A_STAGES;
  uint64_t* b_full = a_empty + HL_A_STAGES;
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
    for (int s = 0; s < HL_A_STAGES; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
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
        const int x0 = tx * HL_BW - p.dil, y0 = ty * HL_BH - p.dil;   // top-left pixel of the region (may be < 0)
        for (int cb = 0; cb < cblocks; ++cb) {
          mbar_wait(&a_empty[as_], aph ^ 1);
          uint8_t* dst = sA + as_ * g.a_stage;
          mbar_expect_tx(&a_full[as_], 16 * g.region_bytes);
#pragma unroll 1
          for (int ch = 0; ch < 8; ++ch) {
            tma_load_4d(dst + ch * g.chunk_pitch, &tmA_hi, &a_full[as_], cb * TC_BLOCK_K + ch * 8, x0, y0, img);
            tma_load_4d(dst + (8 + ch) * g.chunk_pitch, &tmA_lo, &a_full[as_], cb * TC_BLOCK_K + ch * 8, x0, y0, img);
          }
          if (++as_ == HL_A_STAGES) { as_ = 0; aph ^= 1; }
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
    const uint32_t lbo = (uint32_t)(g.chunk_pitch >> 4), sbo = (uint32_t)g.RW;   // 16-byte units
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
          const uint32_t shift = (uint32_t)((ky * p.dil) * g.RW + kx * p.dil) * 16u;
          const bool last_in_chunk = (in_chunk == p.chunk_kb - 1) || (kb == num_kb - 1);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < TC_BLOCK_K / 16; ++k) {
              const uint32_t a_hi_addr = a_base + (2 * k) * g.chunk_pitch + shift;
              const uint32_t a_lo_addr = a_base + (8 + 2 * k) * g.chunk_pitch + shift;
              const uint64_t a_hi = umma_desc_k_noswz(a_hi_addr, lbo, sbo);
              const uint64_t a_lo = umma_desc_k_noswz(a_lo_addr, lbo, sbo);
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
        if (++as_ == HL_A_STAGES) { as_ = 0; aph ^= 1; }
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
static int launch_halo(const CUtensorMap& a_hi, const CUtensorMap& a_lo, const CUtensorMap& b_hi, const CUtensorMap& b_lo,
                       const TcParams& p, const HaloGeom& g, int num_sms, cudaStream_t stream) {
  using Cfg = HaloCfg<BLOCK_N>;
  const int smem = Cfg::B_STAGES * Cfg::B_STAGE + HL_A_STAGES * g.a_stage + 1024 + 512;
  TDN_REQUIRE(smem <= 232448, TDN_ERR_UNSUPPORTED, "conv2d_tc_halo: %d bytes of shared memory", smem);
  static PerDeviceFlag attr_set;
  const int slot = current_device_slot();
  if (!attr_set.is_set(slot)) {
    TDN_CUDA_OK(cudaFuncSetAttribute(tc_conv_halo_kernel<BLOCK_N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    attr_set.set(slot);
  }
  int grid = p.num_tiles < num_sms ? p.num_tiles : num_sms;
  TDN_CUDA_OK(tc_launch(tc_conv_halo_kernel<BLOCK_N>, grid, TC_THREADS, smem, stream, p.num_tiles <= 2 * grid, a_hi, a_lo, b_hi, b_lo, p, g));
  TDN_LAUNCH_OK();
  return TDN_OK;
}

// Called by conv2d_tc() for 3x3 / stride 1 / dilation <= 2 problems; `p` arrives with the epilogue, output and
// residual fields filled in and is re-tiled here.
int conv2d_tc_halo(const tdn_tc_conv_desc* d, TcParams p, int num_sms, int chunk_kb, cudaStream_t stream) {
  const tdn_tensor& in = d->in;
  p.BW = HL_BW; p.BH = HL_BH;
  p.tiles_h = ceil_div(in.h, HL_BH);
  p.tiles_w = ceil_div(in.w, HL_BW);
  const int block_n = d->cout <= 64 ? 64 : 128;
  p.n_tiles_n = ceil_div(d->cout, block_n);
  long long num_tiles = (long long)in.n * p.tiles_h * p.tiles_w * p.n_tiles_n;
  TDN_REQUIRE(num_tiles < (1ll << 31), TDN_ERR_UNSUPPORTED, "conv2d_tc_halo: too many tiles");
  p.num_tiles = (int)num_tiles;
  p.chunk_kb = chunk_kb;
  HaloGeom g;
  g.RW = HL_BW + 2 * d->dilation;
  g.RH = HL_BH + 2 * d->dilation;
  g.region_bytes = g.RW * g.RH * 16;
  g.chunk_pitch = (g.region_bytes + 127) / 128 * 128;
  g.a_stage = 16 * g.chunk_pitch;

  CUtensorMap a_hi, a_lo, b_hi, b_lo;
  int rc;
  {
    cuuint64_t dims[4] = {(cuuint64_t)in.c, (cuuint64_t)in.w, (cuuint64_t)in.h, (cuuint64_t)in.n};
    cuuint64_t str[3] = {(cuuint64_t)in.stride_w * 2, (cuuint64_t)in.stride_h * 2, (cuuint64_t)in.stride_n * 2};
    cuuint32_t box[4] = {8, (cuuint32_t)g.RW, (cuuint32_t)g.RH, 1};   // 8 channels = 16 bytes inner extent, no swizzle
    if ((rc = encode_map_f16(&a_hi, in.data, 4, dims, str, box, "A.hi(halo)", nullptr, 0))) return rc;
    if ((rc = encode_map_f16(&a_lo, in.data_lo, 4, dims, str, box, "A.lo(halo)", nullptr, 0))) return rc;
  }
  {
    cuuint64_t dims[3] = {(cuuint64_t)9 * in.c, (cuuint64_t)d->cout, 1};
    cuuint64_t str[2] = {(cuuint64_t)d->weight_ld * 2, (cuuint64_t)d->weight_ld * 2 * (cuuint64_t)d->cout};
    cuuint32_t box[3] = {(cuuint32_t)TC_BLOCK_K, (cuuint32_t)block_n, 1};
    if ((rc = encode_map_f16(&b_hi, d->weight_hi, 3, dims, str, box, "B.hi", nullptr, 1))) return rc;
    if ((rc = encode_map_f16(&b_lo, d->weight_lo, 3, dims, str, box, "B.lo", nullptr, 1))) return rc;
  }
  if (block_n == 64) return launch_halo<64>(a_hi, a_lo, b_hi, b_lo, p, g, num_sms, stream);
  return launch_halo<128>(a_hi, a_lo, b_hi, b_lo, p, g, num_sms, stream);
}

}  // namespace tdn
