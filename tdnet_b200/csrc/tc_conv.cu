// tcgen05 implicit-GEMM convolution / GEMM for sm_100a, fp32-faithful ("exact mode").
//
//   out[pixel, co] = act( (sum_{tap, ci} A[pixel + tap*dil, ci] * W[co, tap, ci]) * scale[co] + bias + residual )
//
// Operands are SPLIT16: every fp32 value x is carried as two fp16 planes, hi = fp16(x) and
// lo = fp16(x - hi), so hi + lo reproduces x to >= 22 significant bits.  Three tensor-core products
// per K step -- Ah*Bl, Al*Bh, Ah*Bh -- accumulate in ONE fp32 TMEM accumulator (fp16 x fp16 products
// are exact in fp32; the dropped Al*Bl term is below 2^-22 relative).  This is what makes the
// tensor-core path match the reference's fp32 arithmetic instead of being a bf16/tf32 approximation.
//
// Structure (one persistent CTA per SM, 320 threads):
//   warp 0      TMA producer: per K block (one filter tap x 64 input channels) four bulk-tensor loads
//               -- A hi/lo as a [BH x BW pixels] x 64ch box of the NHWC map, shifted by the tap offset
//               (out-of-bounds rows/cols are zero-filled by TMA = the convolution padding), B hi/lo as
//               [BLOCK_N couts] x 64 -- into a STAGES-deep shared-memory ring (128B swizzle).
//   warp 1      MMA issuer: one elected lane issues 12 tcgen05.mma (4 K16 steps x 3 products) per
//               stage into a double-buffered TMEM accumulator; tcgen05.commit releases the stage.
//   warps 2-9   epilogue: tcgen05.ld the 128 x BLOCK_N fp32 tile (one pixel row x half the columns per
//               thread), apply scale/bias/residual/activation, re-split to hi/lo (and/or write fp32),
//               16-byte stores.
// Accumulation is CHUNKED: the tensor core adds into its fp32 TMEM accumulator with truncation (measured
// on B200: -3.8e-5 mean relative error after 864 chained MMAs on positive data), so a TMEM accumulator
// only ever holds `chunk_kb` K blocks (default 4 = 48 MMAs); the epilogue warps pull each finished chunk
// out of TMEM and add it to per-thread fp32 registers with round-to-nearest.  The 512 TMEM columns form
// a ring of 4 (BLOCK_N=128) or 8 (BLOCK_N=64) chunk accumulators, so the MMA warp keeps issuing while
// earlier chunks are drained and while the scale/bias/store phase of the previous tile runs.
#include "tc_common.cuh"

#include <stdlib.h>
#include <string.h>
#include <cuda.h>  // CUtensorMap types only; the encode entry point is resolved at run time

namespace tdn {

// TSA ("A through tensor memory"; TDN_TC_BASE_TS, an experiment kept as an explicit variant): the MMA warp copies the A tile of
// every K block from shared memory into one of two 64-column TMEM buffers (tcgen05.cp.128x256b, 8 slabs of 128 rows x 32
// bytes: hi plane then lo plane, the source named by the same SWIZZLE_128B descriptor an MMA would use) and the twelve
// exact-mode MMAs take A from there (tcgen05.mma [d], [a_tmem], b_desc); copies and MMAs of one thread execute in issue order,
// so nothing else synchronises them.  Shared memory is then read once per K block for A (32 KB) instead of twelve times 4 KB;
// the ring of chunk accumulators shrinks from 512 to 384 columns.  Same products in the same order: bit-identical.
// Hypothesis tested: the N <= 128 kernels (K blocks of 2-3x their tensor time) wait for shared-memory operand reads.
// Result: 12-16 % SLOWER on every shape -- they do not; DESIGN.md section 10, third session.
template <int BLOCK_N, bool TSA = false>
struct TcCfg {
  static constexpr int B_PLANE = BLOCK_N * TC_BLOCK_K * 2;
  static constexpr int STAGE_BYTES = 2 * TC_A_PLANE + 2 * B_PLANE;
  static constexpr int STAGES = (200 * 1024) / STAGE_BYTES > 6 ? 6 : (200 * 1024) / STAGE_BYTES;
  // All 512 TMEM columns as a ring of chunk accumulators (4 x 128 or 8 x 64 columns): the MMA warp can
  // run several chunks ahead of the warps that drain them, so the per-tile store phase of the epilogue
  // (scale/bias/residual/split/store) overlaps the MMAs of the next tile instead of stalling them.
  static constexpr int NUM_ACC = (TSA ? 384 : 512) / BLOCK_N;
  static constexpr int A_TMEM_COL = 384;             // TSA: A buffer b = columns [384 + 64 b, 448 + 64 b): hi pairs | lo pairs
  static constexpr int TMEM_COLS = 512;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 512 /*barriers*/;
};

template <int BLOCK_N, bool TSA>
__global__ void __launch_bounds__(TC_THREADS, 1)
tc_conv_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
               const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
               const TcParams p) {
  using Cfg = TcCfg<BLOCK_N, TSA>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  constexpr int NUM_ACC = Cfg::NUM_ACC;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + NUM_ACC;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + NUM_ACC);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int kc_per_tap = p.Cin / TC_BLOCK_K;
  const int num_kb = p.taps_h * p.taps_w * kc_per_tap;

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tmA_hi);
    prefetch_tensormap(&tmA_lo);
    prefetch_tensormap(&tmB_hi);
    prefetch_tensormap(&tmB_lo);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < NUM_ACC; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], 32 * TC_EPI_WARPS);
    }
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
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x + p.tile_begin; tile < p.num_tiles; tile += gridDim.x) {
        const int nt = tile % p.n_tiles_n;
        int mt = tile / p.n_tiles_n;
        const int tx = mt % p.tiles_w;
        mt /= p.tiles_w;
        const int ty = mt % p.tiles_h;
        const int img = mt / p.tiles_h;
        for (int kb = 0; kb < num_kb; ++kb) {
          const int tap = kb / kc_per_tap;
          const int kc = kb - tap * kc_per_tap;
          const int ky = tap / p.taps_w;
          const int kx = tap - ky * p.taps_w;
          // top-left input pixel of the tap's box; with conv_stride 2 the tensor map picks every second
          // pixel (TMA elementStrides), so the box still lands as BH x BW pixel rows in shared memory
          const int x0 = tx * p.BW * p.conv_stride + (kx - (p.taps_w - 1) / 2) * p.dil;
          const int y0 = ty * p.BH * p.conv_stride + (ky - (p.taps_h - 1) / 2) * p.dil;
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
          uint8_t* sb = sa + 2 * TC_A_PLANE;
          mbar_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
          tma_load_4d(sa, &tmA_hi, &full_bar[stage], kc * TC_BLOCK_K, x0, y0, img);
          tma_load_4d(sa + TC_A_PLANE, &tmA_lo, &full_bar[stage], kc * TC_BLOCK_K, x0, y0, img);
          const int kcol = tap * p.Cin + kc * TC_BLOCK_K;
          const int bz = p.w_batched ? img : 0;
          tma_load_3d(sb, &tmB_hi, &full_bar[stage], kcol, nt * BLOCK_N, bz);
          tma_load_3d(sb + Cfg::B_PLANE, &tmB_lo, &full_bar[stage], kcol, nt * BLOCK_N, bz);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ======================= MMA issuer =======================
    // All 32 lanes run the loop and the barrier waits (keeps stage / phase / descriptors warp-uniform, i.e.
    // in uniform registers); one elected lane issues the tcgen05 instructions.
    constexpr uint32_t idesc = umma_idesc_f16(TC_BLOCK_M, BLOCK_N);
    int stage = 0;
    uint32_t phase = 0;
    int as = 0;
    uint32_t aphase = 0;
    uint32_t kbc = 0;                                  // K blocks issued so far (TSA: A buffer kbc & 1)
    for (int tile = blockIdx.x + p.tile_begin; tile < p.num_tiles; tile += gridDim.x) {
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
            if (TSA) {
              const uint32_t a_t = tmem_base + Cfg::A_TMEM_COL + (kbc & 1) * 64;
#pragma unroll
              for (int k = 0; k < TC_BLOCK_K / 16; ++k) {
                tmem_cp_128x256b(a_t + k * 8, umma_desc_k_sw128(sa + k * 32));
                tmem_cp_128x256b(a_t + 32 + k * 8, umma_desc_k_sw128(sa + TC_A_PLANE + k * 32));
              }
#pragma unroll
              for (int k = 0; k < TC_BLOCK_K / 16; ++k) {
                const uint32_t a_hi = a_t + k * 8, a_lo = a_hi + 32;
                const uint64_t b_hi = umma_desc_k_sw128(sb + k * 32);
                const uint64_t b_lo = umma_desc_k_sw128(sb + Cfg::B_PLANE + k * 32);
                if (p.fast) {
                  umma_f16_ts(d_tmem, a_hi, b_hi, idesc, ((kb - kb0) | k) != 0);
                } else {
                  umma_f16_ts(d_tmem, a_hi, b_lo, idesc, ((kb - kb0) | k) != 0);
                  umma_f16_ts(d_tmem, a_lo, b_hi, idesc, 1);
                  umma_f16_ts(d_tmem, a_hi, b_hi, idesc, 1);
                }
              }
            } else {
#pragma unroll
            for (int k = 0; k < TC_BLOCK_K / 16; ++k) {
              const uint64_t a_hi = umma_desc_k_sw128(sa + k * 32);
              const uint64_t a_lo = umma_desc_k_sw128(sa + TC_A_PLANE + k * 32);
              const uint64_t b_hi = umma_desc_k_sw128(sb + k * 32);
              const uint64_t b_lo = umma_desc_k_sw128(sb + Cfg::B_PLANE + k * 32);
              if (p.fast) {
                umma_f16(d_tmem, a_hi, b_hi, idesc, ((kb - kb0) | k) != 0);
              } else {
                umma_f16(d_tmem, a_hi, b_lo, idesc, ((kb - kb0) | k) != 0);
                umma_f16(d_tmem, a_lo, b_hi, idesc, 1);
                umma_f16(d_tmem, a_hi, b_hi, idesc, 1);
              }
            }
            }
            umma_commit(&empty_bar[stage]);
            if (kb == kb1 - 1) umma_commit(&tmem_full[as]);
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
          ++kbc;
        }
        if (++as == NUM_ACC) { as = 0; aphase ^= 1; }
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
// Host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)ptr;
  }
  return fn;
}

int encode_map_f16(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims,
                   const cuuint64_t* strides_bytes, const cuuint32_t* box, const char* what,
                   const cuuint32_t* elem_strides, int swizzle128) {
  EncodeTiledFn fn = get_encode_fn();
  TDN_REQUIRE(fn != nullptr, TDN_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  if (elem_strides)
    for (int i = 0; i < rank; ++i) estr[i] = elem_strides[i];
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, rank, const_cast<void*>(base), dims, strides_bytes, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  TDN_REQUIRE(r == CUDA_SUCCESS, TDN_ERR_CUDA, "cuTensorMapEncodeTiled(%s) failed with CUresult %d", what, (int)r);
  return TDN_OK;
}

static void pick_tile(int H, int W, int* BH, int* BW) {
  // 128 pixels as BH x BW; minimise the padded area, prefer wide tiles (longer contiguous runs) on ties.
  const int cand[5][2] = {{1, 128}, {2, 64}, {4, 32}, {8, 16}, {16, 8}};
  long long best = -1;
  for (int i = 0; i < 5; ++i) {
    int bh = cand[i][0], bw = cand[i][1];
    long long area = (long long)ceil_div(H, bh) * bh * ceil_div(W, bw) * bw;
    if (best < 0 || area < best) { best = area; *BH = bh; *BW = bw; }
  }
}


template <int BLOCK_N, bool TSA>
static int launch_tc(const CUtensorMap& a_hi, const CUtensorMap& a_lo, const CUtensorMap& b_hi,
                     const CUtensorMap& b_lo, const TcParams& p, cudaStream_t stream) {
  using Cfg = TcCfg<BLOCK_N, TSA>;
  static PerDeviceFlag attr_set;
  const int slot = current_device_slot();
  if (!attr_set.is_set(slot)) {
    TDN_CUDA_OK(cudaFuncSetAttribute(tc_conv_kernel<BLOCK_N, TSA>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     Cfg::SMEM_BYTES));
    attr_set.set(slot);
  }
  const int g_num_sms = device_sm_count();
  TDN_REQUIRE(g_num_sms > 0, TDN_ERR_CUDA, "conv2d_tc: cannot query the SM count");
  const int todo = p.num_tiles - p.tile_begin;
  int grid = todo < g_num_sms ? todo : g_num_sms;
  TDN_CUDA_OK(tc_launch(tc_conv_kernel<BLOCK_N, TSA>, grid, TC_THREADS, Cfg::SMEM_BYTES, stream, todo <= 2 * grid, a_hi, a_lo, b_hi, b_lo, p));
  return TDN_OK;
}

int conv2d_tc_halo(const tdn_tc_conv_desc* d, TcParams p, int num_sms, int chunk_kb, cudaStream_t stream);
int conv2d_tc_halo_sw(const tdn_tc_conv_desc* d, TcParams p, int num_sms, int chunk_kb, cudaStream_t stream);
int conv2d_tc_pair(const tdn_tc_conv_desc* d, TcParams p, int block_n, int num_sms, int max_pair_tiles,
                   cudaStream_t stream, int first_pair_row = 0, bool quad = false);
int conv2d_tc_pair_clusters(int block_n, int num_sms, int* clusters);
int conv2d_tc_quad_clusters(int num_sms, int* clusters);
int conv2d_tc_pair_band(const tdn_tc_conv_desc* d, TcParams p, int num_sms, cudaStream_t stream);

int conv2d_tc(const tdn_tc_conv_desc* d, cudaStream_t stream) {
  const tdn_tensor& in = d->in;
  const tdn_tensor& out = d->out;
  TDN_REQUIRE(in.dtype == TDN_SPLIT16 && in.data && in.data_lo, TDN_ERR_INVALID, "conv2d_tc: input must be SPLIT16");
  TDN_REQUIRE(d->weight_hi && d->weight_lo, TDN_ERR_INVALID, "conv2d_tc: null weights");
  TDN_REQUIRE(in.c % TC_BLOCK_K == 0, TDN_ERR_UNSUPPORTED, "conv2d_tc: cin=%d must be a multiple of 64", in.c);
  TDN_REQUIRE(d->kh % 2 == 1 && d->kw % 2 == 1 && d->kh * d->kw <= 49, TDN_ERR_UNSUPPORTED, "conv2d_tc: odd kernels only");
  const int cs = d->stride <= 1 ? 1 : d->stride;
  TDN_REQUIRE(cs == 1 || cs == 2, TDN_ERR_UNSUPPORTED, "conv2d_tc: stride must be 1 or 2");
  TDN_REQUIRE(out.n == in.n && out.h == (in.h - 1) / cs + 1 && out.w == (in.w - 1) / cs + 1 && out.c == d->cout,
              TDN_ERR_INVALID, "conv2d_tc: 'same'-padded convolution expects out dims == ceil(in dims / stride)");
  TDN_REQUIRE(aligned16(in.data) && aligned16(in.data_lo) && in.stride_w % 8 == 0 && in.stride_h % 8 == 0 &&
                  in.stride_n % 8 == 0, TDN_ERR_INVALID, "conv2d_tc: input planes must be 16-byte aligned");
  TDN_REQUIRE((!d->scale || aligned16(d->scale)) && (!d->bias || aligned16(d->bias)), TDN_ERR_INVALID,
              "conv2d_tc: scale/bias must be 16-byte aligned");
  TDN_REQUIRE(aligned16(d->weight_hi) && aligned16(d->weight_lo) && d->weight_ld % 8 == 0 &&
                  d->weight_batch_stride % 8 == 0, TDN_ERR_INVALID, "conv2d_tc: weights must be 16-byte aligned");
  const bool out16 = out.dtype == TDN_SPLIT16;
  TDN_REQUIRE(out.data != nullptr && (out16 ? out.data_lo != nullptr : true), TDN_ERR_INVALID, "conv2d_tc: null output");
  const int taps = d->kh * d->kw;
  TDN_REQUIRE(d->weight_ld >= (long long)taps * in.c, TDN_ERR_INVALID, "conv2d_tc: weight_ld < K");

  TcParams p;
  memset(&p, 0, sizeof(p));
  p.n_img = in.n; p.Ho = out.h; p.Wo = out.w;
  pick_tile(out.h, out.w, &p.BH, &p.BW);
  p.tiles_h = ceil_div(out.h, p.BH);
  p.tiles_w = ceil_div(out.w, p.BW);
  p.conv_stride = cs;
  p.Cout = d->cout; p.Cin = in.c;
  p.taps_h = d->kh; p.taps_w = d->kw; p.dil = d->dilation;
  // N tile: 128 columns unless the problem is so small that 128-wide tiles would leave the 148 SMs with fewer
  // than two waves of work AND the K loop is short (measured: the fc GEMMs with 8 K blocks gain 25 % from
  // N = 64, the head conv with 72 K blocks loses 35 % because the A tile is then fetched twice as often).
  const int g_num_sms = device_sm_count();
  TDN_REQUIRE(g_num_sms > 0, TDN_ERR_CUDA, "conv2d_tc: cannot query the SM count");
  const long long tiles128 = (long long)in.n * p.tiles_h * p.tiles_w * ceil_div(d->cout, 128);
  const int num_kb_host = taps * (in.c / TC_BLOCK_K);
  // Third case (TD2-FANet's stride-32/64 maps: 512-channel 3x3 convs on a few thousand pixels, 72 K blocks): when
  // 128-wide tiles cover less than half of the SMs, N = 64 doubles the CTAs at work whatever the K length.
  int block_n = (d->cout <= 64 || (tiles128 < 2ll * g_num_sms && num_kb_host <= 24) || 2 * tiles128 <= g_num_sms) ? 64 : 128;
  p.n_tiles_n = ceil_div(d->cout, block_n);
  long long num_tiles = (long long)in.n * p.tiles_h * p.tiles_w * p.n_tiles_n;
  TDN_REQUIRE(num_tiles < (1ll << 31), TDN_ERR_UNSUPPORTED, "conv2d_tc: too many tiles");
  p.num_tiles = (int)num_tiles;
  {
    // K blocks per TMEM accumulation chunk (see the header comment).  4 is the measured sweet spot;
    // TDNET_TC_CHUNK_KB overrides it for experiments (a huge value = plain in-TMEM accumulation).
    static int chunk_kb = 0;
    if (chunk_kb == 0) {
      const char* e = getenv("TDNET_TC_CHUNK_KB");
      chunk_kb = e ? atoi(e) : 4;
      if (chunk_kb < 1) chunk_kb = 4;
    }
    p.chunk_kb = chunk_kb;
  }
  p.w_batched = d->weight_batched;
  p.scale = d->scale; p.bias = d->bias; p.bias_along_m = d->bias_along_m;
  p.act = d->act; p.slope = d->leaky_slope;
  // The epilogue stores 32-channel chunks with 16-byte vectors (a ragged last chunk falls back to scalar
  // stores), so pixel rows must start 16-byte aligned: pitch % 8 (fp16 planes) / % 4 (fp32).
  if (out16) {
    p.out_hi = (__half*)out.data; p.out_lo = (__half*)out.data_lo;
    TDN_REQUIRE(aligned16(out.data) && aligned16(out.data_lo) && out.stride_w % 8 == 0 && out.stride_h % 8 == 0 &&
                    out.stride_n % 8 == 0, TDN_ERR_INVALID, "conv2d_tc: SPLIT16 output rows must be 16-byte aligned");
  } else {
    p.out_f32 = (float*)out.data;
    TDN_REQUIRE(aligned16(out.data) && out.stride_w % 4 == 0 && out.stride_h % 4 == 0 && out.stride_n % 4 == 0,
                TDN_ERR_INVALID, "conv2d_tc: fp32 output rows must be 16-byte aligned");
  }
  if (d->out_f32_copy) {
    TDN_REQUIRE(out16 && aligned16(d->out_f32_copy), TDN_ERR_INVALID, "conv2d_tc: out_f32_copy needs a SPLIT16 primary output");
    p.out_f32 = d->out_f32_copy;  // same element strides as `out`
  }
  p.osn = out.stride_n; p.osh = out.stride_h; p.osw = out.stride_w;
  if (d->residual.data) {
    const tdn_tensor& r = d->residual;
    TDN_REQUIRE(r.n == out.n && r.h == out.h && r.w == out.w && r.c == out.c, TDN_ERR_INVALID,
                "conv2d_tc: residual dims must equal output dims");
    if (r.dtype == TDN_SPLIT16) {
      TDN_REQUIRE(r.data_lo && aligned16(r.data) && aligned16(r.data_lo) && r.stride_w % 8 == 0 &&
                      r.stride_h % 8 == 0 && r.stride_n % 8 == 0, TDN_ERR_INVALID, "conv2d_tc: misaligned residual");
      p.res_hi = (const __half*)r.data; p.res_lo = (const __half*)r.data_lo;
    } else {
      TDN_REQUIRE(aligned16(r.data) && r.stride_w % 4 == 0 && r.stride_h % 4 == 0 && r.stride_n % 4 == 0,
                  TDN_ERR_INVALID, "conv2d_tc: misaligned residual");
      p.res_f32 = (const float*)r.data;
    }
    p.rsn = r.stride_n; p.rsh = r.stride_h; p.rsw = r.stride_w;
  }
  p.range_flag = d->range_flag;
  p.fast = (d->flags & TDN_TC_FLAG_FAST) ? 1 : 0;

  {
    // Kernel choice.  d->variant forces one (tests / tuning); otherwise:
    //  * CTA pairs (tc_conv_pair.cu) for wide layers: M 256 x N 256 tiles over two SMs halve the shared-memory
    //    traffic per MMA, which is what bounds the single-CTA kernel.  TDNET_TC_PAIR = 0 / 1 forces off / on.
    //  * halo regions (tc_conv_halo.cu) for 3x3 / stride 1 / dilation <= 2: one region load per channel block
    //    instead of nine tap loads.  Measured on B200 (profiles/r01_tc_probe_stem_halo.txt): wins where the
    //    per-tap kernel is L2->SM bound with two channel blocks (layer 2: 128 -> 128, 0.048 -> 0.039 ms), loses
    //    on layer 1 (64 -> 64: one block per tile, its 16-byte-row TMA boxes cost more than they save) and on
    //    layer 3 (dilation-2 regions are 2x the tile).  TDNET_TC_HALO = 0 / 1 forces it off / on.
    static int pair_env = -2, halo_env = -2;
    if (pair_env == -2) {
      const char* e = getenv("TDNET_TC_PAIR");
      pair_env = e ? atoi(e) : -1;
      e = getenv("TDNET_TC_HALO");
      halo_env = e ? atoi(e) : -1;
    }
    const bool pair_ok = !d->weight_batched && d->cout % 128 == 0;
    const bool halo_ok = d->kh == 3 && d->kw == 3 && cs == 1 && d->dilation <= 2 && !d->weight_batched && in.w >= 8;
    TDN_REQUIRE(d->variant >= TDN_TC_AUTO && d->variant <= TDN_TC_PAIR_BAND, TDN_ERR_INVALID, "conv2d_tc: unknown variant");
    const bool band_ok = d->kh == 3 && d->kw == 3 && cs == 1 && d->dilation >= 1 && d->dilation <= 4 && !d->weight_batched &&
                         d->cout % 256 == 0 && in.w >= 8;
    TDN_REQUIRE(d->variant != TDN_TC_PAIR_BAND || band_ok, TDN_ERR_UNSUPPORTED,
                "conv2d_tc: the band kernel needs a 3x3 stride-1 convolution with dilation <= 4, cout %% 256 == 0, shared weights");
    if (d->variant == TDN_TC_PAIR_BAND) return conv2d_tc_pair_band(d, p, g_num_sms, stream);
    const bool pair_forced = d->variant == TDN_TC_PAIR || d->variant == TDN_TC_PAIR_TAIL || d->variant == TDN_TC_PAIR_QUAD;
    TDN_REQUIRE(d->variant != TDN_TC_HALO_SW || halo_ok, TDN_ERR_UNSUPPORTED,
                "conv2d_tc: the swizzled halo kernel needs a 3x3 stride-1 convolution with dilation <= 2, width >= 8");
    TDN_REQUIRE((d->variant != TDN_TC_PAIR_TAIL && d->variant != TDN_TC_PAIR_QUAD) || (pair_ok && d->cout % 256 == 0), TDN_ERR_UNSUPPORTED,
                "conv2d_tc: the tail / quad variants of the CTA-pair kernel need cout %% 256 == 0 and shared weights");
    TDN_REQUIRE(!pair_forced || pair_ok, TDN_ERR_UNSUPPORTED,
                "conv2d_tc: the CTA-pair kernel needs cout %% 128 == 0 and shared weights");
    TDN_REQUIRE(d->variant != TDN_TC_HALO || halo_ok, TDN_ERR_UNSUPPORTED,
                "conv2d_tc: the halo kernel needs a 3x3 stride-1 convolution with dilation <= 2");
    // measured (profiles/r01_tc_probe_pair.txt): layer 4 0.342 -> 0.308 ms, layer 3 0.106 -> 0.092 ms; the 1x1
    // 512 -> 512 GEMMs with only 8 K blocks per tile lose 8 % (fewer, longer tiles: 3.5 waves -> 4), N = 128
    // pair tiles are a wash
    // ... and only when the M 256 x N 256 pair tiles fill at least one wave of the 74 clusters: on small maps
    // (TD2-FANet layer 3 / 4 at 1024x2048: 32 / 16 pair tiles, measured 77 us per launch) the single-CTA kernel's
    // smaller tiles keep more SMs busy.
    const long long pair_tiles256 = ((long long)in.n * p.tiles_h * p.tiles_w + 1) / 2 * (d->cout / 256);
    const bool pair_auto = d->cout % 256 == 0 && (pair_env > 0 || (pair_env < 0 && num_kb_host >= 16 &&
                                                                    2 * pair_tiles256 >= g_num_sms));
    if (pair_forced || (d->variant == TDN_TC_AUTO && pair_ok && pair_auto)) {
      const int pair_n = d->cout % 256 == 0 ? 256 : 128;
      // Two alternatives were built, are bit-identical (same products in the same order per output element) and are kept
      // as explicit variants because neither moves the frame rate:
      //  * TDN_TC_PAIR_TAIL -- wave quantisation: pair tiles are big (layer 4 at 1024x2048: 256 tiles on 74 clusters = 3.46
      //    rounds) and the launch time follows the ROUNDS, not the work (tools/quant_probe.py: 208 / 224 / 256 / 288 / 304
      //    tiles take 0.259 / 0.328 / 0.334 / 0.352 / 0.416 ms).  The full rounds run as N = 256 tiles and the remaining
      //    rows of pair tiles as N = 128 tiles in a second launch.  But an N = 128 pair tile is shared-memory bound (6 KB of
      //    operands per 64 tensor cycles) and costs ~0.7 of a full one: 0.341 -> 0.331 ms standalone, nothing in frames.
      //  * TDN_TC_PAIR_QUAD -- clusters of two pairs share the weight tile by TMA multicast (half the L2 reads of B): 33
      //    resident clusters of four instead of 74 of two, tile time 0.0836 -> 0.0823 ms: the kernel is not L2-bound.
      //  (Earlier: the ragged round on the SINGLE-CTA kernel gained nothing either.)  TDNET_TC_PAIR_MODE=tail|quad
      //  applies them to TDN_TC_AUTO for A/B runs.
      static int mode_env = -1;
      if (mode_env < 0) {
        const char* e = getenv("TDNET_TC_PAIR_MODE");
        mode_env = !e ? 0 : !strcmp(e, "tail") ? 1 : !strcmp(e, "quad") ? 2 : !strcmp(e, "band") ? 3 : 0;
      }
      const int mode = d->variant == TDN_TC_PAIR_TAIL ? 1 : d->variant == TDN_TC_PAIR_QUAD ? 2
                       : (d->variant == TDN_TC_AUTO && pair_n == 256) ? mode_env : 0;
      int clusters = 0, rc;
      if (mode == 3 && band_ok) return conv2d_tc_pair_band(d, p, g_num_sms, stream);
      if (mode == 2) {
        int quads = 0;
        if ((rc = conv2d_tc_quad_clusters(g_num_sms, &quads))) return rc;
        TDN_REQUIRE(quads > 0 || d->variant == TDN_TC_AUTO, TDN_ERR_UNSUPPORTED, "conv2d_tc: no resident 4-CTA clusters");
        if (quads > 0) return conv2d_tc_pair(d, p, 256, g_num_sms, 0, stream, 0, true);
      }
      if ((rc = conv2d_tc_pair_clusters(pair_n, g_num_sms, &clusters))) return rc;
      if (mode == 1) {
        const int ntn_pair = ceil_div(d->cout, pair_n);
        const long long m_tiles = (long long)in.n * p.tiles_h * p.tiles_w;
        const long long pair_rows = (m_tiles + 1) / 2;
        const long long pair_tiles = pair_rows * ntn_pair;
        long long head = pair_tiles / clusters * clusters;
        head -= head % ntn_pair;                                    // whole rows of pair tiles only
        const long long tail_tiles = (pair_rows - head / ntn_pair) * (d->cout / 128);
        const double cost_plain = (double)ceil_div((int)pair_tiles, clusters);
        const double cost_split = (double)(head / clusters) + 0.7 * (double)ceil_div((int)tail_tiles, clusters) + 0.05;
        if (head > 0 && head < pair_tiles && (cost_split < cost_plain || d->variant == TDN_TC_PAIR_TAIL)) {
          if ((rc = conv2d_tc_pair(d, p, 256, g_num_sms, (int)head, stream, 0))) return rc;
          return conv2d_tc_pair(d, p, 128, g_num_sms, 0, stream, (int)(head / ntn_pair));
        }
      }
      return conv2d_tc_pair(d, p, pair_n, g_num_sms, 0, stream);
    }
    const bool halo_auto = halo_env > 0 || (halo_env < 0 && d->dilation == 1 && in.c == 128 && d->cout <= 128);
    // the 128-byte-swizzled halo kernel (tc_conv_halo_sw.cu): TDNET_TC_HALO_SW = 0 never, 1 wherever it fits, default: see below
    static int halo_sw_env = -2;
    if (halo_sw_env == -2) {
      const char* e = getenv("TDNET_TC_HALO_SW");
      halo_sw_env = e ? atoi(e) : -1;
    }
    const bool halo_sw_auto = halo_sw_env > 0 || (halo_sw_env < 0 && false);
    if (d->variant == TDN_TC_HALO_SW || (d->variant == TDN_TC_AUTO && halo_ok && halo_sw_auto && d->cout <= 128))
      return conv2d_tc_halo_sw(d, p, g_num_sms, p.chunk_kb, stream);
    if (d->variant == TDN_TC_HALO || (d->variant == TDN_TC_AUTO && halo_ok && halo_auto && p.tile_begin == 0))
      return conv2d_tc_halo(d, p, g_num_sms, p.chunk_kb, stream);
  }

  CUtensorMap a_hi, a_lo, b_hi, b_lo;
  {
    cuuint64_t dims[4] = {(cuuint64_t)in.c, (cuuint64_t)in.w, (cuuint64_t)in.h, (cuuint64_t)in.n};
    cuuint64_t str[3] = {(cuuint64_t)in.stride_w * 2, (cuuint64_t)in.stride_h * 2, (cuuint64_t)in.stride_n * 2};
    // box extents are in source pixels; elementStrides picks every cs-th one, so BW*cs x BH*cs source
    // pixels deliver BW x BH rows of 128 bytes
    cuuint32_t box[4] = {(cuuint32_t)TC_BLOCK_K, (cuuint32_t)(p.BW * cs), (cuuint32_t)(p.BH * cs), 1};
    cuuint32_t est[4] = {1, (cuuint32_t)cs, (cuuint32_t)cs, 1};
    int rc;
    if ((rc = encode_map_f16(&a_hi, in.data, 4, dims, str, box, "A.hi", est, 1))) return rc;
    if ((rc = encode_map_f16(&a_lo, in.data_lo, 4, dims, str, box, "A.lo", est, 1))) return rc;
  }
  {
    const int nb = d->weight_batched ? in.n : 1;
    cuuint64_t bstride = d->weight_batched ? (cuuint64_t)d->weight_batch_stride * 2
                                           : (cuuint64_t)d->weight_ld * 2 * (cuuint64_t)d->cout;
    cuuint64_t dims[3] = {(cuuint64_t)taps * in.c, (cuuint64_t)d->cout, (cuuint64_t)nb};
    cuuint64_t str[2] = {(cuuint64_t)d->weight_ld * 2, bstride};
    cuuint32_t box[3] = {(cuuint32_t)TC_BLOCK_K, (cuuint32_t)block_n, 1};
    int rc;
    if ((rc = encode_map_f16(&b_hi, d->weight_hi, 3, dims, str, box, "B.hi", nullptr, 1))) return rc;
    if ((rc = encode_map_f16(&b_lo, d->weight_lo, 3, dims, str, box, "B.lo", nullptr, 1))) return rc;
  }
  // A through tensor memory (TSA, see TcCfg): TDN_TC_BASE_TS forces it, TDNET_TC_TSA = 0 / 1 switches it for TDN_TC_AUTO / TDN_TC_BASE
  static int tsa_env = -2;
  if (tsa_env == -2) {
    const char* e = getenv("TDNET_TC_TSA");
    tsa_env = e ? atoi(e) : 0;
  }
  const bool tsa = d->variant == TDN_TC_BASE_TS || (d->variant != TDN_TC_BASE && tsa_env > 0);
  if (tsa) {
    if (block_n == 64) return launch_tc<64, true>(a_hi, a_lo, b_hi, b_lo, p, stream);
    return launch_tc<128, true>(a_hi, a_lo, b_hi, b_lo, p, stream);
  }
  if (block_n == 64) return launch_tc<64, false>(a_hi, a_lo, b_hi, b_lo, p, stream);
  return launch_tc<128, false>(a_hi, a_lo, b_hi, b_lo, p, stream);
}

}  // namespace tdn
