// Shared pieces of the tcgen05 convolution kernels (tc_conv.cu, tc_conv_halo.cu): launch constants, the
// parameter block, and the epilogue role (chunk drain with fp32 register accumulation + the fused
// scale / bias / residual / activation / re-split / store phase).
#pragma once
#include "common.cuh"
#include "tc_ptx.cuh"

#include <stdlib.h>

namespace tdn {

using namespace ptx;

constexpr int TC_BLOCK_M = 128;
constexpr int TC_BLOCK_K = 64;                       // fp16 elements = 128 bytes = one swizzle row
constexpr int TC_A_PLANE = TC_BLOCK_M * TC_BLOCK_K * 2;  // 16 KiB per hi or lo plane
constexpr int TC_EPI_WARPS = 8;                      // two per TMEM lane quarter, each owning half of the N columns
constexpr int TC_THREADS = 64 + 32 * TC_EPI_WARPS;   // warp 0 TMA, warp 1 MMA, warps 2-9 epilogue

struct TcParams {
  int n_img, Ho, Wo;
  int tiles_h, tiles_w, BH, BW;
  int Cout, Cin;
  int taps_h, taps_w, dil, conv_stride;
  int n_tiles_n, num_tiles;
  int tile_begin;        // first tile of this launch (tiles [tile_begin, num_tiles))
  int chunk_kb;          // K blocks accumulated inside TMEM before the fp32 register accumulation
  int w_batched;
  const float* scale;
  const float* bias;
  int bias_along_m;
  int act;
  float slope;
  __half* out_hi;
  __half* out_lo;
  float* out_f32;
  long long osn, osh, osw;
  const __half* res_hi;
  const __half* res_lo;
  const float* res_f32;
  long long rsn, rsh, rsw;
  int* range_flag;
  int fast;              // TDN_TC_FLAG_FAST: issue only the hi x hi product of every K step
  int quad;              // tc_conv_pair.cu: clusters of TWO pairs that share the weight tile (see pair_tile_coords)
};

// Pair tile index -> (output-channel tile, row of pair tiles = 256 pixels).  Plain pair launch: the channel tiles of a row
// are consecutive.  Quad launch (clusters of two pairs, p.quad): consecutive indices 2s, 2s+1 are the two pairs of a
// cluster working on "super tile" s -- the SAME channel tile of two consecutive rows, so that they can share its weights.
__device__ __forceinline__ void pair_tile_coords(const TcParams& p, int tile, int& nt, int& mrow) {
  if (p.quad) {
    const int st = tile >> 1;
    nt = st % p.n_tiles_n;
    mrow = 2 * (st / p.n_tiles_n) + (tile & 1);
  } else {
    nt = tile % p.n_tiles_n;
    mrow = tile / p.n_tiles_n;
  }
}

// Launch of a persistent tcgen05 kernel, optionally with programmatic dependent launch: the kernel's setup (barrier
// init, TMEM allocation, tensor-map prefetch) overlaps the tail of the previous kernel in the stream; every such kernel
// calls tc_pdl_sync() after its setup and before it touches global memory.  Safe next to any predecessor: a kernel
// that never triggers its dependents early releases them when it completes, which is the ordinary stream order.
// Measured on B200: launches of a few tiles per CTA gain (TD2-FANet call 1.874 -> 1.823 ms), the long launches of the
// td4-psp18 frame do not (323.8 vs 319.1 frames/s, within run-to-run noise but not a gain), so by default only
// `short_launch` launches (<= 2 work items per CTA) ask for it.  TDNET_PDL = 0: never, 1: short launches, 2: always.
inline int tc_pdl_mode() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("TDNET_PDL");
    v = e ? atoi(e) : 1;
  }
  return v;
}

template <typename... KArgs, typename... Args>
inline cudaError_t tc_launch(void (*kernel)(KArgs...), int grid, int block, size_t smem, cudaStream_t stream,
                             bool short_launch, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(block);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (tc_pdl_mode() >= 2 || (tc_pdl_mode() == 1 && short_launch)) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<Args&&>(args)...);
}

// After the kernel's setup: wait for the producers of our inputs, then let the next kernel begin its own setup.
__device__ __forceinline__ void tc_pdl_sync() {
  griddep_wait();
  griddep_launch_dependents();
}

__device__ __forceinline__ float tc_act(float v, int act, float slope) {
  if (act == TDN_ACT_RELU) return fmaxf(v, 0.f);
  if (act == TDN_ACT_LEAKY_RELU) return v > 0.f ? v : v * slope;
  return v;
}


// Epilogue role of warps 2..9 for a persistent CTA that walks tiles `blockIdx.x + i * gridDim.x`; `num_kb` K
// blocks per tile arrive in chunks of p.chunk_kb through the TMEM accumulator ring.
// PAIR = true (tc_conv_pair.cu): the CTA is one half of a two-CTA cluster that walks PAIR tiles
// `cluster + i * clusters` (two adjacent 128-pixel M tiles x BLOCK_N columns); this CTA owns M tile
// 2 * pair_m + rank, and a drained accumulator is handed back on the LEADER CTA's barrier, one arrival per warp.
template <int BLOCK_N, int NUM_ACC, bool PAIR = false>
__device__ __forceinline__ void tc_epilogue_role(const TcParams& p, uint32_t tmem_base, uint64_t* tmem_full,
                                                 uint64_t* tmem_empty, int warp, int lane, int num_kb) {
    // Warp w may only touch TMEM lanes 32*(w%4)..+31, so the 8 epilogue warps pair up per lane quarter:
    // group 0 (warps 2-5) owns accumulator columns [0, N/2), group 1 (warps 6-9) columns [N/2, N).
    constexpr int COLS = BLOCK_N / 2;
    const int group = (warp - 2) >> 2;
    const int quarter = warp & 3;                 // TMEM lane quarter this warp may access
    const int row = quarter * 32 + lane;          // pixel row of the tile
    const int h_local = row / p.BW;
    const int w_local = row - h_local * p.BW;
    int as = 0;
    uint32_t aphase = 0;
    bool out_of_range = false;
    const int tile_first = (PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x) + p.tile_begin;
    const int tile_step = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    for (int tile = tile_first; tile < p.num_tiles; tile += tile_step) {
      int nt = tile % p.n_tiles_n;
      int mt = tile / p.n_tiles_n;
      if (PAIR) {
        pair_tile_coords(p, tile, nt, mt);
        mt = 2 * mt + (int)(blockIdx.x & 1);
      }
      const int tx = mt % p.tiles_w;
      mt /= p.tiles_w;
      const int ty = mt % p.tiles_h;
      const int img = mt / p.tiles_h;
      const int oh = ty * p.BH + h_local;
      const int ow = tx * p.BW + w_local;
      const bool valid = oh < p.Ho && ow < p.Wo && img < p.n_img;
      const long long ooff = (long long)img * p.osn + (long long)oh * p.osh + (long long)ow * p.osw;
      const long long roff = (long long)img * p.rsn + (long long)oh * p.rsh + (long long)ow * p.rsw;
      const float bias_m = (p.bias && p.bias_along_m && valid) ? __ldg(p.bias + oh * p.Wo + ow) : 0.f;

      float acc[COLS];
#pragma unroll
      for (int j = 0; j < COLS; ++j) acc[j] = 0.f;
      for (int kb0 = 0; kb0 < num_kb; kb0 += p.chunk_kb) {
        mbar_wait(&tmem_full[as], aphase);
        tc_fence_after();
        const uint32_t taddr_c = tmem_base + ((uint32_t)(quarter * 32) << 16) + as * BLOCK_N + group * COLS;
#pragma unroll
        for (int chunk = 0; chunk < COLS / 32; ++chunk) {
          uint32_t r[32];
          tmem_ld_32x32(taddr_c + chunk * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) acc[chunk * 32 + j] += __uint_as_float(r[j]);
        }
        tc_fence_before();
        if (PAIR) {
          __syncwarp();
          if (lane == 0) mbar_arrive_leader(&tmem_empty[as]);
        } else {
          mbar_arrive(&tmem_empty[as]);
        }
        if (++as == NUM_ACC) { as = 0; aphase ^= 1; }
      }
#pragma unroll
      for (int chunk = 0; chunk < COLS / 32; ++chunk) {
        const int c0 = nt * BLOCK_N + group * COLS + chunk * 32;
        if (valid && c0 < p.Cout) {
          float v[32];
          const bool full = (c0 + 32 <= p.Cout);
          if (full) {
            // warp-uniform 16-byte loads of the per-channel scale / bias (L1 broadcast)
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              float4 s4 = p.scale ? __ldg(reinterpret_cast<const float4*>(p.scale + c0) + q)
                                  : make_float4(1.f, 1.f, 1.f, 1.f);
              float4 b4 = (p.bias && !p.bias_along_m) ? __ldg(reinterpret_cast<const float4*>(p.bias + c0) + q)
                                                      : make_float4(bias_m, bias_m, bias_m, bias_m);
              v[q * 4 + 0] = fmaf(acc[chunk * 32 + q * 4 + 0], s4.x, b4.x);
              v[q * 4 + 1] = fmaf(acc[chunk * 32 + q * 4 + 1], s4.y, b4.y);
              v[q * 4 + 2] = fmaf(acc[chunk * 32 + q * 4 + 2], s4.z, b4.z);
              v[q * 4 + 3] = fmaf(acc[chunk * 32 + q * 4 + 3], s4.w, b4.w);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int c = c0 + j;
              float s = 1.f, b = bias_m;
              if (c < p.Cout) {
                if (p.scale) s = __ldg(p.scale + c);
                if (p.bias && !p.bias_along_m) b = __ldg(p.bias + c);
              }
              v[j] = fmaf(acc[chunk * 32 + j], s, b);
            }
          }
          if (p.res_hi) {
            if (full) {
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                uint4 h4 = *reinterpret_cast<const uint4*>(p.res_hi + roff + c0 + q * 8);
                uint4 l4 = *reinterpret_cast<const uint4*>(p.res_lo + roff + c0 + q * 8);
                const __half2* hh = reinterpret_cast<const __half2*>(&h4);
                const __half2* ll = reinterpret_cast<const __half2*>(&l4);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  float2 a = __half22float2(hh[e]), b2 = __half22float2(ll[e]);
                  v[q * 8 + e * 2 + 0] += a.x + b2.x;
                  v[q * 8 + e * 2 + 1] += a.y + b2.y;
                }
              }
            } else {
              for (int j = 0; j < 32 && c0 + j < p.Cout; ++j)
                v[j] += __half2float(p.res_hi[roff + c0 + j]) + __half2float(p.res_lo[roff + c0 + j]);
            }
          } else if (p.res_f32) {
            if (full) {
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                float4 f = *reinterpret_cast<const float4*>(p.res_f32 + roff + c0 + q * 4);
                v[q * 4 + 0] += f.x; v[q * 4 + 1] += f.y; v[q * 4 + 2] += f.z; v[q * 4 + 3] += f.w;
              }
            } else {
              for (int j = 0; j < 32 && c0 + j < p.Cout; ++j) v[j] += p.res_f32[roff + c0 + j];
            }
          }
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = tc_act(v[j], p.act, p.slope);

          if (p.out_f32) {
            if (full) {
#pragma unroll
              for (int q = 0; q < 8; ++q)
                *reinterpret_cast<float4*>(p.out_f32 + ooff + c0 + q * 4) =
                    make_float4(v[q * 4 + 0], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
            } else {
              for (int j = 0; j < 32 && c0 + j < p.Cout; ++j) p.out_f32[ooff + c0 + j] = v[j];
            }
          }
          if (p.out_hi) {
            __half2 hi2[16], lo2[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              out_of_range |= fmaxf(fabsf(v[2 * j]), fabsf(v[2 * j + 1])) > 60000.f;
              split_f32x2(v[2 * j], v[2 * j + 1], hi2[j], lo2[j]);
            }
            const __half* hi = reinterpret_cast<const __half*>(hi2);
            const __half* lo = reinterpret_cast<const __half*>(lo2);
            if (full) {
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                *reinterpret_cast<uint4*>(p.out_hi + ooff + c0 + q * 8) = *reinterpret_cast<const uint4*>(&hi[q * 8]);
                *reinterpret_cast<uint4*>(p.out_lo + ooff + c0 + q * 8) = *reinterpret_cast<const uint4*>(&lo[q * 8]);
              }
            } else {
              for (int j = 0; j < 32 && c0 + j < p.Cout; ++j) {
                p.out_hi[ooff + c0 + j] = hi[j];
                p.out_lo[ooff + c0 + j] = lo[j];
              }
            }
          }
        }
      }
    }
    if (out_of_range && p.range_flag) *reinterpret_cast<volatile int*>(p.range_flag) = 1;   // idempotent store: the flag may live in host-mapped memory
}

}  // namespace tdn
