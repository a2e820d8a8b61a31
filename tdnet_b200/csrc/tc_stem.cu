// Fused ResNet-18/34 stem on the tensor cores: conv 7x7 stride 2 pad 3 (3 -> 64) + folded BatchNorm + ReLU +
// maxpool 3x3 stride 2 pad 1 (resnet.py:133-137, 205-208), exact mode (split-fp16 operands, three tcgen05 MMAs
// per K step, fp32 accumulation in TMEM), from the caller's NCHW fp32 image -- or the uint8 HWC camera frame --
// straight to the pooled NHWC map.  Same contract and results class as stem.cu (the fp32 CUDA-core version,
// ~0.42 ms at 1024x2048 against the fp32 pipe's ~37 TFLOP/s); this one is bounded by shared-memory operand reads.
//
// No im2col.  The image is staged in shared memory as two fp16 planes (hi, lo) with 4 channels per pixel
// (r, g, b, 0) = 8 bytes, so that consecutive stride-2 conv columns are exactly 16 bytes apart -- the row
// pitch of a no-swizzle K-major UMMA core matrix.  One UMMA M tile = 128 consecutive conv columns of one conv
// row; a K16 step = filter row ky, filter columns kx = 4*half .. 4*half+3 (two 16-byte chunks of two pixels x
// four channels each), i.e. the A descriptor is just "start = image row 2*oy+ky, + half*32 B; LBO = 16 B;
// SBO = 128 B" into the staged image.  K = 7 rows x 2 halves x 16 = 224 (kx = 7 and c = 3 carry zero weights).
// The weights sit in shared memory for the whole kernel as [chunk 28][cout 64][8] fp16 per plane.
//
// One CTA walks down a strip of 126 conv columns (63 pooled columns) for 2*PB+1 conv rows (PB pooled rows):
//   warps 9-16  loader   image rows (2 per conv row) -> split fp16 -> ring of 5 row pairs (the next pair's global
//                        loads are in flight in registers while the current one is converted)
//   warp  0     MMA      42 MMAs per conv row into one of two 64-column TMEM accumulators
//   warps 1-8   epilogue TMEM -> BN + ReLU (0 outside the conv map: cannot change a max of ReLU outputs)
//                        -> ring of 4 fp32 conv rows in shared memory (16-byte quads XOR-swizzled by column);
//                        two warps per TMEM lane quarter, 32 channels each
//   warps 17-24 pool     3x3/2 max over three conv rows -> pooled row in global memory (F32 or SPLIT16)
// all hand-offs through mbarriers; the 512x1024x64 pre-pool map never leaves the SM.
// ncu of the round-1 layout (4 epilogue warps of 64 channels, 3 conv-row slots, loads issued pixel by pixel): 5.4 K
// cycles per conv row against 2.0 K of MMA time -- first the loader (one DRAM round trip per pixel), then the epilogue
// warps (700 instructions per row on one warp per scheduler, and blocked while the pool holds all three row slots),
// then the four pool warps (busy 84 % of the time) and the four loader warps (400 instructions per row pair at one warp
// per scheduler): eight of each now, 25 warps in all.
#include "tc_common.cuh"

namespace tdn {

constexpr int TS_M = 128;                              // conv columns per strip = UMMA M
constexpr int TS_PW = 63;                              // pooled columns per strip (needs conv columns 0..126)
constexpr int TS_IPX_USED = 2 * TS_M + 6;              // 262 image pixels feed 128 conv columns (kx 0..7)
constexpr int TS_IPX = 264;                            // staged row pitch in pixels
constexpr int TS_ROW_BYTES = TS_IPX * 8;               // 2112 (multiple of 16)
constexpr int TS_DP = 5;                               // image ring depth in row pairs (4 in use + 1 ahead)
constexpr int TS_RING_PLANE = TS_DP * 2 * TS_ROW_BYTES;
constexpr int TS_W_CHUNKS = 28;                        // 7 filter rows x 4 column pairs
constexpr int TS_W_PLANE = TS_W_CHUNKS * 64 * 16;      // 28 KiB per plane
constexpr int TS_CROW_BYTES = TS_M * 64 * 4;           // one conv row, fp32
constexpr int TS_CROWS = 4;                            // three feed a pooled row, the fourth is being written
constexpr int TS_EPI_WARPS = 8;                        // warps 1-8
constexpr int TS_LOAD_WARP0 = 9, TS_POOL_WARP0 = 17;   // warp 0 MMA, warps 1-8 epilogue, 9-16 loader, 17-24 pool
constexpr int TS_LOAD_THREADS = 256, TS_POOL_THREADS = 256;
constexpr int TS_THREADS = 32 * 25;
constexpr int TS_TMEM_COLS = 128;
constexpr int TS_SMEM_BYTES = 2 * TS_W_PLANE + 2 * TS_RING_PLANE + TS_CROWS * TS_CROW_BYTES + 512 + 256 + 128;
static_assert(TS_SMEM_BYTES <= 232448, "stem kernel exceeds the 227 KB shared-memory limit");

struct TcStemParams {
  const float* img;       // [n,3,H,W] fp32 NCHW ...
  const uint8_t* img_u8;  // ... or [n,H,W,3] uint8 HWC with `lut`
  const float* lut;       // [3][256]
  const uint4* w;         // fp16 [2 planes][28 chunks][64 cout][8]
  const float* scale;     // [64]  BatchNorm scale x weight row scale
  const float* bias;      // [64]
  View out;               // pooled [n,Hp,Wp,64]
  int H, W, Hc, Wc, Hp, Wp;
  int PB;                 // pooled rows per CTA
  int act;                // TDN_ACT_RELU (resnet.py:136) or TDN_ACT_LEAKY_RELU (the td2_fanet ResNet, resnet.py:117)
  float slope;
  int* range_flag;
};

template <bool U8, bool LEAKY>
__global__ void __launch_bounds__(TS_THREADS, 1) tc_stem_kernel(const TcStemParams p) {
  extern __shared__ uint8_t ts_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(ts_smem_raw) + 127) & ~(uintptr_t)127);
  uint8_t* s_w = smem;
  uint8_t* s_ring = s_w + 2 * TS_W_PLANE;
  uint8_t* s_rows = s_ring + 2 * TS_RING_PLANE;
  float* s_sb = reinterpret_cast<float*>(s_rows + TS_CROWS * TS_CROW_BYTES);   // scale[64] | bias[64]
  uint64_t* img_full = reinterpret_cast<uint64_t*>(s_sb + 128);
  uint64_t* img_empty = img_full + TS_DP;
  uint64_t* acc_full = img_empty + TS_DP;
  uint64_t* acc_empty = acc_full + 2;
  uint64_t* row_full = acc_empty + 2;
  uint64_t* row_empty = row_full + TS_CROWS;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(row_empty + TS_CROWS);

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const float pool_pad = LEAKY ? -INFINITY : 0.f;
  const int b = blockIdx.z;
  const int px0 = blockIdx.x * TS_PW, py0 = blockIdx.y * p.PB;
  const int npb = min(p.PB, p.Hp - py0);               // pooled rows of this CTA
  const int NR = 2 * npb + 1;                          // conv rows
  const int cx0 = 2 * px0 - 1, cy0 = 2 * py0 - 1;      // conv pixel of M row 0 / conv row 0 (may be -1)
  const int ix0 = 2 * cx0 - 3, iy0 = 2 * cy0 - 3;      // image pixel of staged column 0 / staged row 0

  if (tid == 0) {
    for (int s = 0; s < TS_DP; ++s) { mbar_init(&img_full[s], TS_LOAD_THREADS); mbar_init(&img_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], TS_EPI_WARPS); }
    for (int s = 0; s < TS_CROWS; ++s) { mbar_init(&row_full[s], TS_EPI_WARPS); mbar_init(&row_empty[s], TS_POOL_THREADS / 32); }
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_ptr, TS_TMEM_COLS);
    tmem_relinquish();
  }
  for (int i = tid; i < 2 * TS_W_PLANE / 16; i += TS_THREADS) reinterpret_cast<uint4*>(s_w)[i] = __ldg(p.w + i);
  if (tid < 64) { s_sb[tid] = __ldg(p.scale + tid); s_sb[64 + tid] = __ldg(p.bias + tid); }
  fence_proxy_async_smem();                            // the weights are read by the tensor core (async proxy)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ======================= MMA issuer (warp-uniform control, one elected lane issues) =======================
    constexpr uint32_t idesc = umma_idesc_f16(TS_M, 64);
    const uint32_t w_hi = smem_u32(s_w), w_lo = w_hi + TS_W_PLANE;
    const uint32_t ring_hi = smem_u32(s_ring), ring_lo = ring_hi + TS_RING_PLANE;
    for (int r = 0; r < NR; ++r) {
      // conv row r reads staged rows 2r .. 2r+6 = row pairs r .. r+3; pairs below r+3 were awaited by earlier rows
      for (int u = (r == 0 ? 0 : r + 3); u <= r + 3; ++u) mbar_wait(&img_full[u % TS_DP], (u / TS_DP) & 1);
      mbar_wait(&acc_empty[r & 1], ((r >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (r & 1) * 64;
      if (elect_one()) {
#pragma unroll
        for (int ky = 0; ky < 7; ++ky) {
          const int i = 2 * r + ky;
          const uint32_t rowoff = (uint32_t)((((i >> 1) % TS_DP) * 2 + (i & 1)) * TS_ROW_BYTES);
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            const uint64_t a_hi = umma_desc_k_noswz(ring_hi + rowoff + half * 32, 1, 8);
            const uint64_t a_lo = umma_desc_k_noswz(ring_lo + rowoff + half * 32, 1, 8);
            const uint64_t b_hi = umma_desc_k_noswz(w_hi + (ky * 4 + 2 * half) * 1024, 64, 8);
            const uint64_t b_lo = umma_desc_k_noswz(w_lo + (ky * 4 + 2 * half) * 1024, 64, 8);
            umma_f16(d_tmem, a_hi, b_lo, idesc, (ky | half) != 0);
            umma_f16(d_tmem, a_lo, b_hi, idesc, 1);
            umma_f16(d_tmem, a_hi, b_hi, idesc, 1);
          }
        }
        umma_commit(&img_empty[r % TS_DP]);            // row pair r is not read again
        umma_commit(&acc_full[r & 1]);
      }
      __syncwarp();
    }
  } else if (warp < TS_LOAD_WARP0) {
    // ======================= epilogue: TMEM -> BN + ReLU -> conv-row ring =======================
    const int quarter = warp & 3;                      // TMEM lanes 32*quarter .. +31
    const int chalf = (warp - 1) >> 2;                 // channels 32*chalf .. +31
    const int m = quarter * 32 + lane;                 // conv column of the strip
    const int cx = cx0 + m;
    const bool colok = cx >= 0 && cx < p.Wc;
    const float4* sc4 = reinterpret_cast<const float4*>(s_sb + chalf * 32);
    const float4* bi4 = reinterpret_cast<const float4*>(s_sb + 64 + chalf * 32);
    for (int r = 0; r < NR; ++r) {
      const int slot = r % TS_CROWS;
      mbar_wait(&acc_full[r & 1], (r >> 1) & 1);
      mbar_wait(&row_empty[slot], ((r / TS_CROWS) & 1) ^ 1);
      tc_fence_after();
      const int cy = cy0 + r;
      const bool ok = colok && cy >= 0 && cy < p.Hc;
      uint8_t* dst = s_rows + slot * TS_CROW_BYTES + m * 256;
      uint32_t v[32];
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (r & 1) * 64 + chalf * 32;
      tmem_ld_32x32(taddr, v);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[r & 1]);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        // outside the conv map: the pool's padding.  0 cannot change a max of ReLU outputs; LeakyReLU outputs may be
        // negative, so there the padding is -inf (every 3x3 window holds at least its centre, which is inside)
        float4 o = make_float4(pool_pad, pool_pad, pool_pad, pool_pad);
        if (ok) {
          const float4 sc = sc4[q], bi = bi4[q];
          o.x = fmaf(__uint_as_float(v[q * 4 + 0]), sc.x, bi.x);
          o.y = fmaf(__uint_as_float(v[q * 4 + 1]), sc.y, bi.y);
          o.z = fmaf(__uint_as_float(v[q * 4 + 2]), sc.z, bi.z);
          o.w = fmaf(__uint_as_float(v[q * 4 + 3]), sc.w, bi.w);
          if (LEAKY) {
            o.x = o.x > 0.f ? o.x : o.x * p.slope; o.y = o.y > 0.f ? o.y : o.y * p.slope;
            o.z = o.z > 0.f ? o.z : o.z * p.slope; o.w = o.w > 0.f ? o.w : o.w * p.slope;
          } else {
            o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f);
          }
        }
        *reinterpret_cast<float4*>(dst + (((chalf * 8 + q) ^ (m & 15)) * 16)) = o;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&row_full[slot]);
    }
  } else if (warp < TS_POOL_WARP0) {
    // ======================= loader: image rows -> split fp16, 4 channels per pixel =======================
    // Software-pipelined: all 18 global loads of row pair u+1 (2 rows x 3 pixels per thread x 3 channels) are issued
    // before row pair u is converted and stored, so a thread always has one row pair of loads in flight.  (With the
    // loads issued pixel by pixel the loader was one DRAM round trip per pixel and row: 6 dependent trips = ~5 K
    // cycles per conv row, the bound of the whole kernel at 128 us per 1024x2048 frame.)
    // 256 threads: thread (rr, pl) owns pixels pl, pl + 128, pl + 256 of row rr of every row pair.
    const int tl = tid - TS_LOAD_WARP0 * 32;
    const int rr = tl >> 7, pl = tl & 127;
    const int NP = NR + 3;
    const float* img = p.img + (long long)b * 3 * p.H * p.W;
    const uint8_t* img8 = p.img_u8 + (long long)b * 3 * p.H * p.W;
    const int plane = p.H * p.W;                        // < 2^31 / 3 for any image this kernel is given (host check)
    constexpr int PXI = (TS_IPX + 127) / 128;           // pixels per thread and row (3)
    constexpr uint32_t PAD = 0xffffffffu;               // uint8 frames: "outside the image" (byte values are <= 255)
    bool out_of_range = false;
    // everything that does not depend on the row pair: image column of this thread's pixels
    int ixs[PXI];
    bool colok[PXI];
#pragma unroll
    for (int it = 0; it < PXI; ++it) {
      const int px = pl + it * 128;
      ixs[it] = ix0 + px;
      colok[it] = ixs[it] >= 0 && ixs[it] < p.W && px < TS_IPX_USED;
    }
    // raw[it][c]: the fp32 bit pattern (NCHW image; 0 = the zero padding) or the byte value (uint8 frame; PAD = padding)
    auto issue = [&](int u, uint32_t (&raw)[PXI][3]) {
      const int iy = iy0 + 2 * u + rr;
      const bool rowok = iy >= 0 && iy < p.H;
      const int rowoff = iy * p.W;
#pragma unroll
      for (int it = 0; it < PXI; ++it) {
        raw[it][0] = raw[it][1] = raw[it][2] = U8 ? PAD : 0u;
        if (rowok && colok[it]) {
          const int idx = rowoff + ixs[it];
          if (U8) {
            const uint8_t* q = img8 + (long long)idx * 3;
            raw[it][0] = __ldg(q); raw[it][1] = __ldg(q + 1); raw[it][2] = __ldg(q + 2);
          } else {
            raw[it][0] = __float_as_uint(__ldg(img + idx));
            raw[it][1] = __float_as_uint(__ldg(img + idx + plane));
            raw[it][2] = __float_as_uint(__ldg(img + idx + 2 * plane));
          }
        }
      }
    };
    uint32_t cur[PXI][3], nxt[PXI][3];
    issue(0, cur);
    for (int u = 0; u < NP; ++u) {
      const int slot = u % TS_DP;
      if (u + 1 < NP) issue(u + 1, nxt);
      mbar_wait(&img_empty[slot], ((u / TS_DP) & 1) ^ 1);
      uint8_t* dst_hi = s_ring + (slot * 2 + rr) * TS_ROW_BYTES + pl * 8;
      uint8_t* dst_lo = dst_hi + TS_RING_PLANE;
#pragma unroll
      for (int it = 0; it < PXI; ++it) {
        if (pl + it * 128 < TS_IPX) {
          float c0, c1, c2;                              // zero padding applies to the normalised tensor
          if (U8) {
            const bool in = cur[it][0] != PAD;
            c0 = in ? __ldg(p.lut + cur[it][0]) : 0.f;
            c1 = in ? __ldg(p.lut + 256 + cur[it][1]) : 0.f;
            c2 = in ? __ldg(p.lut + 512 + cur[it][2]) : 0.f;
          } else {
            c0 = __uint_as_float(cur[it][0]);
            c1 = __uint_as_float(cur[it][1]);
            c2 = __uint_as_float(cur[it][2]);
          }
          out_of_range |= fmaxf(fabsf(c0), fmaxf(fabsf(c1), fabsf(c2))) > 60000.f;
          __half2 h[2], l[2];
          split_f32x2(c0, c1, h[0], l[0]);
          split_f32x2(c2, 0.f, h[1], l[1]);
          *reinterpret_cast<uint2*>(dst_hi + it * 128 * 8) = *reinterpret_cast<const uint2*>(h);
          *reinterpret_cast<uint2*>(dst_lo + it * 128 * 8) = *reinterpret_cast<const uint2*>(l);
        }
      }
      fence_proxy_async_smem();
      mbar_arrive(&img_full[slot]);
#pragma unroll
      for (int it = 0; it < PXI; ++it)
#pragma unroll
        for (int c = 0; c < 3; ++c) cur[it][c] = nxt[it][c];
    }
    if (out_of_range && p.range_flag) *reinterpret_cast<volatile int*>(p.range_flag) = 1;   // idempotent store: the flag may live in host-mapped memory
  } else {
    // ======================= pool: 3x3 stride-2 max over three conv rows =======================
    const int tp = tid - TS_POOL_WARP0 * 32;
    for (int j = 0; j < npb; ++j) {
#pragma unroll
      for (int dy = 0; dy < 3; ++dy) {
        const int r = 2 * j + dy;
        mbar_wait(&row_full[r % TS_CROWS], (r / TS_CROWS) & 1);
      }
      const int py = py0 + j;
      const uint8_t* base0 = s_rows + ((2 * j) % TS_CROWS) * TS_CROW_BYTES;
      const uint8_t* base1 = s_rows + ((2 * j + 1) % TS_CROWS) * TS_CROW_BYTES;
      const uint8_t* base2 = s_rows + ((2 * j + 2) % TS_CROWS) * TS_CROW_BYTES;
      for (int item = tp; item < TS_PW * 16; item += TS_POOL_THREADS) {
        const int quad = item & 15;
        const int pxl = item >> 4;
        const int px = px0 + pxl;
        if (px >= p.Wp) continue;
        float4 mx = make_float4(pool_pad, pool_pad, pool_pad, pool_pad);
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
          const int m = 2 * pxl + dx;
          const int off = m * 256 + ((quad ^ (m & 15)) * 16);
          const float4 a = *reinterpret_cast<const float4*>(base0 + off);
          const float4 c = *reinterpret_cast<const float4*>(base1 + off);
          const float4 d = *reinterpret_cast<const float4*>(base2 + off);
          mx.x = fmaxf(mx.x, fmaxf(a.x, fmaxf(c.x, d.x)));
          mx.y = fmaxf(mx.y, fmaxf(a.y, fmaxf(c.y, d.y)));
          mx.z = fmaxf(mx.z, fmaxf(a.z, fmaxf(c.z, d.z)));
          mx.w = fmaxf(mx.w, fmaxf(a.w, fmaxf(c.w, d.w)));
        }
        st4(p.out, (long long)b * p.out.sn + (long long)py * p.out.sh + (long long)px * p.out.sw + quad * 4, mx);
      }
      __syncwarp();
      if (lane == 0) {                                 // conv row 2j+2 is row 2(j+1) of the next pooled row
        mbar_arrive(&row_empty[(2 * j) % TS_CROWS]);
        mbar_arrive(&row_empty[(2 * j + 1) % TS_CROWS]);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TS_TMEM_COLS);
  }
}

int stem_conv_pool_tc(const float* nchw, const uint8_t* hwc_u8, const float* lut, int n, int h, int w,
                      const void* weight_tc, const float* scale, const float* bias, const tdn_tensor* out,
                      int act, float slope, int* range_flag, cudaStream_t stream) {
  TDN_REQUIRE(act == TDN_ACT_RELU || act == TDN_ACT_LEAKY_RELU, TDN_ERR_UNSUPPORTED,
              "stem_tc: activation must be ReLU or LeakyReLU (max pooling commutes with neither 'none' padding rule)");
  TDN_REQUIRE((nchw != nullptr) != (hwc_u8 != nullptr), TDN_ERR_INVALID,
              "stem_tc: exactly one of the fp32 NCHW image and the uint8 HWC frame must be given");
  TDN_REQUIRE((nchw || lut) && weight_tc && scale && bias, TDN_ERR_INVALID, "stem_tc: null pointer");
  TDN_REQUIRE(n >= 1 && h >= 1 && w >= 1, TDN_ERR_INVALID, "stem_tc: empty image");
  TDN_REQUIRE((long long)h * w * 3 < (1ll << 31), TDN_ERR_UNSUPPORTED, "stem_tc: image planes beyond 32-bit indexing");
  TDN_REQUIRE((reinterpret_cast<uintptr_t>(weight_tc) & 15) == 0, TDN_ERR_INVALID, "stem_tc: weights must be 16-byte aligned");
  int rc;
  if ((rc = check_tensor(out, "stem_tc.out"))) return rc;
  TcStemParams p;
  p.img = nchw; p.img_u8 = hwc_u8; p.lut = lut; p.w = reinterpret_cast<const uint4*>(weight_tc);
  p.scale = scale; p.bias = bias;
  p.out = make_view(*out);
  p.H = h; p.W = w;
  p.Hc = (h - 1) / 2 + 1; p.Wc = (w - 1) / 2 + 1;
  p.Hp = (p.Hc - 1) / 2 + 1; p.Wp = (p.Wc - 1) / 2 + 1;
  p.range_flag = range_flag;
  p.act = act; p.slope = slope;
  TDN_REQUIRE(out->n == n && out->h == p.Hp && out->w == p.Wp && out->c == 64 && vec4_ok(*out), TDN_ERR_INVALID,
              "stem_tc: out must be a vector-aligned [n,%d,%d,64] view", p.Hp, p.Wp);
  static PerDeviceFlag attr_set;
  const int slot = current_device_slot();
  if (!attr_set.is_set(slot)) {
    TDN_CUDA_OK(cudaFuncSetAttribute(tc_stem_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TS_SMEM_BYTES));
    TDN_CUDA_OK(cudaFuncSetAttribute(tc_stem_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TS_SMEM_BYTES));
    TDN_CUDA_OK(cudaFuncSetAttribute(tc_stem_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TS_SMEM_BYTES));
    TDN_CUDA_OK(cudaFuncSetAttribute(tc_stem_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TS_SMEM_BYTES));
    attr_set.set(slot);
  }
  // Pooled rows per CTA: every CTA pays ~6 row-times of prologue (weights, pipeline fill) and one halo conv
  // row, and the grid runs in waves of one CTA per SM -- pick the band height with the shortest critical path.
  const int strips = ceil_div(p.Wp, TS_PW);
  const int sms = device_sm_count();
  TDN_REQUIRE(sms > 0, TDN_ERR_CUDA, "stem_tc: cannot query the SM count");
  int best_pb = 1;
  long long best_cost = -1;
  for (int pb = 1; pb <= 64 && pb <= p.Hp; ++pb) {
    const long long ctas = (long long)strips * ceil_div(p.Hp, pb) * n;
    const long long cost = ((ctas + sms - 1) / sms) * (2 * pb + 1 + 6);
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; best_pb = pb; }
  }
  p.PB = best_pb;
  dim3 grid(strips, ceil_div(p.Hp, p.PB), n);
  const bool leaky = act == TDN_ACT_LEAKY_RELU;
  if (nchw) {
    if (leaky) tc_stem_kernel<false, true><<<grid, TS_THREADS, TS_SMEM_BYTES, stream>>>(p);
    else tc_stem_kernel<false, false><<<grid, TS_THREADS, TS_SMEM_BYTES, stream>>>(p);
  } else {
    if (leaky) tc_stem_kernel<true, true><<<grid, TS_THREADS, TS_SMEM_BYTES, stream>>>(p);
    else tc_stem_kernel<true, false><<<grid, TS_THREADS, TS_SMEM_BYTES, stream>>>(p);
  }
  TDN_LAUNCH_OK();
  return TDN_OK;
}

}  // namespace tdn
