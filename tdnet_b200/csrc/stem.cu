// Fused ResNet-18/34 stem: conv 7x7 stride 2 pad 3 (3 -> 64) + folded BatchNorm + ReLU + maxpool 3x3
// stride 2 pad 1, reading the caller's NCHW fp32 image directly and writing the pooled NHWC map
// (resnet.py:133-137, 205-208).  fp32 CUDA-core arithmetic (K = 147 is too ragged for the tcgen05 tile
// path and the op is ~1 % of the frame's FLOPs); what the fusion buys is memory traffic: the 512x1024x64
// pre-pool map (134 MB at 1024x2048) never leaves shared memory.
//
// One CTA = 8 x 8 pooled pixels = 17 x 17 conv pixels.  Warp r owns conv row r: its 32 lanes hold two
// output channels each (lane, lane+32) for all 17 pixels of the row, as 17 packed fp32x2 accumulators.
// The arithmetic is FFMA2 (fma.rn.f32x2: two IEEE fp32 FMAs per instruction): plain 3-operand FFMA issues
// at half rate on this part (measured: the scalar version of this kernel and conv_simt both sit at ~37 of
// the nominal 75 TFLOP/s), so operands are laid out as pairs: inputs are stored duplicated ({v, v}) and
// weights interleaved ({w[ch], w[ch+32]}) in shared memory, accumulators are the channel pair.
#include "common.cuh"

namespace tdn {

constexpr int ST_P = 8;                              // pooled tile edge
constexpr int ST_C = 2 * ST_P + 1;                   // 17 conv rows / cols
constexpr int ST_IH = 2 * ST_C + 5;                  // 39 input rows
constexpr int ST_IW = 40;                            // 39 input cols, padded to a multiple of 4
constexpr int ST_K = 147;                            // 3 * 7 * 7
constexpr int ST_THREADS = 32 * ST_C;                // 544: one warp per conv row
constexpr int ST_IN_FLOATS = 2 * 3 * ST_IH * ST_IW;  // 9360: every input value is stored twice ({v, v}) so that a
                                                     // shared-memory load lands directly as an FFMA2 operand pair
constexpr int ST_CONV_PITCH = 66;                    // floats per conv pixel in smem (64 + 2: conflict-free float2 rows)
constexpr int ST_SMEM_FLOATS = ST_IN_FLOATS + ST_K * 64 + ST_C * ST_C * ST_CONV_PITCH + 128;

__device__ __forceinline__ void ffma2(uint64_t& c, uint64_t a, uint64_t b) {   // c.{lo,hi} += a.{lo,hi} * b.{lo,hi}
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(c) : "l"(a), "l"(b));
}
__device__ __forceinline__ void unpack2(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}

struct StemParams {
  const float* img;      // [n,3,H,W] fp32 NCHW ...
  const uint8_t* img_u8; // ... or [n,H,W,3] uint8 HWC with `lut` (device-side frame ingest)
  const float* lut;      // [3][256]: lut[c][v] = float((v/255.0 - mean[c]) / std[c]) computed in fp64 on the host
  const float* w;        // [147][64]  (k = (c*7 + ky)*7 + kx)
  const float* scale;    // [64]
  const float* bias;     // [64]
  View out;              // pooled [n,Hp,Wp,64]
  int H, W, Hc, Wc, Hp, Wp;
};

template <bool U8>
__global__ void __launch_bounds__(ST_THREADS, 1) stem_conv_pool_kernel(const StemParams p) {
  extern __shared__ __align__(16) float st_smem[];
  float* s_in = st_smem;                               // [3][ST_IH][ST_IW]
  float* s_w = s_in + ST_IN_FLOATS;                    // [147][64]
  float* s_conv = s_w + ST_K * 64;                     // [17][17][ST_CONV_PITCH]
  float* s_sb = s_conv + ST_C * ST_C * ST_CONV_PITCH;  // scale[64] | bias[64]

  const int tid = threadIdx.x;
  const int b = blockIdx.z;
  const int py0 = blockIdx.y * ST_P, px0 = blockIdx.x * ST_P;
  const int cy0 = 2 * py0 - 1, cx0 = 2 * px0 - 1;      // first conv pixel of the tile (may be -1)
  const int iy0 = 2 * cy0 - 3, ix0 = 2 * cx0 - 3;      // first input pixel of the patch

  // weights -> [k][lane][2] = {w[k][lane], w[k][lane + 32]}
  for (int i = tid; i < ST_K * 64; i += ST_THREADS) {
    const int k = i >> 6, r = i & 63;
    s_w[i] = __ldg(p.w + k * 64 + (r >> 1) + (r & 1) * 32);
  }
  if (tid < 64) { s_sb[tid] = __ldg(p.scale + tid); s_sb[64 + tid] = __ldg(p.bias + tid); }
  const float* img = p.img + (long long)b * 3 * p.H * p.W;
  const uint8_t* img8 = p.img_u8 + (long long)b * 3 * p.H * p.W;
  for (int i = tid; i < 3 * ST_IH * ST_IW; i += ST_THREADS) {
    const int x = i % ST_IW;
    const int t = i / ST_IW;
    const int y = t % ST_IH;
    const int c = t / ST_IH;
    const int iy = iy0 + y, ix = ix0 + x;
    float v = 0.f;   // zero padding applies to the normalised tensor
    if (iy >= 0 && iy < p.H && ix >= 0 && ix < p.W) {
      if (U8) v = __ldg(p.lut + c * 256 + __ldg(img8 + ((long long)iy * p.W + ix) * 3 + c));
      else v = __ldg(img + ((long long)c * p.H + iy) * p.W + ix);
    }
    reinterpret_cast<float2*>(s_in)[i] = make_float2(v, v);
  }
  __syncthreads();

  {
    const int r = tid >> 5;                            // conv row of the tile (one warp each)
    const int lane = tid & 31;
    uint64_t acc[ST_C];                                // packed {channel lane, channel lane + 32}
#pragma unroll
    for (int j = 0; j < ST_C; ++j) acc[j] = 0ull;
#pragma unroll 1
    for (int cky = 0; cky < 21; ++cky) {               // (input channel, filter row)
      const int c = cky / 7, ky = cky - c * 7;
      const ulonglong2* row = reinterpret_cast<const ulonglong2*>(
          reinterpret_cast<const float2*>(s_in) + (c * ST_IH + 2 * r + ky) * ST_IW);   // 2 duplicated inputs / load
      const uint64_t* wk = reinterpret_cast<const uint64_t*>(s_w) + (cky * 7) * 32 + lane;
      uint64_t wp[7];
#pragma unroll
      for (int kx = 0; kx < 7; ++kx) wp[kx] = wk[kx * 32];
      {  // pixels 0..8 use input columns 0..22
        uint64_t in[24];
#pragma unroll
        for (int q = 0; q < 12; ++q) { const ulonglong2 v = row[q]; in[2 * q] = v.x; in[2 * q + 1] = v.y; }
#pragma unroll
        for (int kx = 0; kx < 7; ++kx)
#pragma unroll
          for (int j = 0; j < 9; ++j) ffma2(acc[j], in[2 * j + kx], wp[kx]);
      }
      {  // pixels 9..16 use input columns 18..38
        uint64_t in[22];
#pragma unroll
        for (int q = 0; q < 11; ++q) { const ulonglong2 v = row[9 + q]; in[2 * q] = v.x; in[2 * q + 1] = v.y; }
#pragma unroll
        for (int kx = 0; kx < 7; ++kx)
#pragma unroll
          for (int j = 9; j < ST_C; ++j) ffma2(acc[j], in[2 * j + kx - 18], wp[kx]);
      }
    }
    // BN + ReLU; conv pixels outside the conv map become 0, which cannot change a max over ReLU outputs
    const int cy = cy0 + r;
    const bool rowok = cy >= 0 && cy < p.Hc;
    const float sc0 = s_sb[lane], sc1 = s_sb[lane + 32], bi0 = s_sb[64 + lane], bi1 = s_sb[96 + lane];
    float* dst = s_conv + (r * ST_C) * ST_CONV_PITCH;
#pragma unroll
    for (int j = 0; j < ST_C; ++j) {
      const int cx = cx0 + j;
      const bool ok = rowok && cx >= 0 && cx < p.Wc;
      float a0, a1;
      unpack2(acc[j], a0, a1);
      dst[j * ST_CONV_PITCH + lane] = ok ? fmaxf(fmaf(a0, sc0, bi0), 0.f) : 0.f;
      dst[j * ST_CONV_PITCH + lane + 32] = ok ? fmaxf(fmaf(a1, sc1, bi1), 0.f) : 0.f;
    }
  }
  __syncthreads();

  // 3x3 stride-2 max pool over the shared conv tile: pooled (py, px) covers conv rows/cols 2p .. 2p+2 of the tile
  for (int u = tid; u < ST_P * ST_P * 16; u += ST_THREADS) {
    const int cq = u & 15;
    const int pp = u >> 4;
    const int px = pp % ST_P, py = pp / ST_P;
    const int gy = py0 + py, gx = px0 + px;
    if (gy >= p.Hp || gx >= p.Wp) continue;
    float4 m = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int dy = 0; dy < 3; ++dy)
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        const float* v = s_conv + ((2 * py + dy) * ST_C + 2 * px + dx) * ST_CONV_PITCH + cq * 4;
        const float2 v0 = *reinterpret_cast<const float2*>(v), v1 = *reinterpret_cast<const float2*>(v + 2);
        m.x = fmaxf(m.x, v0.x); m.y = fmaxf(m.y, v0.y); m.z = fmaxf(m.z, v1.x); m.w = fmaxf(m.w, v1.y);
      }
    st4(p.out, b * p.out.sn + gy * p.out.sh + gx * p.out.sw + cq * 4, m);
  }
}

int stem_conv_pool(const float* nchw, const uint8_t* hwc_u8, const float* lut, int n, int h, int w,
                   const float* weight, const float* scale, const float* bias, const tdn_tensor* out,
                   cudaStream_t stream) {
  TDN_REQUIRE((nchw || (hwc_u8 && lut)) && weight && scale && bias, TDN_ERR_INVALID, "stem: null pointer");
  int rc;
  if ((rc = check_tensor(out, "stem.out"))) return rc;
  StemParams p;
  p.img = nchw; p.img_u8 = hwc_u8; p.lut = lut; p.w = weight; p.scale = scale; p.bias = bias;
  p.out = make_view(*out);
  p.H = h; p.W = w;
  p.Hc = (h - 1) / 2 + 1; p.Wc = (w - 1) / 2 + 1;
  p.Hp = (p.Hc - 1) / 2 + 1; p.Wp = (p.Wc - 1) / 2 + 1;
  TDN_REQUIRE(out->n == n && out->h == p.Hp && out->w == p.Wp && out->c == 64 && vec4_ok(*out), TDN_ERR_INVALID,
              "stem: out must be a vector-aligned [n,%d,%d,64] view", p.Hp, p.Wp);
  const int smem = ST_SMEM_FLOATS * (int)sizeof(float);
  static PerDeviceFlag attr_set;
  const int slot = current_device_slot();
  if (!attr_set.is_set(slot)) {
    TDN_CUDA_OK(cudaFuncSetAttribute(stem_conv_pool_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    TDN_CUDA_OK(cudaFuncSetAttribute(stem_conv_pool_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set.set(slot);
  }
  dim3 grid(ceil_div(p.Wp, ST_P), ceil_div(p.Hp, ST_P), n);
  if (nchw) stem_conv_pool_kernel<false><<<grid, ST_THREADS, smem, stream>>>(p);
  else stem_conv_pool_kernel<true><<<grid, ST_THREADS, smem, stream>>>(p);
  TDN_LAUNCH_OK();
  return TDN_OK;
}

}  // namespace tdn
