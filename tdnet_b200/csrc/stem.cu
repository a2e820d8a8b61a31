// Fused ResNet-18/34 stem: conv 7x7 stride 2 pad 3 (3 -> 64) + folded BatchNorm + ReLU + maxpool 3x3
// stride 2 pad 1, reading the caller's NCHW fp32 image directly and writing the pooled NHWC map
// (resnet.py:133-137, 205-208).  fp32 CUDA-core arithmetic (K = 147 is too ragged for the tcgen05 tile
// path and the op is ~1 % of the frame's FLOPs); what the fusion buys is memory traffic: the 512x1024x64
// pre-pool map (134 MB at 1024x2048) never leaves shared memory.
//
// One CTA = 4 x 16 pooled pixels.  It stages the 23 x 71 x 3 input patch and the 147 x 64 weights in
// shared memory, computes the 9 x 33 conv pixels the pooling windows touch (each thread: 2 adjacent
// pixels x 32 channels = 64 fp32 accumulators, weights broadcast from shared memory), applies
// scale/bias/ReLU into a shared conv tile, then max-pools it and stores 4 channels per thread.
#include "common.cuh"

namespace tdn {

constexpr int ST_PH = 4, ST_PW = 16;                 // pooled tile
constexpr int ST_CH = 2 * ST_PH + 1;                 // 9 conv rows
constexpr int ST_CW = 2 * ST_PW + 1;                 // 33 conv cols
constexpr int ST_CWP = ST_CW + 1;                    // padded to a pair count (34)
constexpr int ST_IH = 2 * ST_CH + 5;                 // 23 input rows
constexpr int ST_IW = 2 * ST_CWP + 5;                // 73 input cols (covers the padded pair)
constexpr int ST_K = 147;                            // 3 * 7 * 7
constexpr int ST_PAIRS = ST_CH * (ST_CWP / 2);       // 153 pixel pairs
constexpr int ST_THREADS = 320;
constexpr int ST_IN_FLOATS = (3 * ST_IH * ST_IW + 3) / 4 * 4;   // keep the float4 regions behind it 16-byte aligned
constexpr int ST_SMEM_FLOATS = ST_IN_FLOATS + ST_K * 64 + ST_CH * ST_CWP * 64 + 128;

struct StemParams {
  const float* img;      // [n,3,H,W]
  const float* w;        // [147][64]  (k = (c*7 + ky)*7 + kx)
  const float* scale;    // [64]
  const float* bias;     // [64]
  View out;              // pooled [n,Hp,Wp,64]
  int H, W, Hc, Wc, Hp, Wp;
};

__global__ void __launch_bounds__(ST_THREADS, 1) stem_conv_pool_kernel(const StemParams p) {
  extern __shared__ __align__(16) float st_smem[];
  float* s_in = st_smem;                               // [3][ST_IH][ST_IW]
  float* s_w = s_in + ST_IN_FLOATS;                    // [147][64]
  float* s_conv = s_w + ST_K * 64;                     // [ST_CH][ST_CWP][64]
  float* s_sb = s_conv + ST_CH * ST_CWP * 64;          // scale[64] | bias[64]

  const int tid = threadIdx.x;
  const int b = blockIdx.z;
  const int py0 = blockIdx.y * ST_PH, px0 = blockIdx.x * ST_PW;
  const int cy0 = 2 * py0 - 1, cx0 = 2 * px0 - 1;      // first conv pixel of the tile (may be -1)
  const int iy0 = 2 * cy0 - 3, ix0 = 2 * cx0 - 3;      // first input pixel of the patch

  for (int i = tid; i < ST_K * 64; i += ST_THREADS) s_w[i] = __ldg(p.w + i);
  if (tid < 64) { s_sb[tid] = __ldg(p.scale + tid); s_sb[64 + tid] = __ldg(p.bias + tid); }
  const float* img = p.img + (long long)b * 3 * p.H * p.W;
  for (int i = tid; i < 3 * ST_IH * ST_IW; i += ST_THREADS) {
    const int x = i % ST_IW;
    const int t = i / ST_IW;
    const int y = t % ST_IH;
    const int c = t / ST_IH;
    const int iy = iy0 + y, ix = ix0 + x;
    float v = 0.f;
    if (iy >= 0 && iy < p.H && ix >= 0 && ix < p.W) v = __ldg(img + ((long long)c * p.H + iy) * p.W + ix);
    s_in[i] = v;
  }
  __syncthreads();

  if (tid < 2 * ST_PAIRS) {
    const int half = tid / ST_PAIRS;                   // channels [32*half, 32*half+32)
    const int pair = tid - half * ST_PAIRS;
    const int prow = pair / (ST_CWP / 2);
    const int pcol = pair - prow * (ST_CWP / 2);
    float acc0[32], acc1[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) { acc0[j] = 0.f; acc1[j] = 0.f; }
    const float* in_base = s_in + (2 * prow) * ST_IW + 4 * pcol;   // pixel 2*pcol -> input col 2*(2*pcol)
    const float* w_base = s_w + half * 32;
    for (int c = 0; c < 3; ++c) {
#pragma unroll 1
      for (int ky = 0; ky < 7; ++ky) {
        const float* row = in_base + (c * ST_IH + ky) * ST_IW;
#pragma unroll
        for (int kx = 0; kx < 7; ++kx) {
          const float a0 = row[kx], a1 = row[kx + 2];
          const float4* wv = reinterpret_cast<const float4*>(w_base + ((c * 7 + ky) * 7 + kx) * 64);
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 w4 = wv[q];
            acc0[q * 4 + 0] = fmaf(a0, w4.x, acc0[q * 4 + 0]); acc1[q * 4 + 0] = fmaf(a1, w4.x, acc1[q * 4 + 0]);
            acc0[q * 4 + 1] = fmaf(a0, w4.y, acc0[q * 4 + 1]); acc1[q * 4 + 1] = fmaf(a1, w4.y, acc1[q * 4 + 1]);
            acc0[q * 4 + 2] = fmaf(a0, w4.z, acc0[q * 4 + 2]); acc1[q * 4 + 2] = fmaf(a1, w4.z, acc1[q * 4 + 2]);
            acc0[q * 4 + 3] = fmaf(a0, w4.w, acc0[q * 4 + 3]); acc1[q * 4 + 3] = fmaf(a1, w4.w, acc1[q * 4 + 3]);
          }
        }
      }
    }
    // BN + ReLU; conv pixels outside the conv map become 0, which cannot change a max over ReLU outputs
    const int cy = cy0 + prow, cxa = cx0 + 2 * pcol, cxb = cxa + 1;
    const bool rowok = cy >= 0 && cy < p.Hc;
    const bool oka = rowok && cxa >= 0 && cxa < p.Wc, okb = rowok && cxb >= 0 && cxb < p.Wc;
    float* da = s_conv + (prow * ST_CWP + 2 * pcol) * 64 + half * 32;
    float* db = da + 64;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      float4 ra, rb;
      const float* sc = s_sb + half * 32 + q * 4;
      const float* bi = sc + 64;
      ra.x = oka ? fmaxf(fmaf(acc0[q * 4 + 0], sc[0], bi[0]), 0.f) : 0.f;
      ra.y = oka ? fmaxf(fmaf(acc0[q * 4 + 1], sc[1], bi[1]), 0.f) : 0.f;
      ra.z = oka ? fmaxf(fmaf(acc0[q * 4 + 2], sc[2], bi[2]), 0.f) : 0.f;
      ra.w = oka ? fmaxf(fmaf(acc0[q * 4 + 3], sc[3], bi[3]), 0.f) : 0.f;
      rb.x = okb ? fmaxf(fmaf(acc1[q * 4 + 0], sc[0], bi[0]), 0.f) : 0.f;
      rb.y = okb ? fmaxf(fmaf(acc1[q * 4 + 1], sc[1], bi[1]), 0.f) : 0.f;
      rb.z = okb ? fmaxf(fmaf(acc1[q * 4 + 2], sc[2], bi[2]), 0.f) : 0.f;
      rb.w = okb ? fmaxf(fmaf(acc1[q * 4 + 3], sc[3], bi[3]), 0.f) : 0.f;
      *reinterpret_cast<float4*>(da + q * 4) = ra;
      *reinterpret_cast<float4*>(db + q * 4) = rb;
    }
  }
  __syncthreads();

  // 3x3 stride-2 max pool over the shared conv tile: pooled (py, px) covers conv rows 2py..2py+2 of the tile
  for (int u = tid; u < ST_PH * ST_PW * 16; u += ST_THREADS) {
    const int cq = u & 15;
    const int pp = u >> 4;
    const int px = pp % ST_PW, py = pp / ST_PW;
    const int gy = py0 + py, gx = px0 + px;
    if (gy >= p.Hp || gx >= p.Wp) continue;
    float4 m = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int dy = 0; dy < 3; ++dy)
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        const float4 v = *reinterpret_cast<const float4*>(s_conv + ((2 * py + dy) * ST_CWP + 2 * px + dx) * 64 + cq * 4);
        m.x = fmaxf(m.x, v.x); m.y = fmaxf(m.y, v.y); m.z = fmaxf(m.z, v.z); m.w = fmaxf(m.w, v.w);
      }
    st4(p.out, b * p.out.sn + gy * p.out.sh + gx * p.out.sw + cq * 4, m);
  }
}

int stem_conv_pool(const float* nchw, int n, int h, int w, const float* weight, const float* scale,
                   const float* bias, const tdn_tensor* out, cudaStream_t stream) {
  TDN_REQUIRE(nchw && weight && scale && bias, TDN_ERR_INVALID, "stem: null pointer");
  int rc;
  if ((rc = check_tensor(out, "stem.out"))) return rc;
  StemParams p;
  p.img = nchw; p.w = weight; p.scale = scale; p.bias = bias;
  p.out = make_view(*out);
  p.H = h; p.W = w;
  p.Hc = (h - 1) / 2 + 1; p.Wc = (w - 1) / 2 + 1;
  p.Hp = (p.Hc - 1) / 2 + 1; p.Wp = (p.Wc - 1) / 2 + 1;
  TDN_REQUIRE(out->n == n && out->h == p.Hp && out->w == p.Wp && out->c == 64 && vec4_ok(*out), TDN_ERR_INVALID,
              "stem: out must be a vector-aligned [n,%d,%d,64] view", p.Hp, p.Wp);
  const int smem = ST_SMEM_FLOATS * (int)sizeof(float);
  static bool attr_set = false;
  if (!attr_set) {
    TDN_CUDA_OK(cudaFuncSetAttribute(stem_conv_pool_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set = true;
  }
  dim3 grid(ceil_div(p.Wp, ST_PW), ceil_div(p.Hp, ST_PH), n);
  stem_conv_pool_kernel<<<grid, ST_THREADS, smem, stream>>>(p);
  TDN_LAUNCH_OK();
  return TDN_OK;
}

}  // namespace tdn
