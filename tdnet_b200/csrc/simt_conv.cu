// Generic fp32 implicit-GEMM convolution on the CUDA cores (FFMA), NHWC, any kernel size / stride /
// dilation, fused per-channel scale+bias (folded BatchNorm), residual add and activation.
//
// Role: the always-correct path of tdn_conv2d.  It runs every shape the tensor-core kernels do not
// take (3-channel stem, stride-2 convs, tiny PSP / key GEMMs, ragged channel counts such as the
// 19-class classifier) and it is the fp32 yardstick the tcgen05 kernels are tested against.
// M = n*Ho*Wo output pixels, N = cout, K = kh*kw*cin.  128x64 tile, BK = 16, 256 threads, 8x4
// outputs per thread, double-buffered shared memory with register prefetch.
#include "common.cuh"

namespace tdn {

struct ConvParams {
  View in;    // dims unused; planes + format only (strides below)
  View out;
  View res;
  int has_res;
  const float* w;
  const float* scale;
  const float* bias;
  long long isn, ish, isw;   // input strides (elements)
  long long osn, osh, osw;   // output strides
  long long rsn, rsh, rsw;   // residual strides
  long long in_bs, out_bs, res_bs, w_bs;  // batch strides
  int Hin, Win, Cin;
  int Ho, Wo, Cout;
  int kh, kw, stride, pad, dil;
  int M, K;
  int act;
  float slope;
  int w_kn;       // weight layout [K][Cout] instead of [Cout][K]
  int out_vec4;   // float4 stores allowed
  int res_vec4;
};

constexpr int BM = 128, BN = 64, BK = 16, TM = 8, TN = 4;
constexpr int APITCH = BM + 4, BPITCH = BN + 4;

__device__ __forceinline__ float apply_act(float v, int act, float slope) {
  if (act == TDN_ACT_RELU) return fmaxf(v, 0.f);
  if (act == TDN_ACT_LEAKY_RELU) return v > 0.f ? v : v * slope;
  return v;
}

__global__ void __launch_bounds__(256) conv_simt_kernel(const ConvParams p) {
  __shared__ __align__(16) float As[2][BK][APITCH];
  __shared__ __align__(16) float Bs[2][BK][BPITCH];

  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int b = blockIdx.z;
  const long long in_b = (long long)b * p.in_bs;
  const float* __restrict__ w = p.w + (long long)b * p.w_bs;

  // ---- A-load mapping: two rows (r, r+64), one float4 of K each ----
  const int a_row = tid >> 2;        // 0..63
  const int a_kq = (tid & 3) * 4;    // 0,4,8,12
  long long a_base[2];
  int a_ih0[2], a_iw0[2];
  bool a_ok[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    int m = m0 + a_row + i * 64;
    a_ok[i] = m < p.M;
    int mm = a_ok[i] ? m : 0;
    int ow = mm % p.Wo;
    int t = mm / p.Wo;
    int oh = t % p.Ho;
    int n = t / p.Ho;
    a_base[i] = (long long)n * p.isn;
    a_ih0[i] = oh * p.stride - p.pad;
    a_iw0[i] = ow * p.stride - p.pad;
  }
  // ---- B-load mapping ----
  const int b_col = p.w_kn ? (tid & 15) * 4 : (tid >> 2);  // n index within tile
  const int b_k = p.w_kn ? (tid >> 4) : (tid & 3) * 4;     // k index within tile

  float4 a_reg[2];
  float4 b_reg;

  auto load_tiles = [&](int k0) {
    // A: im2col gather, one tap per float4 (cin % 4 == 0 guarantees no tap crossing)
    int k = k0 + a_kq;
    int tap = k / p.Cin;
    int c = k - tap * p.Cin;
    int ky = tap / p.kw;
    int kx = tap - ky * p.kw;
    bool kvalid = k < p.K;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      int ih = a_ih0[i] + ky * p.dil;
      int iw = a_iw0[i] + kx * p.dil;
      bool ok = kvalid && a_ok[i] && ih >= 0 && ih < p.Hin && iw >= 0 && iw < p.Win;
      if (ok) {
        a_reg[i] = ld4(p.in, in_b + a_base[i] + (long long)ih * p.ish + (long long)iw * p.isw + c);
      } else {
        a_reg[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    // B
    if (!p.w_kn) {
      int n = n0 + b_col;
      int kk = k0 + b_k;
      if (n < p.Cout && kk < p.K) {
        b_reg = *reinterpret_cast<const float4*>(w + (long long)n * p.K + kk);
      } else {
        b_reg = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    } else {
      int kk = k0 + b_k;
      int n = n0 + b_col;
      if (kk < p.K && n + 3 < p.Cout) {
        b_reg = *reinterpret_cast<const float4*>(w + (long long)kk * p.Cout + n);
      } else if (kk < p.K) {
        const float* src = w + (long long)kk * p.Cout;
        b_reg.x = n + 0 < p.Cout ? src[n + 0] : 0.f;
        b_reg.y = n + 1 < p.Cout ? src[n + 1] : 0.f;
        b_reg.z = n + 2 < p.Cout ? src[n + 2] : 0.f;
        b_reg.w = 0.f;
      } else {
        b_reg = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
  };

  auto store_tiles = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      int r = a_row + i * 64;
      As[buf][a_kq + 0][r] = a_reg[i].x;
      As[buf][a_kq + 1][r] = a_reg[i].y;
      As[buf][a_kq + 2][r] = a_reg[i].z;
      As[buf][a_kq + 3][r] = a_reg[i].w;
    }
    if (!p.w_kn) {
      Bs[buf][b_k + 0][b_col] = b_reg.x;
      Bs[buf][b_k + 1][b_col] = b_reg.y;
      Bs[buf][b_k + 2][b_col] = b_reg.z;
      Bs[buf][b_k + 3][b_col] = b_reg.w;
    } else {
      *reinterpret_cast<float4*>(&Bs[buf][b_k][b_col]) = b_reg;
    }
  };

  const int tx = tid & 15;   // n direction
  const int ty = tid >> 4;   // m direction
  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  const int ktiles = (p.K + BK - 1) / BK;
  load_tiles(0);
  store_tiles(0);
  __syncthreads();

  for (int kt = 0; kt < ktiles; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < ktiles) load_tiles((kt + 1) * BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4 + 64]);
      float4 bv = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
      float av[TM] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float bw[TN] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bw[j], acc[i][j]);
    }
    if (kt + 1 < ktiles) {
      store_tiles(buf ^ 1);
      __syncthreads();
    }
  }

  // ---- epilogue: rows ty*4+{0..3} and 64+ty*4+{0..3}, cols n0+tx*4+{0..3} ----
  const long long out_b = (long long)b * p.out_bs;
  const long long res_b = (long long)b * p.res_bs;
  const bool res = p.has_res != 0;
  const int nb = n0 + tx * 4;
  float sc[TN], bi[TN];
#pragma unroll
  for (int j = 0; j < TN; ++j) {
    int n = nb + j;
    sc[j] = (p.scale && n < p.Cout) ? p.scale[n] : 1.f;
    bi[j] = (p.bias && n < p.Cout) ? p.bias[n] : 0.f;
  }
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    int m = m0 + ty * 4 + (i & 3) + (i >> 2) * 64;
    if (m >= p.M) continue;
    int ow = m % p.Wo;
    int t = m / p.Wo;
    int oh = t % p.Ho;
    int n = t / p.Ho;
    long long ooff = out_b + (long long)n * p.osn + (long long)oh * p.osh + (long long)ow * p.osw + nb;
    long long roff = res_b + (long long)n * p.rsn + (long long)oh * p.rsh + (long long)ow * p.rsw + nb;
    float v[TN];
#pragma unroll
    for (int j = 0; j < TN; ++j) v[j] = fmaf(acc[i][j], sc[j], bi[j]);
    if (nb + 3 < p.Cout) {
      if (res) {
        if (p.res_vec4) {
          float4 r = ld4(p.res, roff);
          v[0] += r.x; v[1] += r.y; v[2] += r.z; v[3] += r.w;
        } else {
#pragma unroll
          for (int j = 0; j < TN; ++j) v[j] += ld1(p.res, roff + j);
        }
      }
#pragma unroll
      for (int j = 0; j < TN; ++j) v[j] = apply_act(v[j], p.act, p.slope);
      if (p.out_vec4) {
        st4(p.out, ooff, make_float4(v[0], v[1], v[2], v[3]));
      } else {
#pragma unroll
        for (int j = 0; j < TN; ++j) st1(p.out, ooff + j, v[j]);
      }
    } else {
#pragma unroll
      for (int j = 0; j < TN; ++j) {
        if (nb + j < p.Cout) {
          float x = v[j];
          if (res) x += ld1(p.res, roff + j);
          st1(p.out, ooff + j, apply_act(x, p.act, p.slope));
        }
      }
    }
  }
}

int conv2d_simt(const tdn_conv2d_desc* d, cudaStream_t stream) {
  const tdn_tensor& in = d->in;
  const tdn_tensor& out = d->out;
  int rc0;
  if ((rc0 = check_tensor(&in, "conv2d.in"))) return rc0;
  if ((rc0 = check_tensor(&out, "conv2d.out"))) return rc0;
  TDN_REQUIRE(in.c % 4 == 0, TDN_ERR_UNSUPPORTED, "conv2d_simt: cin=%d must be a multiple of 4", in.c);
  TDN_REQUIRE(vec4_ok(in) && d->in_batch_stride % 4 == 0,
              TDN_ERR_INVALID, "conv2d_simt: input view must be vector aligned with strides %% 4 == 0");
  TDN_REQUIRE(aligned16(d->weight), TDN_ERR_INVALID, "conv2d_simt: weight must be 16-byte aligned");
  const int Ho = (in.h + 2 * d->pad - d->dilation * (d->kh - 1) - 1) / d->stride + 1;
  const int Wo = (in.w + 2 * d->pad - d->dilation * (d->kw - 1) - 1) / d->stride + 1;
  TDN_REQUIRE(Ho == out.h && Wo == out.w && out.n == in.n && out.c == d->cout, TDN_ERR_INVALID,
              "conv2d: output dims [%d,%d,%d,%d] do not match computed [%d,%d,%d,%d]", out.n, out.h,
              out.w, out.c, in.n, Ho, Wo, d->cout);
  ConvParams p;
  p.in = make_view(in);
  p.out = make_view(out);
  p.has_res = d->residual.data != nullptr;
  p.res = p.has_res ? make_view(d->residual) : make_view(out);
  p.w = d->weight; p.scale = d->scale; p.bias = d->bias;
  p.isn = in.stride_n; p.ish = in.stride_h; p.isw = in.stride_w;
  p.osn = out.stride_n; p.osh = out.stride_h; p.osw = out.stride_w;
  p.rsn = d->residual.stride_n; p.rsh = d->residual.stride_h; p.rsw = d->residual.stride_w;
  p.in_bs = d->in_batch_stride; p.out_bs = d->out_batch_stride;
  p.res_bs = d->residual_batch_stride; p.w_bs = d->weight_batch_stride;
  p.Hin = in.h; p.Win = in.w; p.Cin = in.c;
  p.Ho = Ho; p.Wo = Wo; p.Cout = d->cout;
  p.kh = d->kh; p.kw = d->kw; p.stride = d->stride; p.pad = d->pad; p.dil = d->dilation;
  long long M = (long long)in.n * Ho * Wo;
  TDN_REQUIRE(M > 0 && M < (1ll << 31), TDN_ERR_UNSUPPORTED, "conv2d_simt: M out of range");
  p.M = (int)M;
  p.K = d->kh * d->kw * in.c;
  p.act = d->act; p.slope = d->leaky_slope;
  p.w_kn = d->weight_kn;
  p.out_vec4 = vec4_ok(out) && (d->out_batch_stride % 4 == 0);
  p.res_vec4 = p.has_res ? (vec4_ok(d->residual) && (d->residual_batch_stride % 4 == 0)) : 0;
  if (p.has_res) {
    if ((rc0 = check_tensor(&d->residual, "conv2d.residual"))) return rc0;
    TDN_REQUIRE(d->residual.n == out.n && d->residual.h == out.h && d->residual.w == out.w &&
                    d->residual.c == out.c, TDN_ERR_INVALID, "conv2d: residual dims must equal output dims");
  }
  if (p.w_kn) {
    // float4 weight loads along cout need cout % 4 == 0 alignment of each K row
    TDN_REQUIRE(d->cout % 4 == 0, TDN_ERR_UNSUPPORTED, "conv2d_simt: weight_kn needs cout %% 4 == 0");
  }
  dim3 grid(ceil_div(p.M, BM), ceil_div(p.Cout, BN), d->batch);
  conv_simt_kernel<<<grid, 256, 0, stream>>>(p);
  TDN_LAUNCH_OK();
  return TDN_OK;
}

}  // namespace tdn
