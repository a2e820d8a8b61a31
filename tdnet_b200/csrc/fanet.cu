// Kernels of the TD2-FANet inference path that are not convolutions: the fast-attention
// module's linear attention (Training/ptsemseg/models/td2_fanet/td2_fa.py:352-370) and the FPN top-down add
// (:378-395, 398-402).  fp32 arithmetic on F32 or SPLIT16 NHWC views; none of this is tensor-core shaped (the
// contraction over pixels has 32 rows), the rules that matter are coalesced vector access and fixed-order sums.
#include "common.cuh"

namespace tdn {

constexpr int FA_DK = 32;        // w_qs / w_ks output channels (td2_fa.py:339-341)
constexpr int FA_CT = 64;        // value channels per block
constexpr int FA_PX = 64;        // pixels per shared-memory stage
constexpr int FA_CHUNK = 256;    // pixels per partial sum (first reduction stage)

// L2-normalise the 32 channels of one pixel the way F.normalize(p=2, eps=1e-12) does: x / max(||x||, eps).
// Eight consecutive lanes hold one pixel (a float4 each).
__device__ __forceinline__ float4 fa_normalize8(float4 v) {
  float s = v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  s += __shfl_xor_sync(0xffffffffu, s, 1);
  s += __shfl_xor_sync(0xffffffffu, s, 2);
  s += __shfl_xor_sync(0xffffffffu, s, 4);
  const float d = fmaxf(sqrtf(s), 1e-12f);
  return make_float4(__fdiv_rn(v.x, d), __fdiv_rn(v.y, d), __fdiv_rn(v.z, d), __fdiv_rn(v.w, d));
}

// ---------------------------------------------------------------------------------------------
// f[b][j][c] = sum_p khat[b][p][j] * v[b][p][c]      (td2_fa.py:361-366: key normalised over its 32 channels,
// f = matmul(key, value)).  Stage 1: one block per (256-pixel chunk, 64-channel tile, image) -> part.
// Stage 2: chunks summed in index order in fp64 -> bit-reproducible, no atomics.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) fa_context_partial_kernel(View key, View val, float* __restrict__ part,
                                                                 int chunks) {
  __shared__ __align__(16) float sk[FA_PX][FA_DK];
  __shared__ __align__(16) float sv[FA_PX][FA_CT];
  const int chunk = blockIdx.x, c0 = blockIdx.y * FA_CT, b = blockIdx.z;
  const int P = key.h * key.w;
  const int p_begin = chunk * FA_CHUNK, p_end = min(p_begin + FA_CHUNK, P);
  const int tid = threadIdx.x;
  const int c4 = tid & 15, jg = tid >> 4;               // 4 channels x 2 key rows per thread
  float acc[2][4] = {};
  for (int p0 = p_begin; p0 < p_end; p0 += FA_PX) {
    // keys: 64 pixels x 8 float4; thread -> (pixel = idx / 8, quad = idx % 8), two rounds
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int idx = tid + r * 256;
      const int px = idx >> 3, q = idx & 7;
      const int p = p0 + px;
      float4 kv = make_float4(0.f, 0.f, 0.f, 0.f);
      if (p < p_end) {
        const int y = p / key.w, x = p - y * key.w;
        kv = ld4(key, b * key.sn + y * key.sh + x * key.sw + q * 4);
      }
      kv = fa_normalize8(kv);                            // all 8 lanes of a pixel take the same branch above
      *reinterpret_cast<float4*>(&sk[px][q * 4]) = kv;
    }
    // values: 64 pixels x 16 float4, four rounds
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int idx = tid + r * 256;
      const int px = idx >> 4, q = idx & 15;
      const int p = p0 + px;
      float4 vv = make_float4(0.f, 0.f, 0.f, 0.f);
      if (p < p_end && c0 + q * 4 < val.c) {
        const int y = p / val.w, x = p - y * val.w;
        vv = ld4(val, b * val.sn + y * val.sh + x * val.sw + c0 + q * 4);
      }
      *reinterpret_cast<float4*>(&sv[px][q * 4]) = vv;
    }
    __syncthreads();
#pragma unroll 8
    for (int px = 0; px < FA_PX; ++px) {
      const float2 k2 = *reinterpret_cast<const float2*>(&sk[px][jg * 2]);
      const float4 v4 = *reinterpret_cast<const float4*>(&sv[px][c4 * 4]);
      acc[0][0] = fmaf(k2.x, v4.x, acc[0][0]); acc[0][1] = fmaf(k2.x, v4.y, acc[0][1]);
      acc[0][2] = fmaf(k2.x, v4.z, acc[0][2]); acc[0][3] = fmaf(k2.x, v4.w, acc[0][3]);
      acc[1][0] = fmaf(k2.y, v4.x, acc[1][0]); acc[1][1] = fmaf(k2.y, v4.y, acc[1][1]);
      acc[1][2] = fmaf(k2.y, v4.z, acc[1][2]); acc[1][3] = fmaf(k2.y, v4.w, acc[1][3]);
    }
    __syncthreads();
  }
  const int c = c0 + c4 * 4;
  if (c < val.c) {
#pragma unroll
    for (int jj = 0; jj < 2; ++jj) {
      const int j = jg * 2 + jj;
      float* dst = part + (((long long)b * chunks + chunk) * FA_DK + j) * val.c + c;
      *reinterpret_cast<float4*>(dst) = make_float4(acc[jj][0], acc[jj][1], acc[jj][2], acc[jj][3]);
    }
  }
}

__global__ void fa_context_final_kernel(const float* __restrict__ part, float* __restrict__ f, int chunks, int C,
                                        int n) {
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long per_img = (long long)FA_DK * C;
  if (idx >= per_img * n) return;
  const int b = (int)(idx / per_img);
  const long long jc = idx - b * per_img;
  double s = 0.0;
  for (int k = 0; k < chunks; ++k) s += (double)part[((long long)b * chunks + k) * per_img + jc];
  f[idx] = (float)s;
}

size_t fa_context_workspace_bytes(int n, int h, int w, int c) {
  return (size_t)n * ceil_div((long long)h * w, FA_CHUNK) * FA_DK * c * sizeof(float);
}

int fa_context(const tdn_tensor* key, const tdn_tensor* value, float* f, void* workspace, size_t workspace_bytes,
               cudaStream_t stream) {
  int rc;
  if ((rc = check_tensor(key, "fa_context.key"))) return rc;
  if ((rc = check_tensor(value, "fa_context.value"))) return rc;
  TDN_REQUIRE(f != nullptr && aligned16(f), TDN_ERR_INVALID, "fa_context: f must be a 16-byte aligned device pointer");
  TDN_REQUIRE(key->c == FA_DK, TDN_ERR_UNSUPPORTED, "fa_context: key must have %d channels, has %d", FA_DK, key->c);
  TDN_REQUIRE(key->n == value->n && key->h == value->h && key->w == value->w, TDN_ERR_INVALID,
              "fa_context: key / value maps differ in size");
  TDN_REQUIRE(vec4_ok(*key) && vec4_ok(*value), TDN_ERR_INVALID, "fa_context: float4-aligned views required");
  const int chunks = ceil_div((long long)key->h * key->w, FA_CHUNK);
  const size_t need = fa_context_workspace_bytes(key->n, key->h, key->w, value->c);
  TDN_REQUIRE(workspace && workspace_bytes >= need && aligned16(workspace), TDN_ERR_WORKSPACE,
              "fa_context: workspace %zu < %zu bytes", workspace_bytes, need);
  fa_context_partial_kernel<<<dim3(chunks, ceil_div(value->c, FA_CT), key->n), 256, 0, stream>>>(
      make_view(*key), make_view(*value), (float*)workspace, chunks);
  TDN_LAUNCH_OK();
  const long long total = (long long)key->n * FA_DK * value->c;
  fa_context_final_kernel<<<ceil_div(total, 256), 256, 0, stream>>>((const float*)workspace, f, chunks, value->c,
                                                                    key->n);
  TDN_LAUNCH_OK();
  return TDN_OK;
}

// ---------------------------------------------------------------------------------------------
// y[b][p][c] = sum_j qhat[b][p][j] * f[b][j][c]      (td2_fa.py:358-359, 367: query normalised over its 32
// channels, y = matmul(query, f)); block = 64 pixels x 64 channels, 4 x 4 outputs per thread.  y grows with the
// number of pixels of the map (f is an un-normalised sum over all of them): out_scale, a power of two the caller
// undoes exactly in the scale of the convolution that follows, keeps a SPLIT16 output inside the fp16 range.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) fa_apply_kernel(View query, const float* __restrict__ f, View out,
                                                       float out_scale, int* __restrict__ range_flag) {
  __shared__ __align__(16) float sq[FA_PX][FA_DK + 1];
  __shared__ __align__(16) float sf[FA_DK][FA_CT];
  const int p0 = blockIdx.x * FA_PX, c0 = blockIdx.y * FA_CT, b = blockIdx.z;
  const int P = query.h * query.w, C = out.c;
  const int tid = threadIdx.x;
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int idx = tid + r * 256;
    const int px = idx >> 3, q = idx & 7;
    const int p = p0 + px;
    float4 qv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (p < P) {
      const int y = p / query.w, x = p - y * query.w;
      qv = ld4(query, b * query.sn + y * query.sh + x * query.sw + q * 4);
    }
    qv = fa_normalize8(qv);
    sq[px][q * 4 + 0] = qv.x; sq[px][q * 4 + 1] = qv.y; sq[px][q * 4 + 2] = qv.z; sq[px][q * 4 + 3] = qv.w;
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int idx = tid + r * 256;                       // 32 rows x 16 float4
    const int j = idx >> 4, q = idx & 15;
    float4 fv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c0 + q * 4 < C) fv = *reinterpret_cast<const float4*>(f + ((long long)b * FA_DK + j) * C + c0 + q * 4);
    *reinterpret_cast<float4*>(&sf[j][q * 4]) = fv;
  }
  __syncthreads();
  const int c4 = tid & 15, pg = tid >> 4;                // 4 channels x pixels pg, pg+16, pg+32, pg+48
  float acc[4][4] = {};
#pragma unroll 8
  for (int j = 0; j < FA_DK; ++j) {
    const float4 f4 = *reinterpret_cast<const float4*>(&sf[j][c4 * 4]);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float qj = sq[pg + i * 16][j];
      acc[i][0] = fmaf(qj, f4.x, acc[i][0]); acc[i][1] = fmaf(qj, f4.y, acc[i][1]);
      acc[i][2] = fmaf(qj, f4.z, acc[i][2]); acc[i][3] = fmaf(qj, f4.w, acc[i][3]);
    }
  }
  const int c = c0 + c4 * 4;
  if (c >= C) return;
  bool out_of_range = false;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int p = p0 + pg + i * 16;
    if (p >= P) continue;
    const int y = p / out.w, x = p - y * out.w;
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] *= out_scale;
    out_of_range |= fmaxf(fmaxf(fabsf(acc[i][0]), fabsf(acc[i][1])), fmaxf(fabsf(acc[i][2]), fabsf(acc[i][3]))) > 60000.f;
    st4(out, b * out.sn + y * out.sh + x * out.sw + c, make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]));
  }
  if (out_of_range && out.split && range_flag) *reinterpret_cast<volatile int*>(range_flag) = 1;
}

int fa_apply(const tdn_tensor* query, const float* f, const tdn_tensor* out, float out_scale, int* range_flag,
             cudaStream_t stream) {
  int rc;
  if ((rc = check_tensor(query, "fa_apply.query"))) return rc;
  if ((rc = check_tensor(out, "fa_apply.out"))) return rc;
  TDN_REQUIRE(f != nullptr && aligned16(f), TDN_ERR_INVALID, "fa_apply: f must be a 16-byte aligned device pointer");
  TDN_REQUIRE(query->c == FA_DK, TDN_ERR_UNSUPPORTED, "fa_apply: query must have %d channels, has %d", FA_DK,
              query->c);
  TDN_REQUIRE(query->n == out->n && query->h == out->h && query->w == out->w, TDN_ERR_INVALID,
              "fa_apply: query / output maps differ in size");
  TDN_REQUIRE(vec4_ok(*query) && vec4_ok(*out), TDN_ERR_INVALID, "fa_apply: float4-aligned views required");
  TDN_REQUIRE(out_scale > 0.f, TDN_ERR_INVALID, "fa_apply: out_scale must be positive");
  const int P = query->h * query->w;
  fa_apply_kernel<<<dim3(ceil_div(P, FA_PX), ceil_div(out->c, FA_CT), query->n), 256, 0, stream>>>(
      make_view(*query), f, make_view(*out), out_scale, range_flag);
  TDN_LAUNCH_OK();
  return TDN_OK;
}

// ---------------------------------------------------------------------------------------------
// out = bilinear_align_corners(up -> out size) + (a + b)      (td2_fa.py:373 `W_y + feat`, then _upsample_add
// :398-402).  up == nullptr: out = a + b.  The sum order is the reference's: (a + b) is rounded first.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void fa_src_index(int dst, float scale, int in_size, int& i0, int& i1, float& l1) {
  const float s = scale * (float)dst;
  i0 = min((int)s, in_size - 1);
  i1 = min(i0 + 1, in_size - 1);
  l1 = fminf(fmaxf(s - (float)i0, 0.f), 1.f);
}

__global__ void add_upsampled_kernel(View a, View b, View up, int has_up, View out, float sy, float sx) {
  const int c4n = out.c >> 2;
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long total = (long long)out.n * out.h * out.w * c4n;
  if (idx >= total) return;
  const int c = (int)(idx % c4n) * 4;
  long long t = idx / c4n;
  const int x = (int)(t % out.w); t /= out.w;
  const int y = (int)(t % out.h);
  const int img = (int)(t / out.h);
  const float4 av = ld4(a, img * a.sn + y * a.sh + x * a.sw + c);
  const float4 bv = ld4(b, img * b.sn + y * b.sh + x * b.sw + c);
  float4 s = make_float4(__fadd_rn(av.x, bv.x), __fadd_rn(av.y, bv.y), __fadd_rn(av.z, bv.z), __fadd_rn(av.w, bv.w));
  if (has_up) {
    int y0, y1, x0, x1; float ly, lx;
    fa_src_index(y, sy, up.h, y0, y1, ly);
    fa_src_index(x, sx, up.w, x0, x1, lx);
    const long long base = img * up.sn + c;
    const float4 v00 = ld4(up, base + y0 * up.sh + x0 * up.sw), v01 = ld4(up, base + y0 * up.sh + x1 * up.sw);
    const float4 v10 = ld4(up, base + y1 * up.sh + x0 * up.sw), v11 = ld4(up, base + y1 * up.sh + x1 * up.sw);
    const float wx0 = 1.f - lx, wy0 = 1.f - ly;
#define TDN_BLERP(m) ((v00.m * wx0 + v01.m * lx) * wy0 + (v10.m * wx0 + v11.m * lx) * ly)
    s.x = __fadd_rn(TDN_BLERP(x), s.x); s.y = __fadd_rn(TDN_BLERP(y), s.y);
    s.z = __fadd_rn(TDN_BLERP(z), s.z); s.w = __fadd_rn(TDN_BLERP(w), s.w);
#undef TDN_BLERP
  }
  st4(out, img * out.sn + y * out.sh + x * out.sw + c, s);
}

int add_upsampled(const tdn_tensor* a, const tdn_tensor* b, const tdn_tensor* up, const tdn_tensor* out,
                  cudaStream_t stream) {
  int rc;
  if ((rc = check_tensor(a, "add_upsampled.a"))) return rc;
  if ((rc = check_tensor(b, "add_upsampled.b"))) return rc;
  if ((rc = check_tensor(out, "add_upsampled.out"))) return rc;
  const bool has_up = up != nullptr && up->data != nullptr;
  if (has_up && (rc = check_tensor(up, "add_upsampled.up"))) return rc;
  auto same = [&](const tdn_tensor* t) { return t->n == out->n && t->h == out->h && t->w == out->w && t->c == out->c; };
  TDN_REQUIRE(same(a) && same(b), TDN_ERR_INVALID, "add_upsampled: a / b / out dims mismatch");
  TDN_REQUIRE(!has_up || (up->n == out->n && up->c == out->c), TDN_ERR_INVALID, "add_upsampled: up n/c mismatch");
  TDN_REQUIRE(vec4_ok(*a) && vec4_ok(*b) && vec4_ok(*out) && (!has_up || vec4_ok(*up)), TDN_ERR_INVALID,
              "add_upsampled: float4-aligned views required");
  const float sy = has_up && out->h > 1 ? (float)(up->h - 1) / (float)(out->h - 1) : 0.f;
  const float sx = has_up && out->w > 1 ? (float)(up->w - 1) / (float)(out->w - 1) : 0.f;
  const long long total = (long long)out->n * out->h * out->w * (out->c / 4);
  add_upsampled_kernel<<<ceil_div(total, 256), 256, 0, stream>>>(make_view(*a), make_view(*b),
                                                                 has_up ? make_view(*up) : make_view(*a), has_up,
                                                                 make_view(*out), sy, sx);
  TDN_LAUNCH_OK();
  return TDN_OK;
}

}  // namespace tdn
