// HBM-bound layout / pooling / normalisation / resampling kernels of the frame (fp32, NHWC).
// None of these has data reuse worth tensor cores; the rules that matter are coalesced, vectorised
// access and enough CTAs to fill 148 SMs.
#include "common.cuh"

namespace tdn {

__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---------------------------------------------------------------------------------------------
// NCHW image -> NHWC, channels padded with zeros to 4 (one float4 store per pixel).
// ---------------------------------------------------------------------------------------------
__global__ void image_to_nhwc_kernel(const float* __restrict__ src, int n, int c, int h, int w,
                                     View out) {
  long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  long long total = (long long)n * h * w;
  if (idx >= total) return;
  int x = idx % w;
  long long t = idx / w;
  int y = t % h;
  int b = t / h;
  const long long plane = (long long)h * w;
  const float* s = src + (long long)b * c * plane + (long long)y * w + x;
  float v[4] = {0.f, 0.f, 0.f, 0.f};
  for (int ch = 0; ch < c && ch < 4; ++ch) v[ch] = __ldg(s + ch * plane);
  st4(out, b * out.sn + y * out.sh + x * out.sw, make_float4(v[0], v[1], v[2], v[3]));
}

int image_to_nhwc(const float* nchw, int n, int c, int h, int w, const tdn_tensor* out,
                  cudaStream_t stream) {
  TDN_REQUIRE(nchw != nullptr, TDN_ERR_INVALID, "image_to_nhwc: null pointer");
  int rc;
  if ((rc = check_tensor(out, "image_to_nhwc.out"))) return rc;
  TDN_REQUIRE(c <= 4 && out->c == 4 && out->n == n && out->h == h && out->w == w && vec4_ok(*out),
              TDN_ERR_INVALID, "image_to_nhwc: expects c<=4 and a float4-aligned [n,h,w,4] output");
  long long total = (long long)n * h * w;
  image_to_nhwc_kernel<<<ceil_div(total, 256), 256, 0, stream>>>(nchw, n, c, h, w, make_view(*out));
  TDN_LAUNCH_OK();
  return TDN_OK;
}

// ---------------------------------------------------------------------------------------------
// 3x3 stride-2 pad-1 max pool (padding acts as -inf), float4 over channels.
// ---------------------------------------------------------------------------------------------
__global__ void maxpool3x3s2_kernel(View in, View out) {
  const int c4 = out.c >> 2;
  long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  long long total = (long long)out.n * out.h * out.w * c4;
  if (idx >= total) return;
  int cq = idx % c4;
  long long t = idx / c4;
  int ox = t % out.w; t /= out.w;
  int oy = t % out.h;
  int b = t / out.h;
  float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
#pragma unroll
  for (int dy = 0; dy < 3; ++dy) {
    int iy = oy * 2 - 1 + dy;
    if (iy < 0 || iy >= in.h) continue;
#pragma unroll
    for (int dx = 0; dx < 3; ++dx) {
      int ix = ox * 2 - 1 + dx;
      if (ix < 0 || ix >= in.w) continue;
      float4 v = ld4(in, b * in.sn + iy * in.sh + ix * in.sw + cq * 4);
      m.x = fmaxf(m.x, v.x); m.y = fmaxf(m.y, v.y); m.z = fmaxf(m.z, v.z); m.w = fmaxf(m.w, v.w);
    }
  }
  st4(out, b * out.sn + oy * out.sh + ox * out.sw + cq * 4, m);
}

int maxpool3x3s2(const tdn_tensor* in, const tdn_tensor* out, cudaStream_t stream) {
  int rc;
  if ((rc = check_tensor(in, "maxpool.in"))) return rc;
  if ((rc = check_tensor(out, "maxpool.out"))) return rc;
  TDN_REQUIRE(vec4_ok(*in) && vec4_ok(*out), TDN_ERR_INVALID, "maxpool: float4-aligned views required");
  TDN_REQUIRE(out->h == (in->h - 1) / 2 + 1 && out->w == (in->w - 1) / 2 + 1 && out->c == in->c &&
                  out->n == in->n, TDN_ERR_INVALID, "maxpool: output dims mismatch");
  long long total = (long long)out->n * out->h * out->w * (out->c / 4);
  maxpool3x3s2_kernel<<<ceil_div(total, 256), 256, 0, stream>>>(make_view(*in), make_view(*out));
  TDN_LAUNCH_OK();
  return TDN_OK;
}

// ---------------------------------------------------------------------------------------------
// Pyramid pooling, 50 bins (1x1, 2x2, 3x3, 6x6) in two separable passes.
// Pass 1: per image row, the sums over the 12 column ranges -> rowsum[n][H][12][C]   (reads x once
//         from HBM; the 4 pyramid levels re-read the row from L1/L2).
// Pass 2: per bin, sum its rows of the matching column range and divide by the bin area.
// Bin i of o covers [floor(i*L/o), ceil((i+1)*L/o)) (AdaptiveAvgPool2d; bins overlap when o !| L).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int bin_start(int i, int o, int len) { return (i * len) / o; }
__device__ __forceinline__ int bin_end(int i, int o, int len) { return ((i + 1) * len + o - 1) / o; }

constexpr int PSP_PARTS = 1;   // x-segments per row (tdn_psp_pool_workspace_bytes assumes 1); splitting rows into 4 was
                               // measured slower: it only adds pass-2 reads, the pass is not occupancy-bound

__global__ void psp_rowsum_kernel(View in, float* __restrict__ rowsum) {
  // One pass over a quarter of the row: every pyramid level keeps the running sum of its current column
  // range.  Adjacent ranges of a level overlap by at most one column (ceil vs floor), which seeds the next
  // sum.  Ranges cut by the segment boundary leave partial sums; untouched ranges stay zero.
  const int y = blockIdx.x;
  const int part = blockIdx.z % PSP_PARTS, b = blockIdx.z / PSP_PARTS;
  const int c = blockIdx.y * blockDim.x + threadIdx.x;
  if (c >= in.c) return;
  const long long row = b * in.sn + y * in.sh + c;
  float* dst = rowsum + (((long long)(b * in.h + y) * PSP_PARTS + part) * 12) * in.c + c;
  const int W = in.w;
  const int x_lo = (part * W) / PSP_PARTS, x_hi = ((part + 1) * W) / PSP_PARTS;
  const int lv_o[4] = {1, 2, 3, 6};
  const int lv_off[4] = {0, 1, 3, 6};
#pragma unroll
  for (int r = 0; r < 12; ++r) dst[(long long)r * in.c] = 0.f;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  int cur[4], nend[4];
#pragma unroll
  for (int l = 0; l < 4; ++l) {
    int j = 0;
    while (j < lv_o[l] && bin_end(j, lv_o[l], W) <= x_lo) ++j;   // first range that reaches into the segment
    cur[l] = j;
    nend[l] = bin_end(j, lv_o[l], W);
  }
  constexpr int UN = 8;   // loads issued ahead of the (serial) range bookkeeping
  for (int x0 = x_lo; x0 < x_hi; x0 += UN) {
    float vbuf[UN];
#pragma unroll
    for (int u = 0; u < UN; ++u) vbuf[u] = (x0 + u < x_hi) ? ld1(in, row + (long long)(x0 + u) * in.sw) : 0.f;
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const int x = x0 + u;
      if (x >= x_hi) break;
      const float v = vbuf[u];
#pragma unroll
      for (int l = 0; l < 4; ++l) {
        acc[l] += v;
        if (x + 1 == nend[l]) {
          dst[(long long)(lv_off[l] + cur[l]) * in.c] = acc[l];
          ++cur[l];
          nend[l] = bin_end(cur[l], lv_o[l], W);
          acc[l] = (cur[l] < lv_o[l] && bin_start(cur[l], lv_o[l], W) <= x) ? v : 0.f;
        }
      }
    }
  }
#pragma unroll
  for (int l = 0; l < 4; ++l)      // ranges still open at the segment end keep their partial sum
    if (cur[l] < lv_o[l] && bin_start(cur[l], lv_o[l], W) < x_hi) dst[(long long)(lv_off[l] + cur[l]) * in.c] = acc[l];
}

// Vectorised pass 1 (c % 8 == 0, 16-byte aligned planes, rows wide enough): a thread owns 8 channels (one 16-byte load
// per plane and pixel) and one of RS_SEG x-segments of the row, so 16x more loads are in flight than with one serial
// walk per channel (that version: 65 us for the 67 MB map of a 1024x2048 frame, latency-bound).  A segment is shorter
// than the narrowest column range minus one (host check: 6 (seg_len + 1) <= W), so per pyramid level it meets at most
// two ranges: the first one reaching into it ("A") and the next ("B"; the two may share one column).  The partial sums
// go to shared memory and are added per range in segment order (fixed order: bit-reproducible).
constexpr int RS_SEG = 16;                           // x-segments per row
constexpr int RS_CB = 64;                            // channels per block (8 lanes x 8 channels)
constexpr int RS_SLOTS = 7;                          // level 0: A; levels 1..3: A and B
constexpr int RS_PITCH = RS_SLOTS * RS_CB + 8;       // floats per segment (+8: the 4 segments of a warp use different banks)

__device__ __forceinline__ int first_range(int o, int W, int x_lo) {   // first range of level o that ends behind x_lo
  int j = 0;
  while (j < o - 1 && bin_end(j, o, W) <= x_lo) ++j;
  return j;
}

template <bool SPLIT>
__global__ void __launch_bounds__(RS_SEG * RS_CB / 8) psp_rowsum_vec8_kernel(View in, float* __restrict__ rowsum, int seg_len) {
  __shared__ __align__(16) float part[RS_SEG * RS_PITCH];
  __shared__ int first[RS_SEG][4];                   // first range of level l that reaches into segment s
  __shared__ int seg_lo[12], seg_hi[12];             // segments that meet range r
  const int y = blockIdx.x, b = blockIdx.z;
  const int g = threadIdx.x & 7, s = threadIdx.x >> 3;
  const int c0 = blockIdx.y * RS_CB + g * 8;
  const int W = in.w;
  const int x_lo = min(s * seg_len, W), x_hi = min(x_lo + seg_len, W);
  if (threadIdx.x < RS_SEG * 4) {
    const int s2 = threadIdx.x >> 2, l = threadIdx.x & 3;
    first[s2][l] = first_range(l == 0 ? 1 : l == 1 ? 2 : l == 2 ? 3 : 6, W, min(s2 * seg_len, W));
  } else if (threadIdx.x < RS_SEG * 4 + 12) {
    const int r = threadIdx.x - RS_SEG * 4;
    const int l = r < 1 ? 0 : r < 3 ? 1 : r < 6 ? 2 : 3;
    const int o = l == 0 ? 1 : l == 1 ? 2 : l == 2 ? 3 : 6;
    const int j = r - (l == 0 ? 0 : l == 1 ? 1 : l == 2 ? 3 : 6);
    seg_lo[r] = bin_start(j, o, W) / seg_len;
    seg_hi[r] = min((bin_end(j, o, W) - 1) / seg_len, RS_SEG - 1);
  }
  __syncthreads();
  int endA[4], startB[4];
#pragma unroll
  for (int l = 0; l < 4; ++l) {
    const int o = l == 0 ? 1 : l == 1 ? 2 : l == 2 ? 3 : 6;
    const int j = first[s][l];
    endA[l] = bin_end(j, o, W);
    startB[l] = j + 1 < o ? bin_start(j + 1, o, W) : 0x7fffffff;
  }
  float acc[RS_SLOTS][8];
#pragma unroll
  for (int k = 0; k < RS_SLOTS; ++k)
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[k][e] = 0.f;
  if (c0 < in.c) {
    const long long base = b * in.sn + y * in.sh + c0;
    constexpr int UN = 4;                              // pixels whose loads are issued together
    for (int x0 = x_lo; x0 < x_hi; x0 += UN) {
      float v[UN][8];
#pragma unroll
      for (int u = 0; u < UN; ++u) {
        const long long off = base + (long long)min(x0 + u, x_hi - 1) * in.sw;
        if (SPLIT) {
          const uint4 hv = *reinterpret_cast<const uint4*>(in.hi + off);
          const uint4 lv = *reinterpret_cast<const uint4*>(in.lo + off);
          const __half2* h = reinterpret_cast<const __half2*>(&hv);
          const __half2* lo = reinterpret_cast<const __half2*>(&lv);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 a = __half22float2(h[e]), d = __half22float2(lo[e]);
            v[u][2 * e] = a.x + d.x;
            v[u][2 * e + 1] = a.y + d.y;
          }
        } else {
          *reinterpret_cast<float4*>(&v[u][0]) = *reinterpret_cast<const float4*>(in.p + off);
          *reinterpret_cast<float4*>(&v[u][4]) = *reinterpret_cast<const float4*>(in.p + off + 4);
        }
      }
#pragma unroll
      for (int u = 0; u < UN; ++u) {
        const int x = x0 + u;
        if (x < x_hi) {
#pragma unroll
          for (int l = 0; l < 4; ++l) {
            if (x < endA[l]) {
#pragma unroll
              for (int e = 0; e < 8; ++e) acc[l == 0 ? 0 : 2 * l - 1][e] += v[u][e];
            }
            if (l > 0 && x >= startB[l]) {
#pragma unroll
              for (int e = 0; e < 8; ++e) acc[2 * l][e] += v[u][e];
            }
          }
        }
      }
    }
  }
  float* mine = part + s * RS_PITCH + g * 8;
#pragma unroll
  for (int k = 0; k < RS_SLOTS; ++k) {
    *reinterpret_cast<float4*>(mine + k * RS_CB) = make_float4(acc[k][0], acc[k][1], acc[k][2], acc[k][3]);
    *reinterpret_cast<float4*>(mine + k * RS_CB + 4) = make_float4(acc[k][4], acc[k][5], acc[k][6], acc[k][7]);
  }
  __syncthreads();
  // 12 ranges x RS_CB channels: every segment that meets the range adds its A or B slot, in segment order
  float* dst = rowsum + ((long long)(b * in.h + y) * 12) * in.c + blockIdx.y * RS_CB;
  for (int i = threadIdx.x; i < 12 * RS_CB; i += RS_SEG * RS_CB / 8) {
    const int r = i / RS_CB, cc = i - r * RS_CB;
    int l, j;
    if (r < 1) { l = 0; j = 0; } else if (r < 3) { l = 1; j = r - 1; } else if (r < 6) { l = 2; j = r - 3; } else { l = 3; j = r - 6; }
    float sum = 0.f;
    for (int s2 = seg_lo[r]; s2 <= seg_hi[r]; ++s2) {
      const int j0 = first[s2][l];
      if (j == j0) sum += part[s2 * RS_PITCH + (l == 0 ? 0 : 2 * l - 1) * RS_CB + cc];
      else if (l > 0 && j == j0 + 1) sum += part[s2 * RS_PITCH + 2 * l * RS_CB + cc];
    }
    if (blockIdx.y * RS_CB + cc < in.c) dst[(long long)r * in.c + cc] = sum;
  }
}

// Generic fallback (any width, re-reads the row once per level): used when the map is narrower than the
// finest pyramid level, where ranges overlap by more than one column.
__global__ void psp_rowsum_generic_kernel(View in, float* __restrict__ rowsum) {
  const int y = blockIdx.x, b = blockIdx.y;
  const long long row = b * in.sn + y * in.sh;
  float* dst = rowsum + ((long long)(b * in.h + y) * PSP_PARTS * 12) * in.c;   // everything goes to part 0
  for (int c = threadIdx.x; c < in.c; c += blockDim.x) {
    for (int r = 12; r < 12 * PSP_PARTS; ++r) dst[(long long)r * in.c + c] = 0.f;
    int r = 0;
#pragma unroll
    for (int lv = 0; lv < 4; ++lv) {
      const int o = lv == 0 ? 1 : lv == 1 ? 2 : lv == 2 ? 3 : 6;
      for (int j = 0; j < o; ++j, ++r) {
        int x0 = bin_start(j, o, in.w), x1 = bin_end(j, o, in.w);
        float s = 0.f;
        for (int x = x0; x < x1; ++x) s += ld1(in, row + x * in.sw + c);
        dst[(long long)r * in.c + c] = s;
      }
    }
  }
}

// Pass 2.  A block per bin; BS_GROUPS thread groups walk interleaved rows of the bin (the 1x1 bin sums all H rows: one
// serial walk per channel was 14 us of pure load latency) and their partial sums are added in group order (fixed order).
// 1024 threads = 4 groups x 256 channels per pass, eight loads in flight per thread: the 1x1 bin of a 128-row map is 2 passes x
// 4 round trips to L2.  (The first version of this kernel had 512 threads = 4 groups x 128 channels and four loads in flight:
// 4 passes x 8 round trips, 24 us -- slower than the serial walk it replaced; found in the launch list of the third session.)
constexpr int BS_GROUPS = 4;
constexpr int BS_THREADS = 1024;
__global__ void psp_binsum_kernel(const float* __restrict__ rowsum, View out, int H, int W) {
  extern __shared__ float bs_part[];                 // [BS_GROUPS - 1][C]
  const int bin = blockIdx.x, b = blockIdx.y;
  int o, local, roff;
  if (bin < 1) { o = 1; local = bin; roff = 0; }
  else if (bin < 5) { o = 2; local = bin - 1; roff = 1; }
  else if (bin < 14) { o = 3; local = bin - 5; roff = 3; }
  else { o = 6; local = bin - 14; roff = 6; }
  const int i = local / o, j = local % o;
  const int y0 = bin_start(i, o, H), y1 = bin_end(i, o, H);
  const int x0 = bin_start(j, o, W), x1 = bin_end(j, o, W);
  const float inv = 1.f / (float)((y1 - y0) * (x1 - x0));
  const int C = out.c;
  const int per = blockDim.x / BS_GROUPS;            // threads per group
  const int grp = threadIdx.x / per, t = threadIdx.x - grp * per;
  for (int c0 = 0; c0 < C; c0 += per) {
    const int c = c0 + t;
    float s = 0.f;
    if (c < C) {
#pragma unroll 8
      for (int y = y0 + grp; y < y1; y += BS_GROUPS)
        s += rowsum[(((long long)(b * H + y) * PSP_PARTS) * 12 + roff + j) * C + c];
      if (grp > 0) bs_part[(grp - 1) * C + c] = s;
    }
    __syncthreads();
    if (grp == 0 && c < C) {
#pragma unroll
      for (int g2 = 0; g2 < BS_GROUPS - 1; ++g2) s += bs_part[g2 * C + c];
      out.p[b * out.sn + bin * out.sw + c] = s * inv;
    }
    __syncthreads();
  }
}

int psp_pool(const tdn_tensor* in, const tdn_tensor* out, float* workspace, size_t workspace_bytes,
             cudaStream_t stream) {
  int rc;
  if ((rc = check_tensor(in, "psp_pool.in"))) return rc;
  if ((rc = check_f32_tensor(out, "psp_pool.out"))) return rc;
  TDN_REQUIRE(out->n == in->n && out->h == 1 && out->w == 50 && out->c == in->c, TDN_ERR_INVALID,
              "psp_pool: out must be [n,1,50,c]");
  size_t need = (size_t)in->n * in->h * PSP_PARTS * 12 * in->c * sizeof(float);
  TDN_REQUIRE(workspace && workspace_bytes >= need, TDN_ERR_WORKSPACE,
              "psp_pool: workspace %zu < %zu bytes", workspace_bytes, need);
  int threads = in->c >= 512 ? 512 : (in->c >= 256 ? 256 : 128);
  const int seg_len = ceil_div(in->w, RS_SEG);
  const bool split = in->dtype == TDN_SPLIT16;
  const bool vec8 = in->c % 8 == 0 && 6 * (seg_len + 1) <= in->w &&
                    (split ? (aligned16(in->data) && aligned16(in->data_lo) && in->stride_n % 8 == 0 &&
                              in->stride_h % 8 == 0 && in->stride_w % 8 == 0)
                           : vec4_ok(*in));
  if (vec8) {
    const dim3 grid(in->h, ceil_div(in->c, RS_CB), in->n);
    if (split) psp_rowsum_vec8_kernel<true><<<grid, RS_SEG * RS_CB / 8, 0, stream>>>(make_view(*in), workspace, seg_len);
    else psp_rowsum_vec8_kernel<false><<<grid, RS_SEG * RS_CB / 8, 0, stream>>>(make_view(*in), workspace, seg_len);
  } else if (in->w >= 6) {
    psp_rowsum_kernel<<<dim3(in->h, ceil_div(in->c, 128), in->n * PSP_PARTS), 128, 0, stream>>>(make_view(*in),
                                                                                              workspace);
  } else {
    psp_rowsum_generic_kernel<<<dim3(in->h, in->n), threads, 0, stream>>>(make_view(*in), workspace);
  }
  TDN_LAUNCH_OK();
  psp_binsum_kernel<<<dim3(50, in->n), BS_THREADS, (BS_GROUPS - 1) * in->c * sizeof(float), stream>>>(workspace, make_view(*out),
                                                                                              in->h, in->w);
  TDN_LAUNCH_OK();
  return TDN_OK;
}

// ---------------------------------------------------------------------------------------------
// Bilinear resize, align_corners=True, NHWC -> NHWC (channel-slice views allowed).
// src = dst * (in-1)/(out-1); index0 = floor, lambda = frac (ATen upsample semantics).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void src_index(int dst, float scale, int in_size, int& i0, int& i1, float& l1) {
  float s = scale * (float)dst;
  i0 = min((int)s, in_size - 1);
  i1 = min(i0 + 1, in_size - 1);
  l1 = fminf(fmaxf(s - (float)i0, 0.f), 1.f);
}

__global__ void bilinear_nhwc_kernel(View in, View out, float sy, float sx) {
  long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  long long total = (long long)out.n * out.h * out.w * out.c;
  if (idx >= total) return;
  int c = idx % out.c;
  long long t = idx / out.c;
  int x = t % out.w; t /= out.w;
  int y = t % out.h;
  int b = t / out.h;
  int y0, y1, x0, x1; float ly, lx;
  src_index(y, sy, in.h, y0, y1, ly);
  src_index(x, sx, in.w, x0, x1, lx);
  const long long base = b * in.sn + c;
  float v00 = ld1(in, base + y0 * in.sh + x0 * in.sw), v01 = ld1(in, base + y0 * in.sh + x1 * in.sw);
  float v10 = ld1(in, base + y1 * in.sh + x0 * in.sw), v11 = ld1(in, base + y1 * in.sh + x1 * in.sw);
  float top = v00 * (1.f - lx) + v01 * lx;
  float bot = v10 * (1.f - lx) + v11 * lx;
  st1(out, b * out.sn + y * out.sh + x * out.sw + c, top * (1.f - ly) + bot * ly);
}

// Same arithmetic, four channels per thread (16-byte loads / stores) for vector-aligned views.
__global__ void bilinear_nhwc_vec4_kernel(View in, View out, float sy, float sx) {
  const int c4n = out.c >> 2;
  long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  long long total = (long long)out.n * out.h * out.w * c4n;
  if (idx >= total) return;
  int c = (int)(idx % c4n) * 4;
  long long t = idx / c4n;
  int x = t % out.w; t /= out.w;
  int y = t % out.h;
  int b = t / out.h;
  int y0, y1, x0, x1; float ly, lx;
  src_index(y, sy, in.h, y0, y1, ly);
  src_index(x, sx, in.w, x0, x1, lx);
  const long long base = b * in.sn + c;
  const float4 v00 = ld4(in, base + y0 * in.sh + x0 * in.sw), v01 = ld4(in, base + y0 * in.sh + x1 * in.sw);
  const float4 v10 = ld4(in, base + y1 * in.sh + x0 * in.sw), v11 = ld4(in, base + y1 * in.sh + x1 * in.sw);
  float4 o;
#define TDN_BL(m) { float top = v00.m * (1.f - lx) + v01.m * lx; float bot = v10.m * (1.f - lx) + v11.m * lx; \
                    o.m = top * (1.f - ly) + bot * ly; }
  TDN_BL(x) TDN_BL(y) TDN_BL(z) TDN_BL(w)
#undef TDN_BL
  st4(out, b * out.sn + y * out.sh + x * out.sw + c, o);
}

static inline float ac_scale(int in_size, int out_size) {
  return out_size > 1 ? (float)(in_size - 1) / (float)(out_size - 1) : 0.f;
}

int bilinear_nhwc(const tdn_tensor* in, const tdn_tensor* out, cudaStream_t stream) {
  int rc;
  if ((rc = check_tensor(in, "bilinear.in"))) return rc;
  if ((rc = check_tensor(out, "bilinear.out"))) return rc;
  TDN_REQUIRE(in->n == out->n && in->c == out->c, TDN_ERR_INVALID, "bilinear: n/c mismatch");
  if (vec4_ok(*in) && vec4_ok(*out)) {
    long long total4 = (long long)out->n * out->h * out->w * (out->c / 4);
    bilinear_nhwc_vec4_kernel<<<ceil_div(total4, 256), 256, 0, stream>>>(
        make_view(*in), make_view(*out), ac_scale(in->h, out->h), ac_scale(in->w, out->w));
    TDN_LAUNCH_OK();
    return TDN_OK;
  }
  long long total = (long long)out->n * out->h * out->w * out->c;
  bilinear_nhwc_kernel<<<ceil_div(total, 256), 256, 0, stream>>>(
      make_view(*in), make_view(*out), ac_scale(in->h, out->h), ac_scale(in->w, out->w));
  TDN_LAUNCH_OK();
  return TDN_OK;
}

// ---------------------------------------------------------------------------------------------
// The four PSP branch convolutions (conv1x1 c4 -> slice of c4/4, folded BN, ReLU; td4_psp18.py:255-266)
// on the 50 pooled bins in ONE launch: a block per (bin, image) keeps the pooled vector in shared memory,
// each warp produces output channels with a coalesced dot product over K and a shuffle reduction.
// (As four generic conv launches these were single-CTA, latency-bound ~45 us each.)
// ---------------------------------------------------------------------------------------------
struct PspBranch {
  const float* w[4];      // [eighth][c4] fp32, K-major
  const float* scale[4];
  const float* bias[4];
  float* out[4];          // [n, bins, bins, eighth]
};

// Optional second stage (tdn_psp_branch_project): the consumers of z = cat(c4 slice, up(b1), .., up(b6)) are 1x1
// convolutions, and a 1x1 convolution commutes with the bilinear resize: W . up(b_i) = up(W_i . b_i).  With the
// interpolation weights of the 50 bins as 64 extra input channels of every pixel (constant per map size, written once
// by the host), the resized branch maps never have to exist: this stage writes T[bin][o] = sum_c W[o][lv*eighth + c] *
// b_lv[bin][c] into column `bin` of the consumer's K-major SPLIT16 weight matrix, and the consumer reads
// [c4 slice | 64 interpolation channels] as its input (td4_psp18.py:273-284 + transformer.py:53-55 in one GEMM).
struct PspProjection {
  const float* w;         // [4 * eighth][cout] fp32 (already carrying the row scale of the consumer's static weights)
  __half* hi;
  __half* lo;
  long long ld, bs;       // row pitch / per-image stride of the destination (elements)
  int cout;
};
struct PspProjections {
  PspProjection p[TDN_PSP_MAX_PROJECTIONS];
  int n;
  int* range_flag;
};

constexpr int PB_THREADS = 512;
__global__ void __launch_bounds__(PB_THREADS) psp_branch_kernel(const float* __restrict__ pooled, int c4, int eighth,
                                                                PspBranch br, PspProjections pj) {
  extern __shared__ float xs[];         // pooled vector [c4] | branch output of this bin [eighth]
  float* ys = xs + c4;
  const int bin = blockIdx.x, b = blockIdx.y;
  int lv, local, bins;
  if (bin < 1) { lv = 0; local = bin; bins = 1; }
  else if (bin < 5) { lv = 1; local = bin - 1; bins = 2; }
  else if (bin < 14) { lv = 2; local = bin - 5; bins = 3; }
  else { lv = 3; local = bin - 14; bins = 6; }
  const float* x = pooled + ((long long)b * 50 + bin) * c4;
  for (int k = threadIdx.x; k < c4; k += PB_THREADS) xs[k] = x[k];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* w = br.w[lv];
  float* out = br.out[lv] + ((long long)b * bins * bins + local) * eighth;
  for (int co = warp; co < eighth; co += PB_THREADS / 32) {
    const float* wr = w + (long long)co * c4;
    float acc = 0.f;
    for (int k = lane * 4; k < c4; k += 128) {
      const float4 wv = *reinterpret_cast<const float4*>(wr + k);
      const float4 xv = *reinterpret_cast<const float4*>(xs + k);
      acc = fmaf(wv.x, xv.x, acc); acc = fmaf(wv.y, xv.y, acc);
      acc = fmaf(wv.z, xv.z, acc); acc = fmaf(wv.w, xv.w, acc);
    }
    acc = warp_sum_f(acc);
    if (lane == 0) {
      const float v = fmaxf(fmaf(acc, br.scale[lv][co], br.bias[lv][co]), 0.f);
      out[co] = v;
      ys[co] = v;
    }
  }
  if (pj.n == 0) return;
  __syncthreads();
  // One output channel per thread over the concatenated outputs of all projections: the loads of a warp are one
  // coalesced line per input channel and 16 of them are in flight per thread.  (A warp per output with a shuffle
  // reduction measured 100 us here -- 80 dependent round trips to L2 per warp; the projections one after the other
  // with 8 loads in flight 59 us.)
  bool out_of_range = false;
  int total = 0;
  for (int q = 0; q < pj.n; ++q) total += pj.p[q].cout;
  for (int t = threadIdx.x; t < total; t += PB_THREADS) {
    int q = 0, o = t;
    while (o >= pj.p[q].cout) { o -= pj.p[q].cout; ++q; }
    const PspProjection& P = pj.p[q];
    const float* wc = P.w + (long long)lv * eighth * P.cout + o;
    float acc = 0.f;
#pragma unroll 16
    for (int k = 0; k < eighth; ++k) acc = fmaf(__ldg(wc + (long long)k * P.cout), ys[k], acc);   // index order: fixed
    __half h, l;
    split_f32(acc, h, l);
    const long long d = (long long)b * P.bs + (long long)o * P.ld + bin;
    P.hi[d] = h;
    P.lo[d] = l;
    out_of_range |= fabsf(acc) > 60000.f;
  }
  if (out_of_range && pj.range_flag) *reinterpret_cast<volatile int*>(pj.range_flag) = 1;
}

static int psp_branch_launch(const tdn_tensor* pooled, const float* const* w, const float* const* scale,
                             const float* const* bias, int eighth, float* const* out, const PspProjections& pj,
                             cudaStream_t stream) {
  int rc;
  if ((rc = check_f32_tensor(pooled, "psp_branch.pooled"))) return rc;
  TDN_REQUIRE(pooled->h == 1 && pooled->w == 50 && pooled->stride_w == pooled->c && pooled->stride_n == 50ll * pooled->c &&
                  pooled->c % 4 == 0 && aligned16(pooled->data), TDN_ERR_INVALID, "psp_branch: pooled must be dense [n,1,50,c4]");
  TDN_REQUIRE(eighth > 0, TDN_ERR_INVALID, "psp_branch: eighth must be positive");
  PspBranch br;
  for (int i = 0; i < 4; ++i) {
    TDN_REQUIRE(w && scale && bias && out && w[i] && scale[i] && bias[i] && out[i] && aligned16(w[i]), TDN_ERR_INVALID,
                "psp_branch: null or misaligned branch %d", i);
    br.w[i] = w[i]; br.scale[i] = scale[i]; br.bias[i] = bias[i]; br.out[i] = out[i];
  }
  psp_branch_kernel<<<dim3(50, pooled->n), PB_THREADS, (pooled->c + eighth) * sizeof(float), stream>>>(
      (const float*)pooled->data, pooled->c, eighth, br, pj);
  TDN_LAUNCH_OK();
  return TDN_OK;
}

int psp_branch_convs(const tdn_tensor* pooled, const float* const* w, const float* const* scale,
                     const float* const* bias, int eighth, float* const* out, cudaStream_t stream) {
  PspProjections pj;
  pj.n = 0;
  pj.range_flag = nullptr;
  return psp_branch_launch(pooled, w, scale, bias, eighth, out, pj, stream);
}

int psp_branch_project(const tdn_tensor* pooled, const float* const* w, const float* const* scale,
                       const float* const* bias, int eighth, float* const* out, const tdn_psp_projection* proj,
                       int n_proj, int* range_flag, cudaStream_t stream) {
  TDN_REQUIRE(proj && n_proj >= 1 && n_proj <= TDN_PSP_MAX_PROJECTIONS, TDN_ERR_INVALID,
              "psp_branch_project: 1..%d projections expected", TDN_PSP_MAX_PROJECTIONS);
  PspProjections pj;
  pj.n = n_proj;
  pj.range_flag = range_flag;
  for (int q = 0; q < n_proj; ++q) {
    const tdn_psp_projection& s = proj[q];
    TDN_REQUIRE(s.w && s.dst_hi && s.dst_lo && s.cout > 0 && s.ld >= 50 && (pooled == nullptr || pooled->n == 1 || s.batch_stride > 0),
                TDN_ERR_INVALID, "psp_branch_project: bad projection %d", q);
    pj.p[q].w = s.w;
    pj.p[q].hi = (__half*)s.dst_hi;
    pj.p[q].lo = (__half*)s.dst_lo;
    pj.p[q].ld = s.ld;
    pj.p[q].bs = s.batch_stride;
    pj.p[q].cout = s.cout;
  }
  return psp_branch_launch(pooled, w, scale, bias, eighth, out, pj, stream);
}

// ---------------------------------------------------------------------------------------------
// PyramidPooling output in one pass (td4_psp18.py:273-284): z = cat(x[:, slice], up(feat1..4)[:, slice]).
// `x` is already the channel-slice view of c4; the four branch maps are the (sliced) conv+BN+ReLU
// outputs on the 1/2/3/6 grids and are bilinearly resized (align_corners) on the fly.
// ---------------------------------------------------------------------------------------------
struct PspSmall { const float* p[4]; };

__global__ void psp_concat_kernel(View x, PspSmall small, int eighth, View z) {
  const int c4 = z.c >> 2;
  long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  long long total = (long long)z.n * z.h * z.w * c4;
  if (idx >= total) return;
  const int c = (idx % c4) * 4;
  long long t = idx / c4;
  const int xx = t % z.w; t /= z.w;
  const int y = t % z.h;
  const int b = t / z.h;
  const long long zo = b * z.sn + y * z.sh + xx * z.sw + c;
  if (c < x.c) {
    st4(z, zo, ld4(x, b * x.sn + y * x.sh + xx * x.sw + c));
    return;
  }
  const int lv = (c - x.c) / eighth;
  const int cc = (c - x.c) - lv * eighth;
  const int bins = lv == 0 ? 1 : lv == 1 ? 2 : lv == 2 ? 3 : 6;
  int y0, y1, x0, x1; float ly, lx;
  src_index(y, z.h > 1 ? (float)(bins - 1) / (float)(z.h - 1) : 0.f, bins, y0, y1, ly);
  src_index(xx, z.w > 1 ? (float)(bins - 1) / (float)(z.w - 1) : 0.f, bins, x0, x1, lx);
  const float* base = small.p[lv] + ((long long)b * bins * bins) * eighth + cc;
  const float4 v00 = *reinterpret_cast<const float4*>(base + (y0 * bins + x0) * eighth);
  const float4 v01 = *reinterpret_cast<const float4*>(base + (y0 * bins + x1) * eighth);
  const float4 v10 = *reinterpret_cast<const float4*>(base + (y1 * bins + x0) * eighth);
  const float4 v11 = *reinterpret_cast<const float4*>(base + (y1 * bins + x1) * eighth);
  float4 o;
  o.x = (v00.x * (1.f - lx) + v01.x * lx) * (1.f - ly) + (v10.x * (1.f - lx) + v11.x * lx) * ly;
  o.y = (v00.y * (1.f - lx) + v01.y * lx) * (1.f - ly) + (v10.y * (1.f - lx) + v11.y * lx) * ly;
  o.z = (v00.z * (1.f - lx) + v01.z * lx) * (1.f - ly) + (v10.z * (1.f - lx) + v11.z * lx) * ly;
  o.w = (v00.w * (1.f - lx) + v01.w * lx) * (1.f - ly) + (v10.w * (1.f - lx) + v11.w * lx) * ly;
  st4(z, zo, o);
}

int psp_concat(const tdn_tensor* x, const float* const* small, int eighth, const tdn_tensor* z, cudaStream_t stream) {
  int rc;
  if ((rc = check_tensor(x, "psp_concat.x"))) return rc;
  if ((rc = check_tensor(z, "psp_concat.z"))) return rc;
  TDN_REQUIRE(small && small[0] && small[1] && small[2] && small[3], TDN_ERR_INVALID, "psp_concat: null branch map");
  TDN_REQUIRE(eighth > 0 && eighth % 4 == 0 && z->c == x->c + 4 * eighth && x->n == z->n && x->h == z->h &&
                  x->w == z->w && vec4_ok(*x) && vec4_ok(*z), TDN_ERR_INVALID,
              "psp_concat: z must be [n,h,w,x.c + 4*eighth] with vector-aligned views");
  for (int i = 0; i < 4; ++i) TDN_REQUIRE(aligned16(small[i]), TDN_ERR_INVALID, "psp_concat: misaligned branch map");
  PspSmall sm;
  for (int i = 0; i < 4; ++i) sm.p[i] = small[i];
  long long total = (long long)z->n * z->h * z->w * (z->c / 4);
  psp_concat_kernel<<<ceil_div(total, 256), 256, 0, stream>>>(make_view(*x), sm, eighth, make_view(*z));
  TDN_LAUNCH_OK();
  return TDN_OK;
}

// ---------------------------------------------------------------------------------------------
// Strided NHWC copy (float4 when possible).
// ---------------------------------------------------------------------------------------------
template <int VEC>
__global__ void copy_nhwc_kernel(View in, View out) {
  const int cv = out.c / VEC;
  long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  long long total = (long long)out.n * out.h * out.w * cv;
  if (idx >= total) return;
  int c = (idx % cv) * VEC;
  long long t = idx / cv;
  int x = t % out.w; t /= out.w;
  int y = t % out.h;
  int b = t / out.h;
  const long long so = b * in.sn + y * in.sh + x * in.sw + c;
  const long long dof = b * out.sn + y * out.sh + x * out.sw + c;
  if (VEC == 4) st4(out, dof, ld4(in, so));
  else st1(out, dof, ld1(in, so));
}

int copy_nhwc(const tdn_tensor* in, const tdn_tensor* out, cudaStream_t stream) {
  int rc;
  if ((rc = check_tensor(in, "copy.in"))) return rc;
  if ((rc = check_tensor(out, "copy.out"))) return rc;
  TDN_REQUIRE(in->n == out->n && in->h == out->h && in->w == out->w && in->c == out->c,
              TDN_ERR_INVALID, "copy_nhwc: dims mismatch");
  if (vec4_ok(*in) && vec4_ok(*out)) {
    long long total = (long long)out->n * out->h * out->w * (out->c / 4);
    copy_nhwc_kernel<4><<<ceil_div(total, 256), 256, 0, stream>>>(make_view(*in), make_view(*out));
  } else {
    long long total = (long long)out->n * out->h * out->w * out->c;
    copy_nhwc_kernel<1><<<ceil_div(total, 256), 256, 0, stream>>>(make_view(*in), make_view(*out));
  }
  TDN_LAUNCH_OK();
  return TDN_OK;
}

// fp32 <-> SPLIT16 conversions are copies between views of different dtype.
int split16(const tdn_tensor* in, const tdn_tensor* out, cudaStream_t stream) {
  TDN_REQUIRE(in && out && in->dtype == TDN_F32 && out->dtype == TDN_SPLIT16, TDN_ERR_INVALID,
              "split16: expects an F32 input and a SPLIT16 output");
  return copy_nhwc(in, out, stream);
}
int merge16(const tdn_tensor* in, const tdn_tensor* out, cudaStream_t stream) {
  TDN_REQUIRE(in && out && in->dtype == TDN_SPLIT16 && out->dtype == TDN_F32, TDN_ERR_INVALID,
              "merge16: expects a SPLIT16 input and an F32 output");
  return copy_nhwc(in, out, stream);
}

// ---------------------------------------------------------------------------------------------
// Row softmax with pre-scale (one CTA per row; three passes over a row that stays in L1/L2).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void __launch_bounds__(256) softmax_rows_kernel(float* __restrict__ s, int cols,
                                                           long long ld, float scale) {
  __shared__ float red[8];
  float* row = s + blockIdx.x * ld;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  float m = -INFINITY;
  for (int c = threadIdx.x; c < cols; c += 256) m = fmaxf(m, row[c] * scale);
  m = warp_max(m);
  if (lane == 0) red[wid] = m;
  __syncthreads();
  m = red[0];
#pragma unroll
  for (int i = 1; i < 8; ++i) m = fmaxf(m, red[i]);
  __syncthreads();
  float sum = 0.f;
  for (int c = threadIdx.x; c < cols; c += 256) {
    float e = expf(row[c] * scale - m);
    row[c] = e;
    sum += e;
  }
  sum = warp_sum(sum);
  if (lane == 0) red[wid] = sum;
  __syncthreads();
  sum = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) sum += red[i];
  const float inv = 1.f / sum;
  for (int c = threadIdx.x; c < cols; c += 256) row[c] *= inv;
}

// Same softmax, but the probabilities leave as SPLIT16 (times `out_scale`, a power of two that keeps the
// lo plane in the normal fp16 range; the consumer GEMM undoes it in its epilogue scale).  Columns
// [cols, ld_out) are written as zeros so that the row can be used as a K operand padded to 64.
__global__ void __launch_bounds__(256) softmax_rows_split16_kernel(const float* __restrict__ s, int cols,
                                                                   long long ld, float scale,
                                                                   __half* __restrict__ p_hi,
                                                                   __half* __restrict__ p_lo, long long ld_out,
                                                                   float out_scale) {
  __shared__ float red[8];
  const float* row = s + blockIdx.x * ld;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  float m = -INFINITY;
  for (int c = threadIdx.x; c < cols; c += 256) m = fmaxf(m, row[c] * scale);
  m = warp_max(m);
  if (lane == 0) red[wid] = m;
  __syncthreads();
  m = red[0];
#pragma unroll
  for (int i = 1; i < 8; ++i) m = fmaxf(m, red[i]);
  __syncthreads();
  float sum = 0.f;
  for (int c = threadIdx.x; c < cols; c += 256) sum += expf(row[c] * scale - m);
  sum = warp_sum(sum);
  if (lane == 0) red[wid] = sum;
  __syncthreads();
  sum = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) sum += red[i];
  const float inv = out_scale / sum;
  __half* oh = p_hi + blockIdx.x * ld_out;
  __half* ol = p_lo + blockIdx.x * ld_out;
  for (int c = threadIdx.x; c < ld_out; c += 256) {
    float v = c < cols ? expf(row[c] * scale - m) * inv : 0.f;
    __half h, l;
    split_f32(v, h, l);
    oh[c] = h;
    ol[c] = l;
  }
}

int softmax_rows_split16(const float* s, long long rows, int cols, long long ld, float scale, void* p_hi,
                         void* p_lo, long long ld_out, float out_scale, cudaStream_t stream) {
  TDN_REQUIRE(s && p_hi && p_lo && rows > 0 && cols > 0 && ld >= cols && ld_out >= cols, TDN_ERR_INVALID,
              "softmax_rows_split16: bad arguments");
  TDN_REQUIRE(rows < (1ll << 31), TDN_ERR_UNSUPPORTED, "softmax_rows_split16: too many rows");
  softmax_rows_split16_kernel<<<(unsigned)rows, 256, 0, stream>>>(s, cols, ld, scale, (__half*)p_hi, (__half*)p_lo,
                                                                  ld_out, out_scale);
  TDN_LAUNCH_OK();
  return TDN_OK;
}

int softmax_rows(float* s, long long rows, int cols, long long ld, float scale, cudaStream_t stream) {
  TDN_REQUIRE(s && rows > 0 && cols > 0 && ld >= cols, TDN_ERR_INVALID, "softmax_rows: bad arguments");
  TDN_REQUIRE(rows < (1ll << 31), TDN_ERR_UNSUPPORTED, "softmax_rows: too many rows");
  softmax_rows_kernel<<<(unsigned)rows, 256, 0, stream>>>(s, cols, ld, scale);
  TDN_LAUNCH_OK();
  return TDN_OK;
}

// ---------------------------------------------------------------------------------------------
// LayerNorm over the (H, W) map per (n, channel): fixed-order two-stage reduction in fp64 (bit
// reproducible run to run; no atomics), then a fused normalise + spatial affine.
// ---------------------------------------------------------------------------------------------
constexpr int LN_CHUNK = 64;  // pixels per partial

__global__ void ln_partial_kernel(View x, double2* __restrict__ part, int chunks) {
  const int chunk = blockIdx.x, b = blockIdx.y;
  const int P = x.h * x.w;
  const int p0 = chunk * LN_CHUNK, p1 = min(p0 + LN_CHUNK, P);
  for (int c = threadIdx.x; c < x.c; c += blockDim.x) {
    double s = 0.0, q = 0.0;
    for (int p = p0; p < p1; ++p) {
      int y = p / x.w, xx = p - y * x.w;
      double v = (double)ld1(x, b * x.sn + y * x.sh + xx * x.sw + c);
      s += v; q += v * v;
    }
    part[((long long)b * chunks + chunk) * x.c + c] = make_double2(s, q);
  }
}

// Same partial sums (same order per channel, so bit-identical results), four channels per thread with 8- / 16-byte
// loads for vector-aligned views: the scalar version issues one 2-byte load per plane and element.
__global__ void ln_partial_vec4_kernel(View x, double2* __restrict__ part, int chunks) {
  const int chunk = blockIdx.x, b = blockIdx.y;
  const int P = x.h * x.w;
  const int p0 = chunk * LN_CHUNK, p1 = min(p0 + LN_CHUNK, P);
  for (int c = threadIdx.x * 4; c < x.c; c += blockDim.x * 4) {
    double s[4] = {0.0, 0.0, 0.0, 0.0}, q[4] = {0.0, 0.0, 0.0, 0.0};
    int y = p0 / x.w, xx = p0 - y * x.w;
#pragma unroll 4
    for (int p = p0; p < p1; ++p) {
      const float4 v = ld4(x, b * x.sn + y * x.sh + xx * x.sw + c);
      const double d0 = (double)v.x, d1 = (double)v.y, d2 = (double)v.z, d3 = (double)v.w;
      s[0] += d0; q[0] += d0 * d0; s[1] += d1; q[1] += d1 * d1;
      s[2] += d2; q[2] += d2 * d2; s[3] += d3; q[3] += d3 * d3;
      if (++xx == x.w) { xx = 0; ++y; }
    }
    double2* dst = part + ((long long)b * chunks + chunk) * x.c + c;
#pragma unroll
    for (int j = 0; j < 4; ++j) dst[j] = make_double2(s[j], q[j]);
  }
}

// SPLIT16 maps with c % 8 == 0: 8 channels per thread (one 16-byte load per plane and pixel) and four pixel sub-groups
// per chunk, i.e. 4x the warps and 2x the bytes per load of the float4 version (31 us for the 67 MB map of a 1024x2048
// frame: latency-bound).  The sub-group sums are added in index order through shared memory (fixed order per channel:
// bit-reproducible run to run).
constexpr int LN_SUB = 4;
__global__ void __launch_bounds__(64 * LN_SUB) ln_partial_split8_kernel(View x, double2* __restrict__ part, int chunks) {
  __shared__ double2 sm[LN_SUB - 1][512];
  const int chunk = blockIdx.x, b = blockIdx.y;
  const int P = x.h * x.w;
  const int sub = threadIdx.x >> 6, lane_g = threadIdx.x & 63;
  constexpr int PER = LN_CHUNK / LN_SUB;
  const int p0 = min(chunk * LN_CHUNK + sub * PER, P), p1 = min(p0 + PER, P);
  for (int cb = 0; cb < x.c; cb += 512) {
    const int c = cb + lane_g * 8;
    double s[8], q[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) { s[e] = 0.0; q[e] = 0.0; }
    if (c < x.c) {
      int y = p0 / x.w, xx = p0 - y * x.w;
#pragma unroll 4
      for (int p = p0; p < p1; ++p) {
        const long long off = b * x.sn + y * x.sh + xx * x.sw + c;
        const uint4 hv = *reinterpret_cast<const uint4*>(x.hi + off);
        const uint4 lv = *reinterpret_cast<const uint4*>(x.lo + off);
        const __half2* h = reinterpret_cast<const __half2*>(&hv);
        const __half2* l = reinterpret_cast<const __half2*>(&lv);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 a = __half22float2(h[e]), d = __half22float2(l[e]);
          const double d0 = (double)(a.x + d.x), d1 = (double)(a.y + d.y);
          s[2 * e] += d0; q[2 * e] += d0 * d0;
          s[2 * e + 1] += d1; q[2 * e + 1] += d1 * d1;
        }
        if (++xx == x.w) { xx = 0; ++y; }
      }
    }
    if (sub > 0) {
#pragma unroll
      for (int e = 0; e < 8; ++e) sm[sub - 1][lane_g * 8 + e] = make_double2(s[e], q[e]);
    }
    __syncthreads();
    if (sub == 0 && c < x.c) {
      double2* dst = part + ((long long)b * chunks + chunk) * x.c + c;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        double ss = s[e], qq = q[e];
#pragma unroll
        for (int g = 0; g < LN_SUB - 1; ++g) { ss += sm[g][lane_g * 8 + e].x; qq += sm[g][lane_g * 8 + e].y; }
        dst[e] = make_double2(ss, qq);
      }
    }
    __syncthreads();
  }
}

constexpr int LN_FINAL_GROUPS = 32;   // chunk groups per block: the kernel is a chain of dependent loads per thread, so
                                      // more (shorter) chains finish sooner; the order of the additions is fixed
__global__ void __launch_bounds__(32 * LN_FINAL_GROUPS) ln_final_kernel(const double2* __restrict__ part, int chunks, int C,
                                                                        int P, float eps, float* __restrict__ mean,
                                                                        float* __restrict__ rstd) {
  // block = 32 channels x LN_FINAL_GROUPS chunk groups; fixed summation order -> bit-reproducible
  __shared__ double ss[LN_FINAL_GROUPS][32], sq[LN_FINAL_GROUPS][32];
  const int b = blockIdx.y;
  const int cl = threadIdx.x & 31, grp = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cl;
  double s = 0.0, q = 0.0;
  if (c < C) {
    for (int k = grp; k < chunks; k += LN_FINAL_GROUPS) {
      double2 v = part[((long long)b * chunks + k) * C + c];
      s += v.x; q += v.y;
    }
  }
  ss[grp][cl] = s; sq[grp][cl] = q;
  __syncthreads();
  if (grp == 0 && c < C) {
#pragma unroll
    for (int g = 1; g < LN_FINAL_GROUPS; ++g) { s += ss[g][cl]; q += sq[g][cl]; }
    double mu = s / P;
    double var = q / P - mu * mu;
    if (var < 0.0) var = 0.0;
    mean[b * C + c] = (float)mu;
    rstd[b * C + c] = (float)(1.0 / sqrt(var + (double)eps));
  }
}

int layernorm_hw_stats(const tdn_tensor* x, float* mean, float* rstd, float eps, void* workspace,
                       size_t workspace_bytes, cudaStream_t stream) {
  int rc;
  if ((rc = check_tensor(x, "layernorm.x"))) return rc;
  TDN_REQUIRE(mean && rstd, TDN_ERR_INVALID, "layernorm_hw_stats: null output");
  const int P = x->h * x->w;
  const int chunks = ceil_div(P, LN_CHUNK);
  size_t need = (size_t)x->n * chunks * x->c * sizeof(double2);
  TDN_REQUIRE(workspace && workspace_bytes >= need && aligned16(workspace), TDN_ERR_WORKSPACE,
              "layernorm_hw_stats: workspace %zu < %zu bytes", workspace_bytes, need);
  const bool split8 = x->dtype == TDN_SPLIT16 && x->c % 8 == 0 && x->c >= 256 && aligned16(x->data) && aligned16(x->data_lo) &&
                      x->stride_n % 8 == 0 && x->stride_h % 8 == 0 && x->stride_w % 8 == 0;
  if (split8) {
    ln_partial_split8_kernel<<<dim3(chunks, x->n), 64 * LN_SUB, 0, stream>>>(make_view(*x), (double2*)workspace, chunks);
  } else if (vec4_ok(*x)) {
    int threads = x->c / 4 >= 256 ? 256 : (x->c / 4 >= 128 ? 128 : 64);
    ln_partial_vec4_kernel<<<dim3(chunks, x->n), threads, 0, stream>>>(make_view(*x), (double2*)workspace, chunks);
  } else {
    int threads = x->c >= 256 ? 256 : 128;
    ln_partial_kernel<<<dim3(chunks, x->n), threads, 0, stream>>>(make_view(*x), (double2*)workspace, chunks);
  }
  TDN_LAUNCH_OK();
  ln_final_kernel<<<dim3(ceil_div(x->c, 32), x->n), 32 * LN_FINAL_GROUPS, 0, stream>>>((const double2*)workspace, chunks, x->c, P,
                                                                     eps, mean, rstd);
  TDN_LAUNCH_OK();
  return TDN_OK;
}

// SPLIT16 in and out, c % 8 == 0: 8 channels per thread with 16-byte accesses on every plane (same arithmetic).
__global__ void ln_apply_split8_kernel(View x, View out, const float* __restrict__ mean,
                                       const float* __restrict__ rstd, const float* __restrict__ gamma,
                                       const float* __restrict__ beta) {
  const int c8 = x.c >> 3;
  long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  long long total = (long long)x.n * x.h * x.w * c8;
  if (idx >= total) return;
  const int c = (idx % c8) * 8;
  long long t = idx / c8;
  const int xx = t % x.w; t /= x.w;
  const int y = t % x.h;
  const int b = t / x.h;
  const float g = gamma[y * x.w + xx], be = beta[y * x.w + xx];
  const long long ioff = b * x.sn + y * x.sh + xx * x.sw + c;
  const uint4 hv = *reinterpret_cast<const uint4*>(x.hi + ioff);
  const uint4 lv = *reinterpret_cast<const uint4*>(x.lo + ioff);
  const __half2* h = reinterpret_cast<const __half2*>(&hv);
  const __half2* l = reinterpret_cast<const __half2*>(&lv);
  float mu[8], rs[8];
  *reinterpret_cast<float4*>(mu) = *reinterpret_cast<const float4*>(mean + b * x.c + c);
  *reinterpret_cast<float4*>(mu + 4) = *reinterpret_cast<const float4*>(mean + b * x.c + c + 4);
  *reinterpret_cast<float4*>(rs) = *reinterpret_cast<const float4*>(rstd + b * x.c + c);
  *reinterpret_cast<float4*>(rs + 4) = *reinterpret_cast<const float4*>(rstd + b * x.c + c + 4);
  __half2 oh[4], ol[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float2 a = __half22float2(h[e]), d = __half22float2(l[e]);
    const float v0 = a.x + d.x, v1 = a.y + d.y;
    split_f32x2((v0 - mu[2 * e]) * rs[2 * e] * g + be, (v1 - mu[2 * e + 1]) * rs[2 * e + 1] * g + be, oh[e], ol[e]);
  }
  const long long ooff = b * out.sn + y * out.sh + xx * out.sw + c;
  *reinterpret_cast<uint4*>(out.hi + ooff) = *reinterpret_cast<const uint4*>(oh);
  *reinterpret_cast<uint4*>(out.lo + ooff) = *reinterpret_cast<const uint4*>(ol);
}

__global__ void ln_apply_kernel(View x, View out, const float* __restrict__ mean,
                                const float* __restrict__ rstd, const float* __restrict__ gamma,
                                const float* __restrict__ beta) {
  const int c4 = x.c >> 2;
  long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  long long total = (long long)x.n * x.h * x.w * c4;
  if (idx >= total) return;
  int c = (idx % c4) * 4;
  long long t = idx / c4;
  int xx = t % x.w; t /= x.w;
  int y = t % x.h;
  int b = t / x.h;
  const float g = gamma[y * x.w + xx], be = beta[y * x.w + xx];
  float4 v = ld4(x, b * x.sn + y * x.sh + xx * x.sw + c);
  float4 mu = *reinterpret_cast<const float4*>(mean + b * x.c + c);
  float4 rs = *reinterpret_cast<const float4*>(rstd + b * x.c + c);
  float4 o;
  o.x = (v.x - mu.x) * rs.x * g + be;
  o.y = (v.y - mu.y) * rs.y * g + be;
  o.z = (v.z - mu.z) * rs.z * g + be;
  o.w = (v.w - mu.w) * rs.w * g + be;
  st4(out, b * out.sn + y * out.sh + xx * out.sw + c, o);
}

int layernorm_hw_apply(const tdn_tensor* x, const float* mean, const float* rstd, const float* gamma,
                       const float* beta, const tdn_tensor* out, cudaStream_t stream) {
  int rc;
  if ((rc = check_tensor(x, "layernorm.x"))) return rc;
  if ((rc = check_tensor(out, "layernorm.out"))) return rc;
  TDN_REQUIRE(mean && rstd && gamma && beta, TDN_ERR_INVALID, "layernorm_hw_apply: null pointer");
  TDN_REQUIRE(vec4_ok(*x) && vec4_ok(*out) && aligned16(mean) && aligned16(rstd), TDN_ERR_INVALID,
              "layernorm_hw_apply: float4-aligned views required");
  TDN_REQUIRE(x->n == out->n && x->h == out->h && x->w == out->w && x->c == out->c, TDN_ERR_INVALID,
              "layernorm_hw_apply: dims mismatch");
  long long total = (long long)x->n * x->h * x->w * (x->c / 4);
  const bool split8 = x->dtype == TDN_SPLIT16 && out->dtype == TDN_SPLIT16 && x->c % 8 == 0 && aligned16(x->data) &&
                      aligned16(x->data_lo) && aligned16(out->data) && aligned16(out->data_lo) &&
                      x->stride_n % 8 == 0 && x->stride_h % 8 == 0 && x->stride_w % 8 == 0 &&
                      out->stride_n % 8 == 0 && out->stride_h % 8 == 0 && out->stride_w % 8 == 0;
  if (split8)
    ln_apply_split8_kernel<<<ceil_div(total / 2, 256), 256, 0, stream>>>(make_view(*x), make_view(*out), mean, rstd,
                                                                        gamma, beta);
  else
    ln_apply_kernel<<<ceil_div(total, 256), 256, 0, stream>>>(make_view(*x), make_view(*out), mean, rstd,
                                                              gamma, beta);
  TDN_LAUNCH_OK();
  return TDN_OK;
}

// ---------------------------------------------------------------------------------------------
// 1x1 convolution to a handful of output channels (the nclass classifier, td4_psp18.py:299 / pspnet.py:113 /
// td2_fa.py:316): out[p][j] = (sum_c in[p][c] * w[j][c]) * scale[j] + bias[j], cout <= 32.  A 128x64-tile GEMM
// kernel wastes most of its tile on 19 columns and is latency-bound here (32 us at 128x256 pixels); this one stages
// 64 pixels x 64 channels per step in shared memory (coalesced 16-byte loads) and gives every thread one pixel and a
// group of ceil(cout / 4) classes, summing channels in index order (fixed order: bit-reproducible).
// ---------------------------------------------------------------------------------------------
constexpr int PL_PX = 64, PL_CK = 64, PL_MAXG = 8;

__global__ void __launch_bounds__(256) pointwise_linear_kernel(View in, const float* __restrict__ w,
                                                               const float* __restrict__ scale,
                                                               const float* __restrict__ bias, View out, int cout,
                                                               long long total_px) {
  __shared__ float sx[PL_PX][PL_CK + 1];
  __shared__ float sw[32][PL_CK + 1];
  const int tid = threadIdx.x;
  const long long p0 = (long long)blockIdx.x * PL_PX;
  const int px = tid & (PL_PX - 1), grp = tid >> 6;           // 64 pixels x 4 class groups
  const int cpg = (cout + 3) >> 2;                            // classes per group (<= 8)
  const int j0 = grp * cpg;
  float acc[PL_MAXG];
#pragma unroll
  for (int j = 0; j < PL_MAXG; ++j) acc[j] = 0.f;
  const int hw = in.h * in.w;
  for (int c0 = 0; c0 < in.c; c0 += PL_CK) {
    // stage: 64 px x 16 float4 -> four rounds; weights: cout x 16 float4
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int idx = tid + r * 256;
      const int sp = idx >> 4, q = idx & 15;
      const long long p = p0 + sp;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (p < total_px && c0 + q * 4 < in.c) {
        const int b = (int)(p / hw);
        const int rem = (int)(p - (long long)b * hw);
        const int y = rem / in.w, x = rem - y * in.w;
        v = ld4(in, b * in.sn + y * in.sh + x * in.sw + c0 + q * 4);
      }
      sx[sp][q * 4 + 0] = v.x; sx[sp][q * 4 + 1] = v.y; sx[sp][q * 4 + 2] = v.z; sx[sp][q * 4 + 3] = v.w;
    }
    for (int idx = tid; idx < cout * 16; idx += 256) {
      const int j = idx >> 4, q = idx & 15;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c0 + q * 4 < in.c) v = *reinterpret_cast<const float4*>(w + (long long)j * in.c + c0 + q * 4);
      sw[j][q * 4 + 0] = v.x; sw[j][q * 4 + 1] = v.y; sw[j][q * 4 + 2] = v.z; sw[j][q * 4 + 3] = v.w;
    }
    __syncthreads();
#pragma unroll 4
    for (int c = 0; c < PL_CK; ++c) {
      const float xv = sx[px][c];
#pragma unroll
      for (int j = 0; j < PL_MAXG; ++j)
        if (j < cpg && j0 + j < cout) acc[j] = fmaf(xv, sw[j0 + j][c], acc[j]);
    }
    __syncthreads();
  }
  const long long p = p0 + px;
  if (p >= total_px) return;
  const int b = (int)(p / hw);
  const int rem = (int)(p - (long long)b * hw);
  const int y = rem / out.w, x = rem - y * out.w;
  const long long o = b * out.sn + y * out.sh + x * out.sw;
#pragma unroll
  for (int j = 0; j < PL_MAXG; ++j) {
    const int jj = j0 + j;
    if (j < cpg && jj < cout) {
      float v = acc[j];
      if (scale) v *= scale[jj];
      if (bias) v += bias[jj];
      st1(out, o + jj, v);
    }
  }
}

int pointwise_linear(const tdn_tensor* in, const float* weight, const float* scale, const float* bias,
                     const tdn_tensor* out, cudaStream_t stream) {
  int rc;
  if ((rc = check_tensor(in, "pointwise_linear.in"))) return rc;
  if ((rc = check_tensor(out, "pointwise_linear.out"))) return rc;
  TDN_REQUIRE(weight != nullptr && aligned16(weight), TDN_ERR_INVALID, "pointwise_linear: weight must be 16-byte aligned");
  TDN_REQUIRE(out->c >= 1 && out->c <= 32, TDN_ERR_UNSUPPORTED, "pointwise_linear: 1..32 output channels, got %d", out->c);
  TDN_REQUIRE(in->n == out->n && in->h == out->h && in->w == out->w, TDN_ERR_INVALID, "pointwise_linear: map size mismatch");
  TDN_REQUIRE(vec4_ok(*in), TDN_ERR_INVALID, "pointwise_linear: float4-aligned input view required");
  const long long total_px = (long long)in->n * in->h * in->w;
  pointwise_linear_kernel<<<ceil_div(total_px, PL_PX), 256, 0, stream>>>(make_view(*in), weight, scale, bias,
                                                                         make_view(*out), out->c, total_px);
  TDN_LAUNCH_OK();
  return TDN_OK;
}

// Bilinear blend with explicitly rounded products and sums (no FMA contraction), shared by the logits and the
// arg-max upsamplers so that both produce bit-identical interpolated values.
__device__ __forceinline__ float bilerp(float t0, float t1, float u0, float u1, float lx, float ly) {
  const float wx0 = 1.f - lx, wy0 = 1.f - ly;
  const float top = __fadd_rn(__fmul_rn(t0, wx0), __fmul_rn(t1, lx));
  const float bot = __fadd_rn(__fmul_rn(u0, wx0), __fmul_rn(u1, lx));
  return __fadd_rn(__fmul_rn(top, wy0), __fmul_rn(bot, ly));
}

// ---------------------------------------------------------------------------------------------
// Final x8 bilinear upsample (align_corners) NHWC low-res logits -> NCHW fp32.
// One thread produces 4 consecutive x of one (n, y) for all classes: float4 streaming stores that are
// contiguous across the warp (the 159 MB/frame output write is the whole cost; the 2.5 MB input
// stays in L1/L2).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) upsample_logits_kernel(View in, float* __restrict__ out, int H,
                                                              int W, float sy, float sx) {
  const int W4 = (W + 3) >> 2;
  long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  long long total = (long long)in.n * H * W4;
  if (idx >= total) return;
  int xq = idx % W4;
  long long t = idx / W4;
  int y = t % H;
  int b = t / H;
  int y0, y1; float ly;
  src_index(y, sy, in.h, y0, y1, ly);
  int x0[4], x1[4]; float lx[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) src_index(min(xq * 4 + j, W - 1), sx, in.w, x0[j], x1[j], lx[j]);
  const float* r0 = in.p + b * in.sn + y0 * in.sh;
  const float* r1 = in.p + b * in.sn + y1 * in.sh;
  const long long plane = (long long)H * W;
  float* o = out + (long long)b * in.c * plane + (long long)y * W + xq * 4;
  const bool full = (xq * 4 + 3 < W) && ((W & 3) == 0);
  // 4 consecutive outputs of a >= 3x upsample touch at most 3 consecutive low-res columns: load
  // those once per channel instead of 16 scalar loads
  const int xa = x0[0];
  const bool compact = sx <= (1.f / 3.f);   // then x0[3] - x0[0] <= 1 and x1[3] - x0[0] <= 2
  const int xb = min(xa + 1, in.w - 1), xc = min(xa + 2, in.w - 1);
  for (int c = 0; c < in.c; ++c) {
    float v[4];
    if (compact) {
      const float t0 = r0[xa * in.sw + c], t1 = r0[xb * in.sw + c], t2 = r0[xc * in.sw + c];
      const float u0 = r1[xa * in.sw + c], u1 = r1[xb * in.sw + c], u2 = r1[xc * in.sw + c];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int d0 = x0[j] - xa, d1 = x1[j] - xa;            // 0..1 and 0..2
        const float ta = d0 == 0 ? t0 : t1, tb = d1 == 0 ? t0 : (d1 == 1 ? t1 : t2);
        const float ua = d0 == 0 ? u0 : u1, ub = d1 == 0 ? u0 : (d1 == 1 ? u1 : u2);
        v[j] = bilerp(ta, tb, ua, ub, lx[j], ly);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        v[j] = bilerp(r0[x0[j] * in.sw + c], r0[x1[j] * in.sw + c], r1[x0[j] * in.sw + c], r1[x1[j] * in.sw + c],
                      lx[j], ly);
      }
    }
    float* oc = o + c * plane;
    if (full) {
      __stcs(reinterpret_cast<float4*>(oc), make_float4(v[0], v[1], v[2], v[3]));
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (xq * 4 + j < W) oc[j] = v[j];
    }
  }
}

// Same outputs (same bilerp, same operands) for the common >= 3x upsample: a block produces ONE output row and first
// stages the two low-resolution rows it needs in shared memory as [class][x][row] (the column behind the last one
// repeats it, so x + 1 is always readable; its weight is 0 there).  A thread then needs two 8-byte shared-memory loads
// per class and output pixel -- (top, bottom) at x0 and at x0 + 1 -- and no index arithmetic inside the class loop.
// History: per-thread global loads (19-float pixel pitch, ~11 cache lines per warp load: 20 M L1 wavefronts per
// 1024x2048 frame) 53 us; rows staged as [class][x] with three loads per row and register selects 47 us, issue-bound
// (ncu: 88 % issue slots busy, the selects); HBM needs 25 us for the 159 MB write.
__device__ __forceinline__ void upsample_stage_rows(const View& in, float2* srow, int pitch, int b, int y0, int y1) {
  const int C = in.c;
  const float* g0 = in.p + b * in.sn + y0 * in.sh;
  const float* g1 = in.p + b * in.sn + y1 * in.sh;
  const int row_elems = in.w * C;
  constexpr int UN = 4;                                // elements whose (two) loads are issued before the first store
  for (int i0 = threadIdx.x; i0 < row_elems; i0 += blockDim.x * UN) {
    float2 v[UN];
    int dst[UN];
#pragma unroll
    for (int k = 0; k < UN; ++k) {
      const int i = i0 + k * blockDim.x;
      dst[k] = -1;
      if (i < row_elems) {
        const int x = i / C, c = i - x * C;
        v[k] = make_float2(__ldg(g0 + x * in.sw + c), __ldg(g1 + x * in.sw + c));
        dst[k] = (c * pitch + x) | (x == in.w - 1 ? 0x40000000 : 0);
      }
    }
#pragma unroll
    for (int k = 0; k < UN; ++k) {
      if (dst[k] >= 0) {
        const int d = dst[k] & 0x3fffffff;
        srow[d] = v[k];
        if (dst[k] & 0x40000000) srow[d + 1] = v[k];
      }
    }
  }
}

__global__ void __launch_bounds__(256) upsample_logits_row_kernel(View in, float* __restrict__ out, int H, int W, float sy,
                                                                  float sx) {
  extern __shared__ float2 srow2[];                  // [C][in.w + 1] (top, bottom)
  const int C = in.c, pitch = in.w + 1;
  const int y = blockIdx.x, b = blockIdx.y;
  int y0, y1; float ly;
  src_index(y, sy, in.h, y0, y1, ly);
  upsample_stage_rows(in, srow2, pitch, b, y0, y1);
  __syncthreads();
  const int W4 = W >> 2;
  const long long plane = (long long)H * W;
  for (int xq = threadIdx.x; xq < W4; xq += 256) {
    const float2* p[4]; float lx[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int x0, x1;
      src_index(xq * 4 + j, sx, in.w, x0, x1, lx[j]);
      p[j] = srow2 + x0;
    }
    float* o = out + (long long)b * C * plane + (long long)y * W + xq * 4;
    for (int c = 0; c < C; ++c) {
      float v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 a = p[j][0], d = p[j][1];       // (top, bottom) at x0 and at x0 + 1 (= x1, or weight 0)
        v[j] = bilerp(a.x, d.x, a.y, d.y, lx[j], ly);
        p[j] += pitch;
      }
      __stcs(reinterpret_cast<float4*>(o + c * plane), make_float4(v[0], v[1], v[2], v[3]));
    }
  }
}

// Same interpolation, but only the arg-max class leaves the chip: uint8 label map [n, H, W]
// (Testing/test.py:61 takes output.max(1)[1] right after the forward; lowest index wins ties, like torch).
// Saves the 159 MB fp32 logits write and turns a 16.8 MB int64 D2H into 2 MB.
__global__ void __launch_bounds__(256) upsample_argmax_kernel(View in, uint8_t* __restrict__ labels, int H, int W,
                                                              float sy, float sx) {
  const int W4 = (W + 3) >> 2;
  long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  long long total = (long long)in.n * H * W4;
  if (idx >= total) return;
  int xq = idx % W4;
  long long t = idx / W4;
  int y = t % H;
  int b = t / H;
  int y0, y1; float ly;
  src_index(y, sy, in.h, y0, y1, ly);
  int x0[4], x1[4]; float lx[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) src_index(min(xq * 4 + j, W - 1), sx, in.w, x0[j], x1[j], lx[j]);
  const float* r0 = in.p + b * in.sn + y0 * in.sh;
  const float* r1 = in.p + b * in.sn + y1 * in.sh;
  float best[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
  int arg[4] = {0, 0, 0, 0};
  for (int c = 0; c < in.c; ++c) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      // identical arithmetic to upsample_logits_kernel, so labels == argmax of the logits it would write
      const float v = bilerp(r0[x0[j] * in.sw + c], r0[x1[j] * in.sw + c], r1[x0[j] * in.sw + c],
                             r1[x1[j] * in.sw + c], lx[j], ly);
      if (v > best[j]) { best[j] = v; arg[j] = c; }
    }
  }
  uint8_t* o = labels + ((long long)b * H + y) * W + xq * 4;
  if ((xq * 4 + 3 < W) && ((W & 3) == 0)) {
    *reinterpret_cast<uchar4*>(o) = make_uchar4((uint8_t)arg[0], (uint8_t)arg[1], (uint8_t)arg[2], (uint8_t)arg[3]);
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (xq * 4 + j < W) o[j] = (uint8_t)arg[j];
  }
}

// Shared-memory version of the arg-max upsampler, built like upsample_logits_row_kernel (one output row per block, the two
// low-resolution rows staged as [class][x][row]) with the identical bilerp on identical operands: labels == arg-max of the
// logits either logits kernel writes.  No 159 MB logits write.
__global__ void __launch_bounds__(256) upsample_argmax_row_kernel(View in, uint8_t* __restrict__ labels, int H, int W,
                                                                  float sy, float sx) {
  extern __shared__ float2 srow2[];                  // [C][in.w + 1] (top, bottom)
  const int C = in.c, pitch = in.w + 1;
  const int y = blockIdx.x, b = blockIdx.y;
  int y0, y1; float ly;
  src_index(y, sy, in.h, y0, y1, ly);
  upsample_stage_rows(in, srow2, pitch, b, y0, y1);
  __syncthreads();
  const int W4 = W >> 2;
  for (int xq = threadIdx.x; xq < W4; xq += 256) {
    const float2* p[4]; float lx[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int x0, x1;
      src_index(xq * 4 + j, sx, in.w, x0, x1, lx[j]);
      p[j] = srow2 + x0;
    }
    float best[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
    int arg[4] = {0, 0, 0, 0};
    for (int c = 0; c < C; ++c) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 a = p[j][0], d = p[j][1];
        const float v = bilerp(a.x, d.x, a.y, d.y, lx[j], ly);
        if (v > best[j]) { best[j] = v; arg[j] = c; }
        p[j] += pitch;
      }
    }
    *reinterpret_cast<uchar4*>(labels + ((long long)b * H + y) * W + xq * 4) =
        make_uchar4((uint8_t)arg[0], (uint8_t)arg[1], (uint8_t)arg[2], (uint8_t)arg[3]);
  }
}

int upsample_argmax(const tdn_tensor* in, uint8_t* labels, int out_h, int out_w, cudaStream_t stream) {
  int rc;
  if ((rc = check_f32_tensor(in, "upsample_argmax.in"))) return rc;
  TDN_REQUIRE(labels && out_h > 0 && out_w > 0 && in->c <= 256, TDN_ERR_INVALID, "upsample_argmax: bad output / > 256 classes");
  TDN_REQUIRE((((uintptr_t)labels) & 3u) == 0, TDN_ERR_INVALID, "upsample_argmax: labels must be 4-byte aligned");
  const float sx = ac_scale(in->w, out_w);
  const size_t row_smem = (size_t)2 * in->c * (in->w + 1) * sizeof(float);
  if (sx <= 1.f / 3.f && (out_w & 3) == 0 && row_smem <= 96 * 1024 && out_h >= 128) {
    static PerDeviceFlag attr_set;
    const int slot = current_device_slot();
    if (!attr_set.is_set(slot)) {
      TDN_CUDA_OK(cudaFuncSetAttribute(upsample_argmax_row_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
      attr_set.set(slot);
    }
    upsample_argmax_row_kernel<<<dim3(out_h, in->n), 256, row_smem, stream>>>(make_view(*in), labels, out_h, out_w,
                                                                             ac_scale(in->h, out_h), sx);
    TDN_LAUNCH_OK();
    return TDN_OK;
  }
  long long total = (long long)in->n * out_h * ((out_w + 3) / 4);
  upsample_argmax_kernel<<<ceil_div(total, 256), 256, 0, stream>>>(make_view(*in), labels, out_h, out_w,
                                                                   ac_scale(in->h, out_h), sx);
  TDN_LAUNCH_OK();
  return TDN_OK;
}

// The labels test.py actually keeps (Testing/test.py:61-64): arg-max, then cv2.resize(..., INTER_NEAREST) to a quarter
// of the frame.  Nearest resampling picks full-resolution pixels (ys[oy], xs[ox]); only those are interpolated and
// arg-maxed here (1/16 of the work of the full label map), with the same bilerp as the two kernels above.
__global__ void __launch_bounds__(256) upsample_argmax_sampled_kernel(View in, uint8_t* __restrict__ labels, float sy,
                                                                      float sx, const int* __restrict__ ys,
                                                                      const int* __restrict__ xs, int Ho, int Wo) {
  long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  long long total = (long long)in.n * Ho * Wo;
  if (idx >= total) return;
  const int ox = (int)(idx % Wo);
  long long t = idx / Wo;
  const int oy = (int)(t % Ho);
  const int b = (int)(t / Ho);
  int y0, y1, x0, x1; float ly, lx;
  src_index(__ldg(ys + oy), sy, in.h, y0, y1, ly);
  src_index(__ldg(xs + ox), sx, in.w, x0, x1, lx);
  const float* r0 = in.p + b * in.sn + y0 * in.sh;
  const float* r1 = in.p + b * in.sn + y1 * in.sh;
  float best = -INFINITY;
  int arg = 0;
  for (int c = 0; c < in.c; ++c) {
    const float v = bilerp(r0[x0 * in.sw + c], r0[x1 * in.sw + c], r1[x0 * in.sw + c], r1[x1 * in.sw + c], lx, ly);
    if (v > best) { best = v; arg = c; }
  }
  labels[idx] = (uint8_t)arg;
}

int upsample_argmax_sampled(const tdn_tensor* in, uint8_t* labels, int full_h, int full_w, const int* ys, const int* xs,
                            int out_h, int out_w, cudaStream_t stream) {
  int rc;
  if ((rc = check_f32_tensor(in, "upsample_argmax_sampled.in"))) return rc;
  TDN_REQUIRE(labels && ys && xs && out_h > 0 && out_w > 0 && full_h > 0 && full_w > 0 && in->c <= 256, TDN_ERR_INVALID,
              "upsample_argmax_sampled: bad output / tables / > 256 classes");
  long long total = (long long)in->n * out_h * out_w;
  upsample_argmax_sampled_kernel<<<ceil_div(total, 256), 256, 0, stream>>>(
      make_view(*in), labels, ac_scale(in->h, full_h), ac_scale(in->w, full_w), ys, xs, out_h, out_w);
  TDN_LAUNCH_OK();
  return TDN_OK;
}

// ---------------------------------------------------------------------------------------------
// cv2.resize(frame, (W, H)) of the uint8 RGB camera frame (Testing/dataloader.py:63; INTER_LINEAR, OpenCV's 8-bit
// fixed-point path): integer arithmetic only, bit-exact.  The per-column / per-row taps (source offsets and 11-bit
// weights, which OpenCV derives with float / double arithmetic) come from the host as int4 tables
// {offset0, offset1, weight0, weight1}; a thread produces the three channels of one output pixel:
//   t_r = S[row_r][x0] * a0 + S[row_r][x1] * a1            (r = 0, 1; values scaled by 2048)
//   out = (((b0 * (t_0 >> 4)) >> 16) + ((b1 * (t_1 >> 4)) >> 16) + 2) >> 2
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) resize_linear_u8_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst,
                                                               const int4* __restrict__ xt, const int4* __restrict__ yt,
                                                               int n, int h, int w, int H, int W) {
  long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  long long total = (long long)n * H * W;
  if (idx >= total) return;
  const int x = (int)(idx % W);
  long long t = idx / W;
  const int y = (int)(t % H);
  const int b = (int)(t / H);
  const int4 xc = __ldg(xt + x), yc = __ldg(yt + y);
  const uint8_t* r0 = src + ((long long)b * h + yc.x) * w * 3;
  const uint8_t* r1 = src + ((long long)b * h + yc.y) * w * 3;
  uint8_t* o = dst + idx * 3;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const int t0 = (int)__ldg(r0 + xc.x * 3 + c) * xc.z + (int)__ldg(r0 + xc.y * 3 + c) * xc.w;
    const int t1 = (int)__ldg(r1 + xc.x * 3 + c) * xc.z + (int)__ldg(r1 + xc.y * 3 + c) * xc.w;
    int v = (((yc.z * (t0 >> 4)) >> 16) + ((yc.w * (t1 >> 4)) >> 16) + 2) >> 2;
    o[c] = (uint8_t)min(max(v, 0), 255);
  }
}

int resize_linear_u8(const uint8_t* src, int n, int h, int w, const int* xt, const int* yt, uint8_t* dst, int H, int W,
                     cudaStream_t stream) {
  TDN_REQUIRE(src && dst && xt && yt, TDN_ERR_INVALID, "resize_linear_u8: null pointer");
  TDN_REQUIRE(n > 0 && h > 0 && w > 0 && H > 0 && W > 0, TDN_ERR_INVALID, "resize_linear_u8: empty image");
  TDN_REQUIRE(aligned16(xt) && aligned16(yt), TDN_ERR_INVALID, "resize_linear_u8: tap tables must be 16-byte aligned");
  long long total = (long long)n * H * W;
  resize_linear_u8_kernel<<<ceil_div(total, 256), 256, 0, stream>>>(src, dst, reinterpret_cast<const int4*>(xt),
                                                                    reinterpret_cast<const int4*>(yt), n, h, w, H, W);
  TDN_LAUNCH_OK();
  return TDN_OK;
}

int upsample_logits(const tdn_tensor* in, float* out_nchw, int out_h, int out_w, cudaStream_t stream) {
  int rc;
  if ((rc = check_f32_tensor(in, "upsample.in"))) return rc;
  TDN_REQUIRE(out_nchw && out_h > 0 && out_w > 0, TDN_ERR_INVALID, "upsample_logits: bad output");
  TDN_REQUIRE(aligned16(out_nchw), TDN_ERR_INVALID, "upsample_logits: output must be 16-byte aligned");
  const float sx = ac_scale(in->w, out_w);
  const size_t row_smem = (size_t)2 * in->c * (in->w + 1) * sizeof(float);
  if (sx <= 1.f / 3.f && (out_w & 3) == 0 && row_smem <= 96 * 1024 && out_h >= 128) {
    static PerDeviceFlag attr_set;
    const int slot = current_device_slot();
    if (!attr_set.is_set(slot)) {
      TDN_CUDA_OK(cudaFuncSetAttribute(upsample_logits_row_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
      attr_set.set(slot);
    }
    upsample_logits_row_kernel<<<dim3(out_h, in->n), 256, row_smem, stream>>>(make_view(*in), out_nchw, out_h, out_w,
                                                                             ac_scale(in->h, out_h), sx);
    TDN_LAUNCH_OK();
    return TDN_OK;
  }
  long long total = (long long)in->n * out_h * ((out_w + 3) / 4);
  upsample_logits_kernel<<<ceil_div(total, 256), 256, 0, stream>>>(
      make_view(*in), out_nchw, out_h, out_w, ac_scale(in->h, out_h), ac_scale(in->w, out_w));
  TDN_LAUNCH_OK();
  return TDN_OK;
}

}  // namespace tdn
