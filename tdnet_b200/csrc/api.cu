// extern "C" surface of libtdnet_b200.so (see include/tdnet_b200.h).  Argument validation and
// kernel selection live here; no function throws or synchronises.
#include "common.cuh"

#include <string.h>

namespace tdn {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* get_error() { return g_err; }

int check_tensor(const tdn_tensor* t, const char* what) {
  TDN_REQUIRE(t != nullptr && t->data != nullptr, TDN_ERR_INVALID, "%s: null tensor", what);
  TDN_REQUIRE(t->dtype == TDN_F32 || (t->dtype == TDN_SPLIT16 && t->data_lo != nullptr), TDN_ERR_INVALID,
              "%s: dtype must be F32 or SPLIT16 (with a lo plane)", what);
  TDN_REQUIRE(t->n > 0 && t->h > 0 && t->w > 0 && t->c > 0, TDN_ERR_INVALID, "%s: empty dims [%d,%d,%d,%d]",
              what, t->n, t->h, t->w, t->c);
  return TDN_OK;
}

int check_f32_tensor(const tdn_tensor* t, const char* what) {
  int rc = check_tensor(t, what);
  if (rc) return rc;
  TDN_REQUIRE(t->dtype == TDN_F32, TDN_ERR_UNSUPPORTED, "%s: expected an fp32 plane", what);
  return TDN_OK;
}

int conv2d_simt(const tdn_conv2d_desc* d, cudaStream_t stream);
int image_to_nhwc(const float*, int, int, int, int, const tdn_tensor*, cudaStream_t);
int maxpool3x3s2(const tdn_tensor*, const tdn_tensor*, cudaStream_t);
int psp_pool(const tdn_tensor*, const tdn_tensor*, float*, size_t, cudaStream_t);
int bilinear_nhwc(const tdn_tensor*, const tdn_tensor*, cudaStream_t);
int copy_nhwc(const tdn_tensor*, const tdn_tensor*, cudaStream_t);
int psp_concat(const tdn_tensor*, const float* const*, int, const tdn_tensor*, cudaStream_t);
int psp_branch_convs(const tdn_tensor*, const float* const*, const float* const*, const float* const*, int, float* const*,
                     cudaStream_t);
int psp_branch_project(const tdn_tensor*, const float* const*, const float* const*, const float* const*, int, float* const*,
                       const tdn_psp_projection*, int, int*, cudaStream_t);
int softmax_rows(float*, long long, int, long long, float, cudaStream_t);
int softmax_rows_split16(const float*, long long, int, long long, float, void*, void*, long long, float, cudaStream_t);
int layernorm_hw_stats(const tdn_tensor*, float*, float*, float, void*, size_t, cudaStream_t);
int layernorm_hw_apply(const tdn_tensor*, const float*, const float*, const float*, const float*,
                       const tdn_tensor*, cudaStream_t);
int upsample_logits(const tdn_tensor*, float*, int, int, cudaStream_t);
int upsample_argmax(const tdn_tensor*, uint8_t*, int, int, cudaStream_t);
int conv2d_tc(const tdn_tc_conv_desc*, cudaStream_t);
int stem_conv_pool(const float*, const uint8_t*, const float*, int, int, int, const float*, const float*, const float*,
                   const tdn_tensor*, cudaStream_t);
int attention_tc(const tdn_attention_desc*, cudaStream_t, int* count_only);
int stem_conv_pool_tc(const float*, const uint8_t*, const float*, int, int, int, const void*, const float*, const float*,
                      const tdn_tensor*, int, float, int*, cudaStream_t);
int split16(const tdn_tensor*, const tdn_tensor*, cudaStream_t);
int pointwise_linear(const tdn_tensor*, const float*, const float*, const float*, const tdn_tensor*, cudaStream_t);
int upsample_argmax_sampled(const tdn_tensor*, uint8_t*, int, int, const int*, const int*, int, int, cudaStream_t);
int resize_linear_u8(const uint8_t*, int, int, int, const int*, const int*, uint8_t*, int, int, cudaStream_t);
int fa_context(const tdn_tensor*, const tdn_tensor*, float*, void*, size_t, cudaStream_t);
size_t fa_context_workspace_bytes(int, int, int, int);
int fa_apply(const tdn_tensor*, const float*, const tdn_tensor*, float, int*, cudaStream_t);
int add_upsampled(const tdn_tensor*, const tdn_tensor*, const tdn_tensor*, const tdn_tensor*, cudaStream_t);
int merge16(const tdn_tensor*, const tdn_tensor*, cudaStream_t);

int device_sm_count() {
  static PerDeviceInt cache;
  const int slot = current_device_slot();
  int n = cache.get(slot);
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
    cache.set(slot, n);
  }
  return n;
}

}  // namespace tdn

using namespace tdn;

namespace tdn {
// one warp per CTA; lane 0 spins on %globaltimer and reports how many SM cycles passed meanwhile
__global__ void sm_clock_probe_kernel(unsigned long long* out, long long min_ns) {
  if (threadIdx.x != 0) return;
  unsigned long long t0, t1, c0, c1;
  unsigned int smid;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  asm volatile("mov.u64 %0, %%clock64;" : "=l"(c0));
  do {
    __nanosleep(200);
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
  } while ((long long)(t1 - t0) < min_ns);
  asm volatile("mov.u64 %0, %%clock64;" : "=l"(c1));
  asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
  out[3 * blockIdx.x + 0] = c1 - c0;
  out[3 * blockIdx.x + 1] = t1 - t0;
  out[3 * blockIdx.x + 2] = smid;
}
}  // namespace tdn

extern "C" {

int tdn_abi_version(void) { return TDN_ABI_VERSION; }

int tdn_sm_clock_probe(uint64_t* out, int32_t blocks, int64_t min_ns, void* stream) {
  TDN_REQUIRE(out && blocks > 0 && min_ns > 0, TDN_ERR_INVALID, "sm_clock_probe: bad arguments");
  tdn::sm_clock_probe_kernel<<<blocks, 32, 0, (cudaStream_t)stream>>>((unsigned long long*)out, (long long)min_ns);
  TDN_LAUNCH_OK();
  return TDN_OK;
}

const char* tdn_strerror(int status) {
  switch (status) {
    case TDN_OK: return "ok";
    case TDN_ERR_INVALID: return "invalid argument";
    case TDN_ERR_UNSUPPORTED: return "unsupported shape or option";
    case TDN_ERR_CUDA: return "CUDA error";
    case TDN_ERR_ARCH: return "device is not sm_100 (B200)";
    case TDN_ERR_WORKSPACE: return "workspace too small";
    default: return "unknown status";
  }
}

const char* tdn_last_error(void) { return get_error(); }

int tdn_device_arch(void) {
  int dev = 0, major = 0, minor = 0;
  TDN_CUDA_OK(cudaGetDevice(&dev));
  TDN_CUDA_OK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  TDN_CUDA_OK(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
  return major * 10 + minor;
}

// Compute capability of the current device, cached per device ordinal (negative: a tdn_status error).
static int cached_device_arch() {
  static tdn::PerDeviceInt cache;
  const int slot = tdn::current_device_slot();
  int arch = cache.get(slot);
  if (arch == 0) {
    arch = tdn_device_arch();
    if (arch > 0) cache.set(slot, arch);
  }
  return arch;
}

int tdn_conv2d(const tdn_conv2d_desc* d, void* stream) {
  TDN_REQUIRE(d != nullptr, TDN_ERR_INVALID, "conv2d: null descriptor");
  TDN_REQUIRE(d->in.data && d->out.data && d->weight, TDN_ERR_INVALID, "conv2d: null data pointer");
  TDN_REQUIRE(d->kh > 0 && d->kw > 0 && d->stride > 0 && d->dilation > 0 && d->pad >= 0 && d->cout > 0 &&
                  d->batch >= 1, TDN_ERR_INVALID, "conv2d: bad geometry");
  TDN_REQUIRE(d->in.n > 0 && d->in.h > 0 && d->in.w > 0 && d->in.c > 0, TDN_ERR_INVALID,
              "conv2d: empty input");
  return conv2d_simt(d, (cudaStream_t)stream);
}

int tdn_conv2d_tc(const tdn_tc_conv_desc* d, void* stream) {
  TDN_REQUIRE(d != nullptr, TDN_ERR_INVALID, "conv2d_tc: null descriptor");
  const int arch = cached_device_arch();
  if (arch < 0) return arch;
  TDN_REQUIRE(arch / 10 == 10, TDN_ERR_ARCH, "conv2d_tc: tcgen05 kernels need sm_100, device is sm_%d", arch);
  TDN_REQUIRE(d->cout > 0 && d->kh > 0 && d->kw > 0 && d->dilation > 0, TDN_ERR_INVALID, "conv2d_tc: bad geometry");
  return conv2d_tc(d, (cudaStream_t)stream);
}

int tdn_attention_tc(const tdn_attention_desc* d, void* stream) {
  TDN_REQUIRE(d != nullptr, TDN_ERR_INVALID, "attention_tc: null descriptor");
  const int arch = cached_device_arch();
  if (arch < 0) return arch;
  TDN_REQUIRE(arch / 10 == 10, TDN_ERR_ARCH, "attention_tc: tcgen05 kernels need sm_100, device is sm_%d", arch);
  return attention_tc(d, (cudaStream_t)stream, nullptr);
}

int tdn_attention_tc_launches(const tdn_attention_desc* d, int32_t* launches) {
  TDN_REQUIRE(d != nullptr && launches != nullptr, TDN_ERR_INVALID, "attention_tc_launches: null argument");
  int n = 0;
  const int rc = attention_tc(d, nullptr, &n);
  *launches = n;
  return rc;
}

int tdn_split16(const tdn_tensor* in, const tdn_tensor* out, void* stream) {
  return split16(in, out, (cudaStream_t)stream);
}

int tdn_merge16(const tdn_tensor* in, const tdn_tensor* out, void* stream) {
  return merge16(in, out, (cudaStream_t)stream);
}

int tdn_image_to_nhwc(const float* nchw, int32_t n, int32_t c, int32_t h, int32_t w, const tdn_tensor* out,
                      void* stream) {
  return image_to_nhwc(nchw, n, c, h, w, out, (cudaStream_t)stream);
}

int tdn_stem_conv_pool(const float* nchw, int32_t n, int32_t h, int32_t w, const float* weight, const float* scale,
                       const float* bias, const tdn_tensor* out, void* stream) {
  return stem_conv_pool(nchw, nullptr, nullptr, n, h, w, weight, scale, bias, out, (cudaStream_t)stream);
}

int tdn_stem_conv_pool_tc(const float* nchw, const uint8_t* hwc_u8, const float* lut, int32_t n, int32_t h, int32_t w,
                          const void* weight_tc, const float* scale, const float* bias, const tdn_tensor* out,
                          int32_t* range_flag, void* stream) {
  const int arch = cached_device_arch();
  if (arch < 0) return arch;
  TDN_REQUIRE(arch / 10 == 10, TDN_ERR_ARCH, "stem_conv_pool_tc: tcgen05 kernels need sm_100, device is sm_%d", arch);
  return stem_conv_pool_tc(nchw, hwc_u8, lut, n, h, w, weight_tc, scale, bias, out, TDN_ACT_RELU, 0.f, range_flag,
                           (cudaStream_t)stream);
}

int tdn_stem_conv_pool_tc_act(const float* nchw, const uint8_t* hwc_u8, const float* lut, int32_t n, int32_t h, int32_t w,
                              const void* weight_tc, const float* scale, const float* bias, const tdn_tensor* out,
                              int32_t act, float leaky_slope, int32_t* range_flag, void* stream) {
  const int arch = cached_device_arch();
  if (arch < 0) return arch;
  TDN_REQUIRE(arch / 10 == 10, TDN_ERR_ARCH, "stem_conv_pool_tc: tcgen05 kernels need sm_100, device is sm_%d", arch);
  return stem_conv_pool_tc(nchw, hwc_u8, lut, n, h, w, weight_tc, scale, bias, out, act, leaky_slope, range_flag,
                           (cudaStream_t)stream);
}

int tdn_stem_conv_pool_u8(const uint8_t* hwc, const float* lut, int32_t n, int32_t h, int32_t w, const float* weight,
                          const float* scale, const float* bias, const tdn_tensor* out, void* stream) {
  TDN_REQUIRE(hwc && lut, TDN_ERR_INVALID, "stem_u8: null image / lut");
  return stem_conv_pool(nullptr, hwc, lut, n, h, w, weight, scale, bias, out, (cudaStream_t)stream);
}

int tdn_maxpool3x3s2(const tdn_tensor* in, const tdn_tensor* out, void* stream) {
  return maxpool3x3s2(in, out, (cudaStream_t)stream);
}

uint64_t tdn_psp_pool_workspace_bytes(int32_t n, int32_t h, int32_t c) {
  return (uint64_t)n * h * 12 * c * sizeof(float);
}

int tdn_psp_pool(const tdn_tensor* in, const tdn_tensor* out, void* workspace, uint64_t workspace_bytes,
                 void* stream) {
  return psp_pool(in, out, (float*)workspace, (size_t)workspace_bytes, (cudaStream_t)stream);
}

int tdn_bilinear_nhwc(const tdn_tensor* in, const tdn_tensor* out, void* stream) {
  return bilinear_nhwc(in, out, (cudaStream_t)stream);
}

int tdn_psp_branch_convs(const tdn_tensor* pooled, const float* const* w, const float* const* scale,
                         const float* const* bias, int32_t eighth, float* const* out, void* stream) {
  return psp_branch_convs(pooled, w, scale, bias, eighth, out, (cudaStream_t)stream);
}

int tdn_psp_branch_project(const tdn_tensor* pooled, const float* const* w, const float* const* scale,
                           const float* const* bias, int32_t eighth, float* const* out,
                           const tdn_psp_projection* proj, int32_t n_proj, int32_t* range_flag, void* stream) {
  return psp_branch_project(pooled, w, scale, bias, eighth, out, proj, n_proj, range_flag, (cudaStream_t)stream);
}

int tdn_psp_concat(const tdn_tensor* x, const float* const* small, int32_t eighth, const tdn_tensor* z, void* stream) {
  return psp_concat(x, small, eighth, z, (cudaStream_t)stream);
}

int tdn_copy_nhwc(const tdn_tensor* in, const tdn_tensor* out, void* stream) {
  return copy_nhwc(in, out, (cudaStream_t)stream);
}

int tdn_softmax_rows(float* s, int64_t rows, int32_t cols, int64_t ld, float scale, void* stream) {
  return softmax_rows(s, rows, cols, ld, scale, (cudaStream_t)stream);
}

int tdn_softmax_rows_split16(const float* s, int64_t rows, int32_t cols, int64_t ld, float scale, void* p_hi,
                             void* p_lo, int64_t ld_out, float out_scale, void* stream) {
  return softmax_rows_split16(s, rows, cols, ld, scale, p_hi, p_lo, ld_out, out_scale, (cudaStream_t)stream);
}

uint64_t tdn_layernorm_hw_workspace_bytes(int32_t n, int32_t h, int32_t w, int32_t c) {
  uint64_t chunks = ((uint64_t)h * w + 63) / 64;
  return (uint64_t)n * chunks * c * 16;
}

int tdn_layernorm_hw_stats(const tdn_tensor* x, float* mean, float* rstd, float eps, void* workspace,
                           uint64_t workspace_bytes, void* stream) {
  return layernorm_hw_stats(x, mean, rstd, eps, workspace, (size_t)workspace_bytes, (cudaStream_t)stream);
}

int tdn_layernorm_hw_apply(const tdn_tensor* x, const float* mean, const float* rstd, const float* gamma,
                           const float* beta, const tdn_tensor* out, void* stream) {
  return layernorm_hw_apply(x, mean, rstd, gamma, beta, out, (cudaStream_t)stream);
}

int tdn_upsample_logits(const tdn_tensor* in, float* out_nchw, int32_t out_h, int32_t out_w, void* stream) {
  return upsample_logits(in, out_nchw, out_h, out_w, (cudaStream_t)stream);
}

int tdn_upsample_argmax(const tdn_tensor* in, uint8_t* labels, int32_t out_h, int32_t out_w, void* stream) {
  return upsample_argmax(in, labels, out_h, out_w, (cudaStream_t)stream);
}

int tdn_upsample_argmax_sampled(const tdn_tensor* in, uint8_t* labels, int32_t full_h, int32_t full_w, const int32_t* ys,
                                const int32_t* xs, int32_t out_h, int32_t out_w, void* stream) {
  return upsample_argmax_sampled(in, labels, full_h, full_w, ys, xs, out_h, out_w, (cudaStream_t)stream);
}

int tdn_resize_linear_u8(const uint8_t* src, int32_t n, int32_t h, int32_t w, const int32_t* x_taps, const int32_t* y_taps,
                         uint8_t* dst, int32_t out_h, int32_t out_w, void* stream) {
  return resize_linear_u8(src, n, h, w, x_taps, y_taps, dst, out_h, out_w, (cudaStream_t)stream);
}

int tdn_pointwise_linear(const tdn_tensor* in, const float* weight, const float* scale, const float* bias,
                         const tdn_tensor* out, void* stream) {
  return pointwise_linear(in, weight, scale, bias, out, (cudaStream_t)stream);
}

int tdn_fa_context(const tdn_tensor* key, const tdn_tensor* value, float* f, void* workspace,
                   uint64_t workspace_bytes, void* stream) {
  return fa_context(key, value, f, workspace, (size_t)workspace_bytes, (cudaStream_t)stream);
}

uint64_t tdn_fa_context_workspace_bytes(int32_t n, int32_t h, int32_t w, int32_t c) {
  return (uint64_t)fa_context_workspace_bytes(n, h, w, c);
}

int tdn_fa_apply(const tdn_tensor* query, const float* f, const tdn_tensor* out, float out_scale, int32_t* range_flag,
                 void* stream) {
  return fa_apply(query, f, out, out_scale, range_flag, (cudaStream_t)stream);
}

int tdn_add_upsampled(const tdn_tensor* a, const tdn_tensor* b, const tdn_tensor* up, const tdn_tensor* out,
                      void* stream) {
  return add_upsampled(a, b, up, out, (cudaStream_t)stream);
}

}  // extern "C"
