// Shared device/host helpers for the tdnet_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/tdnet_b200.h"

namespace tdn {

// Thread-local detail string behind tdn_last_error().
void set_error(const char* fmt, ...);
const char* get_error();

#define TDN_REQUIRE(cond, status, ...)            \
  do {                                            \
    if (!(cond)) {                                \
      ::tdn::set_error(__VA_ARGS__);              \
      return (status);                            \
    }                                             \
  } while (0)

#define TDN_CUDA_OK(expr)                                                            \
  do {                                                                               \
    cudaError_t e__ = (expr);                                                        \
    if (e__ != cudaSuccess) {                                                        \
      ::tdn::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__),      \
                       __FILE__, __LINE__);                                          \
      return TDN_ERR_CUDA;                                                           \
    }                                                                                \
  } while (0)

// Checks the launch that was just issued (no synchronisation).
#define TDN_LAUNCH_OK() TDN_CUDA_OK(cudaGetLastError())

static inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// Device-side NHWC view (fp32 plane).
struct View {
  float* p;
  int n, h, w, c;
  long long sn, sh, sw;
};

static inline View make_view(const tdn_tensor& t) {
  View v;
  v.p = (float*)t.data;
  v.n = t.n; v.h = t.h; v.w = t.w; v.c = t.c;
  v.sn = t.stride_n; v.sh = t.stride_h; v.sw = t.stride_w;
  return v;
}

static inline bool aligned16(const void* p) { return (((uintptr_t)p) & 15u) == 0; }

// A float4-friendly view: 16-byte aligned base, channel count and all strides multiples of 4.
static inline bool vec4_ok(const tdn_tensor& t) {
  return aligned16(t.data) && (t.c % 4 == 0) && (t.stride_n % 4 == 0) && (t.stride_h % 4 == 0) &&
         (t.stride_w % 4 == 0);
}

int check_f32_tensor(const tdn_tensor* t, const char* what);

}  // namespace tdn
