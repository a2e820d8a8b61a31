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

// One process may drive several GPUs (model.to('cuda:1')): the > 48 KB dynamic shared-memory opt-in of a kernel
// (cudaFuncSetAttribute), the SM count and the compute capability are properties of the CURRENT device, so every
// "done once" cache in the library is kept per device ordinal.
constexpr int TDN_MAX_DEVICES = 64;
static inline int current_device_slot() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= TDN_MAX_DEVICES) return -1;   // -1: never cached
  return dev;
}
struct PerDeviceFlag {
  bool done[TDN_MAX_DEVICES] = {};
  bool is_set(int slot) const { return slot >= 0 && done[slot]; }
  void set(int slot) { if (slot >= 0) done[slot] = true; }
};
struct PerDeviceInt {
  int v[TDN_MAX_DEVICES] = {};
  int get(int slot) const { return slot >= 0 ? v[slot] : 0; }
  void set(int slot, int x) { if (slot >= 0) v[slot] = x; }
};
// SM count of the current device (cached per device); 0 on failure.
int device_sm_count();

// Device-side NHWC view: one fp32 plane (p) or two fp16 planes hi/lo with value = hi + lo
// (TDN_SPLIT16).  Every layout kernel goes through ld4/st4/ld1/st1, so it accepts either format.
struct View {
  float* p;
  __half* hi;
  __half* lo;
  int split;
  int n, h, w, c;
  long long sn, sh, sw;
};

static inline View make_view(const tdn_tensor& t) {
  View v;
  v.split = t.dtype == TDN_SPLIT16;
  v.p = v.split ? nullptr : (float*)t.data;
  v.hi = v.split ? (__half*)t.data : nullptr;
  v.lo = v.split ? (__half*)t.data_lo : nullptr;
  v.n = t.n; v.h = t.h; v.w = t.w; v.c = t.c;
  v.sn = t.stride_n; v.sh = t.stride_h; v.sw = t.stride_w;
  return v;
}

__device__ __forceinline__ void split_f32(float x, __half& hi, __half& lo) {
  hi = __float2half_rn(x);
  lo = __float2half_rn(x - __half2float(hi));
}
// Two values at once through the packed converter (cvt.rn.f16x2.f32 -> one F2FP per pair instead of two
// F2F on the quarter-rate conversion pipe); identical results to split_f32.
__device__ __forceinline__ void split_f32x2(float a, float b, __half2& hi, __half2& lo) {
  hi = __floats2half2_rn(a, b);
  const float2 hf = __half22float2(hi);
  lo = __floats2half2_rn(a - hf.x, b - hf.y);
}

// Four consecutive channels at element offset `off` (off % 4 == 0, planes suitably aligned).
__device__ __forceinline__ float4 ld4(const View& v, long long off) {
  if (!v.split) return *reinterpret_cast<const float4*>(v.p + off);
  uint2 hv = *reinterpret_cast<const uint2*>(v.hi + off);
  uint2 lv = *reinterpret_cast<const uint2*>(v.lo + off);
  const __half2* h = reinterpret_cast<const __half2*>(&hv);
  const __half2* l = reinterpret_cast<const __half2*>(&lv);
  float2 h0 = __half22float2(h[0]), h1 = __half22float2(h[1]);
  float2 l0 = __half22float2(l[0]), l1 = __half22float2(l[1]);
  return make_float4(h0.x + l0.x, h0.y + l0.y, h1.x + l1.x, h1.y + l1.y);
}
__device__ __forceinline__ void st4(const View& v, long long off, float4 x) {
  if (!v.split) {
    *reinterpret_cast<float4*>(v.p + off) = x;
    return;
  }
  __half2 h[2], l[2];
  split_f32x2(x.x, x.y, h[0], l[0]);
  split_f32x2(x.z, x.w, h[1], l[1]);
  *reinterpret_cast<uint2*>(v.hi + off) = *reinterpret_cast<const uint2*>(h);
  *reinterpret_cast<uint2*>(v.lo + off) = *reinterpret_cast<const uint2*>(l);
}
__device__ __forceinline__ float ld1(const View& v, long long off) {
  if (!v.split) return v.p[off];
  return __half2float(v.hi[off]) + __half2float(v.lo[off]);
}
__device__ __forceinline__ void st1(const View& v, long long off, float x) {
  if (!v.split) { v.p[off] = x; return; }
  __half h, l;
  split_f32(x, h, l);
  v.hi[off] = h; v.lo[off] = l;
}

static inline bool aligned16(const void* p) { return (((uintptr_t)p) & 15u) == 0; }
static inline bool aligned8(const void* p) { return (((uintptr_t)p) & 7u) == 0; }

// A 4-channel-vector friendly view: aligned base(s), channel count and all strides multiples of 4.
static inline bool vec4_ok(const tdn_tensor& t) {
  bool base = t.dtype == TDN_SPLIT16 ? (aligned8(t.data) && aligned8(t.data_lo)) : aligned16(t.data);
  return base && (t.c % 4 == 0) && (t.stride_n % 4 == 0) && (t.stride_h % 4 == 0) && (t.stride_w % 4 == 0);
}

// Validates a view of either dtype (non-null planes, positive dims).
int check_tensor(const tdn_tensor* t, const char* what);
// Same, and requires an fp32 plane.
int check_f32_tensor(const tdn_tensor* t, const char* what);

}  // namespace tdn
