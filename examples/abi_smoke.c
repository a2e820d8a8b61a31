/* Plain-C consumer of the tdnet_b200 C ABI (include/tdnet_b200.h, tdnet_b200/lib/libtdnet_b200.so).
 *
 *   gcc -std=c99 -Iinclude examples/abi_smoke.c -Ltdnet_b200/lib -ltdnet_b200 -Wl,-rpath,$PWD/tdnet_b200/lib -o abi_smoke
 *
 * Without a GPU it exercises what does not need one: version, error strings, workspace sizing, and the argument
 * validation that every entry point performs before its first CUDA call.  On a B200 it additionally runs one
 * 1x1 convolution through tdn_conv2d on device buffers it allocates through the CUDA runtime (link with -lcudart and
 * define WITH_CUDA).  tests/test_cabi.py builds and runs the GPU-less variant. */
#include <stdio.h>
#include <string.h>

#include "tdnet_b200.h"

#ifdef WITH_CUDA
#include <cuda_runtime_api.h>
#endif

#define CHECK(cond)                                              \
  do {                                                           \
    if (!(cond)) {                                               \
      fprintf(stderr, "abi_smoke: check failed: %s\n", #cond);   \
      return 1;                                                  \
    }                                                            \
  } while (0)

int main(void) {
  tdn_conv2d_desc d;
  CHECK(tdn_abi_version() == TDN_ABI_VERSION);
  CHECK(strcmp(tdn_strerror(TDN_OK), "ok") == 0);
  CHECK(strstr(tdn_strerror(TDN_ERR_ARCH), "sm_100") != NULL);
  CHECK(tdn_psp_pool_workspace_bytes(1, 128, 512) == (uint64_t)128 * 12 * 512 * 4);
  CHECK(tdn_layernorm_hw_workspace_bytes(1, 128, 256, 512) == (uint64_t)512 * 512 * 16);
  CHECK(tdn_fa_context_workspace_bytes(1, 128, 256, 64) == (uint64_t)128 * 32 * 64 * 4);

  /* argument validation happens before any CUDA call */
  CHECK(tdn_conv2d(NULL, NULL) == TDN_ERR_INVALID);
  CHECK(strstr(tdn_last_error(), "null descriptor") != NULL);
  memset(&d, 0, sizeof(d));
  CHECK(tdn_conv2d(&d, NULL) == TDN_ERR_INVALID);           /* null data pointers */
  CHECK(tdn_resize_linear_u8(NULL, 1, 4, 4, NULL, NULL, NULL, 8, 8, NULL) == TDN_ERR_INVALID);

#ifdef WITH_CUDA
  {
    /* out[p][co] = sum_ci in[p][ci] * w[co][ci]: 8 pixels, 4 -> 4 channels, identity weights */
    float h_in[8 * 4], h_w[4 * 4], h_out[8 * 4];
    float *in, *w, *out;
    int i;
    for (i = 0; i < 32; ++i) h_in[i] = (float)i;
    memset(h_w, 0, sizeof(h_w));
    for (i = 0; i < 4; ++i) h_w[i * 4 + i] = 1.f;
    CHECK(cudaMalloc((void**)&in, sizeof(h_in)) == cudaSuccess);
    CHECK(cudaMalloc((void**)&w, sizeof(h_w)) == cudaSuccess);
    CHECK(cudaMalloc((void**)&out, sizeof(h_out)) == cudaSuccess);
    cudaMemcpy(in, h_in, sizeof(h_in), cudaMemcpyHostToDevice);
    cudaMemcpy(w, h_w, sizeof(h_w), cudaMemcpyHostToDevice);
    memset(&d, 0, sizeof(d));
    d.in.data = in;   d.in.dtype = TDN_F32;  d.in.n = 1;  d.in.h = 2;  d.in.w = 4;  d.in.c = 4;
    d.in.stride_n = 32; d.in.stride_h = 16; d.in.stride_w = 4;
    d.out = d.in;     d.out.data = out;
    d.weight = w;     d.cout = 4; d.kh = d.kw = 1; d.stride = 1; d.dilation = 1; d.batch = 1;
    CHECK(tdn_conv2d(&d, NULL) == TDN_OK);
    CHECK(cudaMemcpy(h_out, out, sizeof(h_out), cudaMemcpyDeviceToHost) == cudaSuccess);
    for (i = 0; i < 32; ++i) CHECK(h_out[i] == h_in[i]);
    cudaFree(in); cudaFree(w); cudaFree(out);
  }
#endif
  printf("abi_smoke: ok (ABI version %d)\n", tdn_abi_version());
  return 0;
}
