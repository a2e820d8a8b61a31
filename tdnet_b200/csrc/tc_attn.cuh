// Shared pieces of the fused attention-propagation kernels (tc_attn.cu: P through shared memory;
// tc_attn_ts.cu: P through tensor memory): tile constants, the parameter block and the exp2 helper.
#pragma once
#include "common.cuh"
#include "tc_common.cuh"   // tc_launch / tc_pdl_sync (programmatic dependent launch); pulls in tc_ptx.cuh

#include <cuda.h>

namespace tdn {

using namespace ptx;

constexpr int AT_BQ = 128;       // queries per item
constexpr int AT_BK = 64;        // keys per tile (= one 128-byte swizzle row of fp16)
constexpr int AT_BK1 = 128;      // keys per PASS-1 tile (hi planes only: two 64-key boxes fill one K stage)
constexpr int AT_DK = 64;        // d_k (fixed by the model: Encoding(d_model, 64, d_v))
constexpr int AT_DVH = 128;      // V'^T rows per shared-memory stage / per PV MMA (N = 128)
constexpr int AT_THREADS = 352;  // warp 0 TMA, warp 1 S-MMA issuer, warps 2-9 softmax + epilogue, warp 10 PV-MMA issuer
constexpr int AT_PV_WARP = 10;
constexpr int AT_SOFTMAX_THREADS = 256;
constexpr int AT_SOFTMAX_WARPS = AT_SOFTMAX_THREADS / 32;
constexpr int AT_Q_PLANE = AT_BQ * AT_DK * 2;   // 16 KB
constexpr int AT_K_PLANE = AT_BK * AT_DK * 2;   // 8 KB
constexpr int AT_V_PLANE = AT_DVH * AT_BK * 2;  // 16 KB
constexpr int AT_P_PLANE = AT_BQ * AT_BK * 2;   // 16 KB
constexpr int AT_KSTAGES = 2, AT_VSTAGES = 3;
constexpr int AT_SMEM_DATA = 2 * AT_Q_PLANE + AT_KSTAGES * 2 * AT_K_PLANE + AT_VSTAGES * 2 * AT_V_PLANE + 2 * 2 * AT_P_PLANE;
constexpr int AT_TMEM_COLS = 512;   // S: 2 x 64 columns, O: up to 256 columns
constexpr float AT_P_SCALE = 1024.f;
constexpr int ATTN_DEFAULT_VARIANT = 1;   // TDNET_ATTN_TS default: 0 tc_attn.cu, 1 tc_attn_ts.cu, 3 tc_attn_ts.cu with Q in TMEM

// 2^x through one MUFU.EX2 (2 ulp; results below the normal range flush to zero, which is what a
// probability that small should do).  The libm exp2f spends ~6 more instructions on range handling.
__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

struct AttnParams {
  int n_img, Pq, Pk;
  int q_tiles, dv_tiles, k_tiles, k_tiles1, num_items;   // k_tiles: 64-key tiles (pass 2); k_tiles1: 128-key tiles (pass 1)
  int qt_begin;             // first query tile of this launch (q_tiles counts the tiles of the launch)
  // tc_attn_ts.cu only: a launch may end with 128-channel-wide "tail" items (the ragged last round of a 256-wide
  // launch): items [0, items_a) are decoded with (dv_tiles, qt_begin, q_tiles) and are DVT wide, items
  // [items_a, num_items) with (dv_tiles_b, qt_begin_b, q_tiles_b) and are 128 wide.  items_a == num_items: no tail.
  int items_a, dv_tiles_b, qt_begin_b, q_tiles_b;
  // tc_attn_ts.cu: the items [0, items_a) are dealt out in contiguous blocks of per_cta items per CTA, so that consecutive
  // items of a CTA are the d_v slices of the same query tile whenever possible and share its row maxima (pass 1 runs once
  // per query tile and CTA instead of once per item); tail items go one per CTA and round.  0: strided walk (tc_attn.cu).
  int per_cta;
  float scale_log2;         // log2(e) / sqrt(d_k)
  __half* out_hi;
  __half* out_lo;
  float* out_f32;
  long long o_bs, o_ld;     // batch stride / row pitch (elements)
  const __half* res_hi;
  const __half* res_lo;
  const float* res_f32;
  long long r_bs, r_ld;
  int* range_flag;
  int fast;                 // TDN_TC_FLAG_FAST (3): bit 0 Qhi.Khi^T only, bit 1 Phi.V'hi^T only (opt-in, not fp32-faithful)
};


// One work item of a (possibly mixed-width) launch.
struct AttnItem { int qt, img, dv0, halves; };
template <int DVT>
__device__ __forceinline__ AttnItem attn_item(const AttnParams& p, int item) {
  AttnItem it;
  if (item < p.items_a) {
    const int t = item / p.dv_tiles;
    it.dv0 = (item - t * p.dv_tiles) * DVT;
    it.img = t / p.q_tiles;
    it.qt = p.qt_begin + (t - it.img * p.q_tiles);
    it.halves = DVT / AT_DVH;
  } else {
    const int j = item - p.items_a;
    const int t = j / p.dv_tiles_b;
    it.dv0 = (j - t * p.dv_tiles_b) * AT_DVH;
    it.img = t / p.q_tiles_b;
    it.qt = p.qt_begin_b + (t - it.img * p.q_tiles_b);
    it.halves = 1;
  }
  return it;
}

// k-th work item of this CTA in the blocked walk described at AttnParams::per_cta (-1: no more items).
__device__ __forceinline__ int attn_walk(const AttnParams& p, int k) {
  const int c = (int)blockIdx.x, G = (int)gridDim.x;
  const int base = c * p.per_cta;
  int n_main = p.items_a - base;
  n_main = n_main < 0 ? 0 : (n_main > p.per_cta ? p.per_cta : n_main);
  if (k < n_main) return base + k;
  const int t = p.items_a + c + (k - n_main) * G;
  return t < p.num_items ? t : -1;
}
// Item `item` may reuse the row maxima of the item this CTA processed just before it: same image and query tile.
__device__ __forceinline__ bool attn_shares_rowmax(const AttnParams& p, int item, int prev) {
  return prev >= 0 && item < p.items_a && prev < p.items_a && item / p.dv_tiles == prev / p.dv_tiles;
}

// tc_attn_ts.cu: launch of the TMEM-operand kernel family (dvt = 128 or 256 output channels per work item).
cudaError_t attention_ts_launch(int dvt, bool qt, int grid, cudaStream_t stream, bool short_launch, const CUtensorMap& mq_h,
                                const CUtensorMap& mq_l, const CUtensorMap& mk_h, const CUtensorMap& mk_l,
                                const CUtensorMap& mv_h, const CUtensorMap& mv_l, const AttnParams& p);

// tc_attn_s128.cu: the TS kernel with 128-key S MMAs in pass 2 (one S buffer of 128 TMEM columns, two P slots).
cudaError_t attention_s128_launch(int dvt, int grid, cudaStream_t stream, bool short_launch, const CUtensorMap& mq_h,
                                  const CUtensorMap& mq_l, const CUtensorMap& mk_h, const CUtensorMap& mk_l,
                                  const CUtensorMap& mv_h, const CUtensorMap& mv_l, const AttnParams& p);

}  // namespace tdn
