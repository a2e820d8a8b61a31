// EXPERIMENTAL, NOT PART OF THE PRODUCT LIBRARY, NEVER EXECUTED (written after the last GPU minute of round 1).
// Second variant of tc_attn_cluster.cu (read its header first; that variant ran and is correct but slow).  Differences:
//   * THREE P buffers instead of two (the V'^T ring shrinks from 3 to 2 stages to pay for it): tile kt lives in buffer
//     kt % 3 and is produced by CTA kt & 1, so a producer may write tile kt as soon as P.V(kt-3) has retired in both
//     CTAs -- two P.V tile times of slack for the hand-off instead of one;
//   * the hand-off is always the bulk copy: the 8 softmax warps write the tile locally and arrive on `p_written[b]`;
//     one thread then arrives on the local `p_full[b]` and forwards the 32 KB tile with cp.async.bulk.shared::cluster,
//     complete_tx on the peer's `p_full[b]`, which the peer's P.V issuer arms with expect_tx just before it waits
//     (transaction bytes may land before the expect_tx of the same phase).
// `p_empty[b]` completes one phase per use of buffer b (two multicast commits), whoever produced the tile; producers
// count the uses of every buffer, their own and the peer's, to know the parity to wait for.
#include "../../tdnet_b200/csrc/common.cuh"
#include "../../tdnet_b200/csrc/tc_common.cuh"

#include <string.h>
#include <cuda.h>

namespace tdn {
namespace experimental3 {

using namespace ptx;

// ---- cluster helpers local to this file
// shared::cta -> shared::cluster bulk copy; completion is signalled as transaction bytes on the REMOTE mbarrier
__device__ __forceinline__ void bulk_copy_to_cluster(uint32_t dst_cluster_addr, const void* src_local, uint32_t bytes,
                                                     uint32_t mbar_cluster_addr) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst_cluster_addr), "r"(smem_u32(src_local)), "r"(bytes), "r"(mbar_cluster_addr)
               : "memory");
}
__device__ __forceinline__ uint32_t map_to_cta(const void* smem_ptr, uint32_t cta) {   // shared::cluster address of the
  uint32_t r;                                                                          // same offset in CTA `cta`
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(smem_ptr)), "r"(cta));
  return r;
}
__device__ __forceinline__ void st_cluster_f32(uint32_t addr, float v) {
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ void st_cluster_v4(uint32_t addr, uint4 v) {
  asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {           // release at cluster scope
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
// every thread that wrote into the peer's shared memory orders those writes at cluster scope itself, before the
// warp-level hand-off to the lane that arrives on the peer's barrier
__device__ __forceinline__ void fence_cluster() { asm volatile("fence.acq_rel.cluster;" ::: "memory"); }
__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {  // acquire at cluster scope
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t"
      "}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait_cluster(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait_cluster(bar, parity)) {
    if (clock64() - t0 > 4000000000ll) {
      printf("tc_attn_cluster: mbarrier wait timed out (block %d thread %d parity %u)\n", blockIdx.x, threadIdx.x, parity);
      __trap();
    }
  }
}
// arrive on the barrier at this offset in both CTAs of the cluster when all earlier MMAs of this thread are done
__device__ __forceinline__ void umma_commit_both(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(smem_u32(bar)), "h"((uint16_t)3)
      : "memory");
}

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
constexpr int AT_KSTAGES = 2, AT_VSTAGES = 2, AT_PBUFS = 3;
constexpr int AT_SMEM_DATA = 2 * AT_Q_PLANE + AT_KSTAGES * 2 * AT_K_PLANE + AT_VSTAGES * 2 * AT_V_PLANE + AT_PBUFS * 2 * AT_P_PLANE;
constexpr int AT_TMEM_COLS = 512;   // S: 2 x 64 columns, O: up to 256 columns
constexpr float AT_P_SCALE = 1024.f;

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
};

struct AttnBars {
  uint64_t q_full, q_empty;
  uint64_t k_full[AT_KSTAGES], k_empty[AT_KSTAGES];
  uint64_t v_full[AT_VSTAGES], v_empty[AT_VSTAGES];
  uint64_t s_full[2], s_empty[2];
  uint64_t p_written[AT_PBUFS], p_full[AT_PBUFS], p_empty[AT_PBUFS];
  uint64_t o_full, o_empty;
  uint32_t tmem_ptr;
  float xch[2][AT_BQ];     // row max / row sum exchange between the two softmax warp groups
  uint64_t xm_full, xl_full;   // the peer CTA has written its partial row maxima / row sums below
  float xm[AT_BQ], xl[AT_BQ];  // written by the PEER through st.shared::cluster
};

// no alignment slack here (the two exchange arrays need that kilobyte): the dynamic shared memory is declared
// __align__(1024) and the kernel traps if the base is not 1024-byte aligned
constexpr int AT_SMEM_BYTES = AT_SMEM_DATA + ((int)sizeof(AttnBars) + 127) / 128 * 128;
static_assert(AT_SMEM_BYTES <= 232448, "attention kernel exceeds the 227 KB shared-memory limit");

template <int DVT>   // d_v slice per CTA: 256
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(AT_THREADS, 1)
tc_attn_cluster3_kernel(const __grid_constant__ CUtensorMap tmQ_hi, const __grid_constant__ CUtensorMap tmQ_lo,
               const __grid_constant__ CUtensorMap tmK_hi, const __grid_constant__ CUtensorMap tmK_lo,
               const __grid_constant__ CUtensorMap tmV_hi, const __grid_constant__ CUtensorMap tmV_lo,
               const AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if ((smem_u32(smem) & 1023u) != 0) __trap();   // the 128B-swizzled operand tiles need a 1024-byte aligned base
  uint8_t* sQ = smem;                                            // hi | lo
  uint8_t* sK = sQ + 2 * AT_Q_PLANE;                             // stages x (hi | lo)
  uint8_t* sV = sK + AT_KSTAGES * 2 * AT_K_PLANE;
  uint8_t* sP = sV + AT_VSTAGES * 2 * AT_V_PLANE;                // 3 buffers x (hi | lo)
  AttnBars* bars = reinterpret_cast<AttnBars*>(sP + AT_PBUFS * 2 * AT_P_PLANE);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();       // 0 / 1: the d_v slice of the pair and the parity of "own" key tiles
  const uint32_t peer = rank ^ 1u;
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tmQ_hi); prefetch_tensormap(&tmQ_lo);
    prefetch_tensormap(&tmK_hi); prefetch_tensormap(&tmK_lo);
    prefetch_tensormap(&tmV_hi); prefetch_tensormap(&tmV_lo);
    mbar_init(&bars->q_full, 1);
    mbar_init(&bars->q_empty, 1);
    for (int s = 0; s < AT_KSTAGES; ++s) { mbar_init(&bars->k_full[s], 1); mbar_init(&bars->k_empty[s], 1); }
    for (int s = 0; s < AT_VSTAGES; ++s) { mbar_init(&bars->v_full[s], 1); mbar_init(&bars->v_empty[s], 1); }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&bars->s_full[s], 1);
      mbar_init(&bars->s_empty[s], AT_SOFTMAX_WARPS);   // one arrival per softmax warp (lane 0 after __syncwarp)
    }
    for (int s = 0; s < AT_PBUFS; ++s) {
      mbar_init(&bars->p_written[s], AT_SOFTMAX_WARPS); // local tile: the 8 softmax warps have written buffer s
      mbar_init(&bars->p_full[s], 1);                   // local tile: the forwarder's arrive; remote tile: expect_tx + bytes
      mbar_init(&bars->p_empty[s], 2);                  // released by the P.V issuers of both CTAs
    }
    mbar_init(&bars->o_full, 1);
    mbar_init(&bars->o_empty, AT_SOFTMAX_WARPS);
    mbar_init(&bars->xm_full, AT_SOFTMAX_WARPS / 2);    // the peer's four group-0 warps
    mbar_init(&bars->xl_full, AT_SOFTMAX_WARPS / 2);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(&bars->tmem_ptr, AT_TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                            // the peer's barriers are initialised before anything arrives on them
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_ptr;
  tc_pdl_sync();
  const int T_own = (p.k_tiles - (int)rank + 1) / 2;     // 64-key tiles kt with kt % 2 == rank
  const int T1_own = (p.k_tiles1 - (int)rank + 1) / 2;   // 128-key pass-1 tiles likewise
  const uint32_t tmem_S = tmem_base;            // + buf * 64
  const uint32_t tmem_O = tmem_base + 128;
  const int T = p.k_tiles;
  const int T1 = p.k_tiles1;
  constexpr int HALVES = DVT / AT_DVH;

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      int ks = 0, vs = 0;
      uint32_t kph = 0, vph = 0, qph = 0;
      for (int item = cluster_id; item < p.num_items; item += num_clusters) {
        const int dvt = 2 * (item % p.dv_tiles) + (int)rank;      // p.dv_tiles counts PAIRS of 256-wide slices
        int t = item / p.dv_tiles;
        const int qt = p.qt_begin + t % p.q_tiles;
        const int img = t / p.q_tiles;
        mbar_wait(&bars->q_empty, qph ^ 1);
        mbar_expect_tx(&bars->q_full, 2 * AT_Q_PLANE);
        tma_load_3d(sQ, &tmQ_hi, &bars->q_full, 0, qt * AT_BQ, img);
        tma_load_3d(sQ + AT_Q_PLANE, &tmQ_lo, &bars->q_full, 0, qt * AT_BQ, img);
        qph ^= 1;
        // pass 1: the hi plane of the keys only (S~ = Qhi.Khi^T), 128 keys per stage: two 64-key boxes land
        // back to back = one 128-row swizzled tile (a box past the last key is zero-filled)
        for (int kt = (int)rank; kt < T1; kt += 2) {                // own pass-1 tiles only
          mbar_wait(&bars->k_empty[ks], kph ^ 1);
          uint8_t* dst = sK + ks * 2 * AT_K_PLANE;
          mbar_expect_tx(&bars->k_full[ks], 2 * AT_K_PLANE);
          tma_load_3d(dst, &tmK_hi, &bars->k_full[ks], 0, kt * AT_BK1, img);
          tma_load_3d(dst + AT_K_PLANE, &tmK_hi, &bars->k_full[ks], 0, kt * AT_BK1 + AT_BK, img);
          if (++ks == AT_KSTAGES) { ks = 0; kph ^= 1; }
        }
        // pass 2: keys (hi+lo) and the V'^T slice (hi+lo).  The key tile is fetched ONE TILE AHEAD of the
        // values: S(kt+1) is issued before P.V(kt), and a V stage only frees up when P.V(kt-1) retires, so a
        // K load queued behind the V loads would arrive a TMA latency too late and stall the tensor pipe.
        auto load_k = [&](int kt) {
          mbar_wait(&bars->k_empty[ks], kph ^ 1);
          uint8_t* dk = sK + ks * 2 * AT_K_PLANE;
          mbar_expect_tx(&bars->k_full[ks], 2 * AT_K_PLANE);
          tma_load_3d(dk, &tmK_hi, &bars->k_full[ks], 0, kt * AT_BK, img);
          tma_load_3d(dk + AT_K_PLANE, &tmK_lo, &bars->k_full[ks], 0, kt * AT_BK, img);
          if (++ks == AT_KSTAGES) { ks = 0; kph ^= 1; }
        };
        // own key tiles (kt % 2 == rank), requested up to two tiles ahead of the values; the values of EVERY key tile
        int nk = (int)rank;
        for (int kt = 0; kt < T; ++kt) {
          while (nk < T && nk <= kt + 2) { load_k(nk); nk += 2; }
          for (int h = 0; h < HALVES; ++h) {
            mbar_wait(&bars->v_empty[vs], vph ^ 1);
            uint8_t* dv = sV + vs * 2 * AT_V_PLANE;
            mbar_expect_tx(&bars->v_full[vs], 2 * AT_V_PLANE);
            tma_load_3d(dv, &tmV_hi, &bars->v_full[vs], kt * AT_BK, dvt * DVT + h * AT_DVH, img);
            tma_load_3d(dv + AT_V_PLANE, &tmV_lo, &bars->v_full[vs], kt * AT_BK, dvt * DVT + h * AT_DVH, img);
            if (++vs == AT_VSTAGES) { vs = 0; vph ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer 1: S = Q.K^T (both passes) ================================
    // The whole warp runs the loop and the barrier waits so that stage indices, phases and descriptors
    // stay warp-uniform (uniform registers feed tcgen05.mma directly); one elected lane issues.
    constexpr uint32_t idesc_s = umma_idesc_f16(AT_BQ, AT_BK);    // 128 x 64  (pass 2)
    constexpr uint32_t idesc_s1 = umma_idesc_f16(AT_BQ, AT_BK1);  // 128 x 128 (pass 1)
    int ks = 0;
    uint32_t kph = 0, qph = 0, oph = 0;
    uint32_t sn = 0;                                             // S tiles issued so far: buffer sn & 1, phase (sn >> 1) & 1
    const uint32_t q_hi = smem_u32(sQ), q_lo = q_hi + AT_Q_PLANE;
    for (int item = cluster_id; item < p.num_items; item += num_clusters) {
      mbar_wait(&bars->q_full, qph);
      // Pass-1 tiles are 128 keys wide: buffer 0 = the S columns, buffer 1 = the first 128 O columns, which are
      // idle until P.V starts -- once the epilogue of the previous item has read them.
      mbar_wait(&bars->o_empty, oph ^ 1);
      for (int it = 0; it < T1_own; ++it, ++sn) {
        const int sb = sn & 1;
        mbar_wait(&bars->k_full[ks], kph);
        mbar_wait(&bars->s_empty[sb], ((sn >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t k_hi = smem_u32(sK + ks * 2 * AT_K_PLANE);
        const uint32_t d = sb ? tmem_O : tmem_S;
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < AT_DK / 16; ++k)
            umma_f16(d, umma_desc_k_sw128(q_hi + k * 32), umma_desc_k_sw128(k_hi + k * 32), idesc_s1, k != 0);
          umma_commit(&bars->s_full[sb]);
          umma_commit(&bars->k_empty[ks]);
        }
        __syncwarp();
        if (++ks == AT_KSTAGES) { ks = 0; kph ^= 1; }
      }
      // pass 2 re-uses the S columns as two 64-column buffers: the last pass-1 tile must have been read (the one
      // before it is covered by the regular s_empty wait of the first pass-2 tile)
      if (T1_own > 0) mbar_wait(&bars->s_empty[(sn - 1) & 1], ((sn - 1) >> 1) & 1);
      if (T_own == 0 && elect_one()) umma_commit(&bars->q_empty);   // no own pass-2 tile: Q is free after pass 1
      __syncwarp();
      for (int it = 0; it < T_own; ++it, ++sn) {
        const int sb = sn & 1;
        mbar_wait(&bars->k_full[ks], kph);
        mbar_wait(&bars->s_empty[sb], ((sn >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t k_hi = smem_u32(sK + ks * 2 * AT_K_PLANE), k_lo = k_hi + AT_K_PLANE;
        const uint32_t d = tmem_S + sb * AT_BK;
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < AT_DK / 16; ++k) {
            const uint64_t a_h = umma_desc_k_sw128(q_hi + k * 32), b_h = umma_desc_k_sw128(k_hi + k * 32);
            const uint64_t a_l = umma_desc_k_sw128(q_lo + k * 32), b_l = umma_desc_k_sw128(k_lo + k * 32);
            umma_f16(d, a_h, b_l, idesc_s, k != 0);
            umma_f16(d, a_l, b_h, idesc_s, 1);
            umma_f16(d, a_h, b_h, idesc_s, 1);
          }
          umma_commit(&bars->s_full[sb]);
          umma_commit(&bars->k_empty[ks]);
          if (it == T_own - 1) umma_commit(&bars->q_empty);        // Q tile free once the last own S has retired
        }
        __syncwarp();
        if (++ks == AT_KSTAGES) { ks = 0; kph ^= 1; }
      }
      qph ^= 1;
      oph ^= 1;
    }
  } else if (warp == AT_PV_WARP) {
    // ================================ MMA issuer 2: O += P.V'^T ================================
    constexpr uint32_t idesc_o = umma_idesc_f16(AT_BQ, AT_DVH);  // 128 x 128
    int vs = 0;
    uint32_t vph = 0, oph = 0;
    uint32_t pmask = 0;                                          // bit b: phase of p_full[b] expected next
    for (int item = cluster_id; item < p.num_items; item += num_clusters) {
      mbar_wait(&bars->o_empty, oph ^ 1);                        // epilogue of the previous item has read O
      int pb = 0;                                                // kt % 3
      for (int kt = 0; kt < T; ++kt) {
        if ((uint32_t)(kt & 1) != rank) {                        // the peer forwards this tile: arm its bytes
          if (elect_one()) mbar_expect_tx(&bars->p_full[pb], 2 * AT_P_PLANE);
          __syncwarp();
        }
        mbar_wait_cluster(&bars->p_full[pb], (pmask >> pb) & 1u);
        const uint32_t p_hi = smem_u32(sP + pb * 2 * AT_P_PLANE), p_lo = p_hi + AT_P_PLANE;
        for (int h = 0; h < HALVES; ++h) {
          mbar_wait(&bars->v_full[vs], vph);
          tc_fence_after();
          const uint32_t v_hi = smem_u32(sV + vs * 2 * AT_V_PLANE), v_lo = v_hi + AT_V_PLANE;
          const uint32_t d = tmem_O + h * AT_DVH;
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < AT_BK / 16; ++k) {
              const uint64_t a_h = umma_desc_k_sw128(p_hi + k * 32), a_l = umma_desc_k_sw128(p_lo + k * 32);
              const uint64_t b_h = umma_desc_k_sw128(v_hi + k * 32), b_l = umma_desc_k_sw128(v_lo + k * 32);
              umma_f16(d, a_h, b_l, idesc_o, (kt | k) != 0);
              umma_f16(d, a_l, b_h, idesc_o, 1);
              umma_f16(d, a_h, b_h, idesc_o, 1);
            }
            umma_commit(&bars->v_empty[vs]);
            if (h == HALVES - 1) {
              umma_commit_both(&bars->p_empty[pb]);             // the producer of buffer pb may be the peer CTA
              if (kt == T - 1) umma_commit(&bars->o_full);
            }
          }
          __syncwarp();
          if (++vs == AT_VSTAGES) { vs = 0; vph ^= 1; }
        }
        pmask ^= 1u << pb;
        if (++pb == AT_PBUFS) pb = 0;
      }
      oph ^= 1;
    }
  } else {
    // ================================ softmax + epilogue warps ================================
    // Two warps per TMEM lane quarter: group g (warps 2-5 / 6-9) owns key columns [32g, 32g+32) of every
    // 64-key tile and output channels [g*DVT/2, (g+1)*DVT/2) of the O tile.  Row max and row sum are
    // exchanged through shared memory (named barrier 1 over the 256 softmax threads).
    const int quarter = warp & 3;
    const int group = (warp - 2) >> 2;
    const int row = quarter * 32 + lane;                      // query row inside the tile = TMEM lane
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    uint32_t oph = 0, xph = 0;
    uint32_t emask = 0;                                         // bit b: uses of P buffer b so far (own and peer tiles), mod 2
    uint32_t wmask = 0;                                         // bit b: tiles THIS CTA has written into buffer b so far, mod 2
    uint32_t sn = 0;                                            // OWN S tiles consumed so far (same counting as MMA issuer 1)
    const uint32_t peer_xm = map_to_cta(&bars->xm[0], peer), peer_xl = map_to_cta(&bars->xl[0], peer);
    const uint32_t peer_xm_full = map_to_cta(&bars->xm_full, peer), peer_xl_full = map_to_cta(&bars->xl_full, peer);
    bool out_of_range = false;
    auto group_sync = [] { asm volatile("bar.sync 1, 256;" ::: "memory"); };
    for (int item = cluster_id; item < p.num_items; item += num_clusters) {
      const int dvt = 2 * (item % p.dv_tiles) + (int)rank;
      int t = item / p.dv_tiles;
      const int qt = p.qt_begin + t % p.q_tiles;
      const int img = t / p.q_tiles;
      const int q_idx = qt * AT_BQ + row;
      const bool valid = q_idx < p.Pq;

      // ---- pass 1: row maximum of S~; 128-key tiles, this group's 64 key columns of each
      float m = -INFINITY;
      for (int kt = (int)rank; kt < T1; kt += 2, ++sn) {
        const int sb = sn & 1;
        mbar_wait(&bars->s_full[sb], (sn >> 1) & 1);
        tc_fence_after();
        uint32_t r0[32], r1[32];
        const uint32_t src = (sb ? tmem_O : tmem_S) + lane_addr + group * 64;
        tmem_ld_32x32(src, r0);
        tmem_ld_32x32(src + 32, r1);
        tmem_ld_wait();
        const int kbase = kt * AT_BK1 + group * 64;
        if (kbase + 64 <= p.Pk) {                               // only the last key tile can be ragged
          float m0 = m, m1 = -INFINITY;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            m0 = fmaxf(m0, __uint_as_float(r0[j]));
            m1 = fmaxf(m1, __uint_as_float(r1[j]));
          }
          m = fmaxf(m0, m1);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            if (kbase + j < p.Pk) m = fmaxf(m, __uint_as_float(r0[j]));
            if (kbase + 32 + j < p.Pk) m = fmaxf(m, __uint_as_float(r1[j]));
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->s_empty[sb]);
      }
      bars->xch[group][row] = m;
      group_sync();
      m = fmaxf(m, bars->xch[group ^ 1][row]);
      group_sync();                                           // xch is reused for the row sums below
      // m covers this CTA's key tiles only: hand it to the peer, take the peer's
      if (group == 0) {
        st_cluster_f32(peer_xm + row * 4, m);
        fence_cluster();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(peer_xm_full);
      }
      mbar_wait_cluster(&bars->xm_full, xph);
      m = fmaxf(m, bars->xm[row]);
      // exponent offset of pass 2: the row maximum AND log2 of the 2^10 probability scale, so that one FMA + one
      // MUFU.EX2 yield p * 2^10 directly (the row sum l is then scaled by 2^10 as well: out = O / l)
      const float m_scaled = m * p.scale_log2 - 10.f;
      static_assert(AT_P_SCALE == 1024.f, "the exponent offset above assumes a 2^10 probability scale");

      // ---- pass 2: probabilities -> shared memory (UMMA K-major, 128B swizzle), partial row sum
      float l = 0.f;
      int pb = 0;                                               // kt % 3
      for (int kt = 0; kt < T; ++kt, pb = (pb + 1 == AT_PBUFS ? 0 : pb + 1)) {
        if ((uint32_t)(kt & 1) != rank) {                       // the peer's tile: only the buffer's use count advances
          emask ^= 1u << pb;
          continue;
        }
        const int sb = sn & 1;
        mbar_wait(&bars->s_full[sb], (sn >> 1) & 1);
        tc_fence_after();
        float pr[32];
        {
          uint32_t r[32];
          tmem_ld_32x32(tmem_S + lane_addr + sb * AT_BK + group * 32, r);
          tmem_ld_wait();
          const int kbase = kt * AT_BK + group * 32;
#pragma unroll
          for (int j = 0; j < 32; ++j) pr[j] = fast_exp2(fmaf(__uint_as_float(r[j]), p.scale_log2, -m_scaled));
          if (kbase + 32 > p.Pk) {                              // ragged last tile: keys past P' contribute nothing
#pragma unroll
            for (int j = 0; j < 32; ++j) pr[j] = (kbase + j < p.Pk) ? pr[j] : 0.f;
          }
#pragma unroll
          for (int j = 0; j < 32; ++j) l += pr[j];
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->s_empty[sb]);

        // split before waiting for the buffer: the conversions overlap the P.V MMAs that still read it
        uint4 phv[4], plv[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          __half2 hi[4], lo[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) split_f32x2(pr[c * 8 + 2 * e], pr[c * 8 + 2 * e + 1], hi[e], lo[e]);
          phv[c] = *reinterpret_cast<const uint4*>(hi);
          plv[c] = *reinterpret_cast<const uint4*>(lo);
        }
        mbar_wait_cluster(&bars->p_empty[pb], ((emask >> pb) & 1u) ^ 1u);   // the previous use of this buffer is released
        uint8_t* pbuf = sP + pb * 2 * AT_P_PLANE;
        uint8_t* ph = pbuf + row * 128;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int phys = ((group * 4 + c) ^ (row & 7)) << 4;   // 16-byte chunk inside the 128-byte swizzled row
          *reinterpret_cast<uint4*>(ph + phys) = phv[c];
          *reinterpret_cast<uint4*>(ph + AT_P_PLANE + phys) = plv[c];
        }
        fence_proxy_async_smem();                              // local writes -> async proxy (UMMA and the bulk copy)
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->p_written[pb]);
        if (warp == 2 && lane == 0) {
          // forwarder: all 8 warps have written the tile -> release it to the local P.V issuer and send one
          // asynchronous 32 KB copy into the peer's buffer pb (completion = transaction bytes on the peer's p_full[pb])
          mbar_wait(&bars->p_written[pb], (wmask >> pb) & 1u);
          mbar_arrive(&bars->p_full[pb]);
          bulk_copy_to_cluster(map_to_cta(pbuf, peer), pbuf, 2 * AT_P_PLANE, map_to_cta(&bars->p_full[pb], peer));
        }
        __syncwarp();
        emask ^= 1u << pb;
        wmask ^= 1u << pb;
        ++sn;
      }
      bars->xch[group][row] = l;
      group_sync();
      l += bars->xch[group ^ 1][row];
      group_sync();
      // l covers this CTA's key tiles only: exchange with the peer (same protocol as the row maxima)
      if (group == 0) {
        st_cluster_f32(peer_xl + row * 4, l);
        fence_cluster();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(peer_xl_full);
      }
      mbar_wait_cluster(&bars->xl_full, xph);
      l += bars->xl[row];
      xph ^= 1;

      // ---- epilogue: out = O / l + residual; this group's half of the channel slice.  The SPLIT16 residual of
      //      chunk c+1 is requested before chunk c is processed (and chunk 0 before O is even complete): its
      //      global-memory latency was the longest serial piece of an item.
      constexpr int COLS = DVT / 2;
      constexpr int NCHUNK = COLS / 32;
      const int cbase = dvt * DVT + group * COLS;
      const long long obase = (long long)img * p.o_bs + (long long)q_idx * p.o_ld + cbase;
      const long long rbase = (long long)img * p.r_bs + (long long)q_idx * p.r_ld + cbase;
      const bool res16 = valid && p.res_hi != nullptr;
      uint4 rbuf[2][8];                                        // [buffer][4 x hi | 4 x lo] = 32 channels
      auto load_res = [&](int chunk, uint4 (&dst)[8]) {
        if (res16) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            dst[q] = __ldg(reinterpret_cast<const uint4*>(p.res_hi + rbase + chunk * 32 + q * 8));
            dst[4 + q] = __ldg(reinterpret_cast<const uint4*>(p.res_lo + rbase + chunk * 32 + q * 8));
          }
        }
      };
      load_res(0, rbuf[0]);
      mbar_wait(&bars->o_full, oph);
      tc_fence_after();
      oph ^= 1;
      const float inv = 1.f / l;                               // l carries the 2^10 scale of P
#pragma unroll
      for (int chunk = 0; chunk < NCHUNK; ++chunk) {
        if (chunk + 1 < NCHUNK) load_res(chunk + 1, rbuf[(chunk + 1) & 1]);
        uint32_t r[32];
        tmem_ld_32x32(tmem_O + lane_addr + group * COLS + chunk * 32, r);
        tmem_ld_wait();
        if (valid) {
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) * inv;
          const int c0 = chunk * 32;
          if (p.res_hi) {
            const uint4 (&rb)[8] = rbuf[chunk & 1];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const __half2* hh = reinterpret_cast<const __half2*>(&rb[q]);
              const __half2* ll = reinterpret_cast<const __half2*>(&rb[4 + q]);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                float2 a = __half22float2(hh[e]), b2 = __half22float2(ll[e]);
                v[q * 8 + e * 2 + 0] += a.x + b2.x;
                v[q * 8 + e * 2 + 1] += a.y + b2.y;
              }
            }
          } else if (p.res_f32) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              float4 f = *reinterpret_cast<const float4*>(p.res_f32 + rbase + c0 + q * 4);
              v[q * 4 + 0] += f.x; v[q * 4 + 1] += f.y; v[q * 4 + 2] += f.z; v[q * 4 + 3] += f.w;
            }
          }
          if (p.out_f32) {
#pragma unroll
            for (int q = 0; q < 8; ++q)
              *reinterpret_cast<float4*>(p.out_f32 + obase + c0 + q * 4) =
                  make_float4(v[q * 4 + 0], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
          }
          if (p.out_hi) {
            __half2 hi[16], lo[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              out_of_range |= fmaxf(fabsf(v[2 * j]), fabsf(v[2 * j + 1])) > 60000.f;
              split_f32x2(v[2 * j], v[2 * j + 1], hi[j], lo[j]);
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              *reinterpret_cast<uint4*>(p.out_hi + obase + c0 + q * 8) = *reinterpret_cast<const uint4*>(&hi[q * 4]);
              *reinterpret_cast<uint4*>(p.out_lo + obase + c0 + q * 8) = *reinterpret_cast<const uint4*>(&lo[q * 4]);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->o_empty);
    }
    if (out_of_range && p.range_flag) atomicOr(p.range_flag, 1);
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                            // neither CTA exits while the peer may still write into its memory
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, AT_TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------
}  // namespace experimental3
int encode_map_f16(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims,
                   const cuuint64_t* strides_bytes, const cuuint32_t* box, const char* what,
                   const cuuint32_t* elem_strides, int swizzle128 = 1);
namespace experimental3 {

int attention_tc_cluster3(const tdn_attention_desc* d, cudaStream_t stream) {
  TDN_REQUIRE(d->q_hi && d->q_lo && d->k_hi && d->k_lo && d->vt_hi && d->vt_lo, TDN_ERR_INVALID,
              "attention_tc: null operand");
  TDN_REQUIRE(d->d_k == AT_DK, TDN_ERR_UNSUPPORTED, "attention_tc: d_k must be 64 (got %d)", d->d_k);
  TDN_REQUIRE(d->d_v % 512 == 0, TDN_ERR_UNSUPPORTED, "attention_tc_cluster: d_v=%d must be a multiple of 512", d->d_v);
  // 256-wide slices halve the QK^T / softmax recompute; small problems (the FIFO hops with P' queries) would
  // not fill the SMs with them, so they take 128-wide slices = twice as many work items.
  static int num_sms_cached = 0;
  if (num_sms_cached == 0) {
    int dev = 0;
    TDN_CUDA_OK(cudaGetDevice(&dev));
    TDN_CUDA_OK(cudaDeviceGetAttribute(&num_sms_cached, cudaDevAttrMultiProcessorCount, dev));
  }
  const int dvt_size = 256;
  TDN_REQUIRE(d->n > 0 && d->pq > 0 && d->pk > 0, TDN_ERR_INVALID, "attention_tc: empty problem");
  TDN_REQUIRE(d->vt_ld % 8 == 0 && d->vt_ld >= ((d->pk + 63) / 64) * 64, TDN_ERR_INVALID,
              "attention_tc: V'^T row pitch must cover the keys padded to 64 (zero-filled) and be 16-byte aligned");
  TDN_REQUIRE(d->q_ld % 8 == 0 && d->k_ld % 8 == 0 && d->q_batch_stride % 8 == 0 && d->k_batch_stride % 8 == 0 &&
                  d->vt_batch_stride % 8 == 0, TDN_ERR_INVALID, "attention_tc: operand pitches must be 16-byte aligned");
  const tdn_tensor& out = d->out;
  TDN_REQUIRE(out.data && out.n == d->n && out.h == 1 && out.w == d->pq && out.c == d->d_v, TDN_ERR_INVALID,
              "attention_tc: out must be a [n,1,pq,d_v] token view");
  AttnParams p;
  memset(&p, 0, sizeof(p));
  p.n_img = d->n; p.Pq = d->pq; p.Pk = d->pk;
  p.q_tiles = ceil_div(d->pq, AT_BQ);
  p.dv_tiles = d->d_v / (2 * dvt_size);          // pairs of 256-wide slices, one pair per cluster
  p.k_tiles = ceil_div(d->pk, AT_BK);
  p.k_tiles1 = ceil_div(d->pk, AT_BK1);
  long long items = (long long)d->n * p.q_tiles * p.dv_tiles;
  TDN_REQUIRE(items < (1ll << 31), TDN_ERR_UNSUPPORTED, "attention_tc: too many work items");
  p.num_items = (int)items;
  p.scale_log2 = 1.4426950408889634f / sqrtf((float)d->d_k);
  if (out.dtype == TDN_SPLIT16) {
    TDN_REQUIRE(out.data_lo && aligned16(out.data) && aligned16(out.data_lo) && out.stride_w % 8 == 0 &&
                    out.stride_n % 8 == 0, TDN_ERR_INVALID, "attention_tc: misaligned SPLIT16 output");
    p.out_hi = (__half*)out.data; p.out_lo = (__half*)out.data_lo;
  } else {
    TDN_REQUIRE(aligned16(out.data) && out.stride_w % 4 == 0 && out.stride_n % 4 == 0, TDN_ERR_INVALID,
                "attention_tc: misaligned fp32 output");
    p.out_f32 = (float*)out.data;
  }
  p.o_bs = out.stride_n; p.o_ld = out.stride_w;
  if (d->residual.data) {
    const tdn_tensor& r = d->residual;
    TDN_REQUIRE(r.n == out.n && r.h == 1 && r.w == out.w && r.c == out.c, TDN_ERR_INVALID,
                "attention_tc: residual dims must equal output dims");
    if (r.dtype == TDN_SPLIT16) {
      TDN_REQUIRE(r.data_lo && aligned16(r.data) && aligned16(r.data_lo) && r.stride_w % 8 == 0 && r.stride_n % 8 == 0,
                  TDN_ERR_INVALID, "attention_tc: misaligned residual");
      p.res_hi = (const __half*)r.data; p.res_lo = (const __half*)r.data_lo;
    } else {
      TDN_REQUIRE(aligned16(r.data) && r.stride_w % 4 == 0 && r.stride_n % 4 == 0, TDN_ERR_INVALID,
                  "attention_tc: misaligned residual");
      p.res_f32 = (const float*)r.data;
    }
    p.r_bs = r.stride_n; p.r_ld = r.stride_w;
  }
  p.range_flag = d->range_flag;

  CUtensorMap mq_h, mq_l, mk_h, mk_l, mv_h, mv_l;
  int rc;
  {
    cuuint64_t dims[3] = {(cuuint64_t)AT_DK, (cuuint64_t)d->pq, (cuuint64_t)d->n};
    cuuint64_t str[2] = {(cuuint64_t)d->q_ld * 2, (cuuint64_t)(d->n > 1 ? d->q_batch_stride : d->q_ld * (long long)d->pq) * 2};
    cuuint32_t box[3] = {(cuuint32_t)AT_DK, (cuuint32_t)AT_BQ, 1};
    if ((rc = encode_map_f16(&mq_h, d->q_hi, 3, dims, str, box, "Q.hi", nullptr))) return rc;
    if ((rc = encode_map_f16(&mq_l, d->q_lo, 3, dims, str, box, "Q.lo", nullptr))) return rc;
  }
  {
    cuuint64_t dims[3] = {(cuuint64_t)AT_DK, (cuuint64_t)d->pk, (cuuint64_t)d->n};
    cuuint64_t str[2] = {(cuuint64_t)d->k_ld * 2, (cuuint64_t)(d->n > 1 ? d->k_batch_stride : d->k_ld * (long long)d->pk) * 2};
    cuuint32_t box[3] = {(cuuint32_t)AT_DK, (cuuint32_t)AT_BK, 1};
    if ((rc = encode_map_f16(&mk_h, d->k_hi, 3, dims, str, box, "K.hi", nullptr))) return rc;
    if ((rc = encode_map_f16(&mk_l, d->k_lo, 3, dims, str, box, "K.lo", nullptr))) return rc;
  }
  {
    // V'^T: [d_v rows][keys], keys contiguous; the key extent is the padded pitch so that the pad
    // columns (zeros written by the producer) are read rather than treated as out of bounds.
    cuuint64_t dims[3] = {(cuuint64_t)(((d->pk + 63) / 64) * 64), (cuuint64_t)d->d_v, (cuuint64_t)d->n};
    cuuint64_t str[2] = {(cuuint64_t)d->vt_ld * 2, (cuuint64_t)(d->n > 1 ? d->vt_batch_stride : d->vt_ld * (long long)d->d_v) * 2};
    cuuint32_t box[3] = {(cuuint32_t)AT_BK, (cuuint32_t)AT_DVH, 1};
    if ((rc = encode_map_f16(&mv_h, d->vt_hi, 3, dims, str, box, "Vt.hi", nullptr))) return rc;
    if ((rc = encode_map_f16(&mv_l, d->vt_lo, 3, dims, str, box, "Vt.lo", nullptr))) return rc;
  }
  static bool attr_set = false;
  if (!attr_set) {
    TDN_CUDA_OK(cudaFuncSetAttribute(tc_attn_cluster3_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM_BYTES));
    attr_set = true;
  }
  static int num_sms = 0;
  if (num_sms == 0) {
    int dev = 0;
    TDN_CUDA_OK(cudaGetDevice(&dev));
    TDN_CUDA_OK(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  const int max_clusters = num_sms / 2;
  const int clusters = p.num_items < max_clusters ? p.num_items : max_clusters;
  TDN_CUDA_OK(tc_launch(tc_attn_cluster3_kernel<256>, 2 * clusters, AT_THREADS, AT_SMEM_BYTES, stream,
                        p.num_items <= 2 * clusters, mq_h, mq_l, mk_h, mk_l, mv_h, mv_l, p));
  return TDN_OK;
}

}  // namespace experimental3
}  // namespace tdn

extern "C" int tdnx_attention_tc_cluster3(const tdn_attention_desc* d, void* stream) {
  if (d == nullptr) return TDN_ERR_INVALID;
  return tdn::experimental3::attention_tc_cluster3(d, (cudaStream_t)stream);
}
