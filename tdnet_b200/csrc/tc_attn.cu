// Fused attention-propagation kernel for sm_100a (tcgen05 + TMEM + TMA), fp32-faithful exact mode.
//
//   out[q, :] = softmax_k( Q[q,:] . K[k,:] / sqrt(d_k) ) @ V'[k, :]  (+ residual[q, :])
//
// replaces transformer.py:126-139 (ScaledDotProductAttention: bmm -> /temperature -> softmax -> bmm)
// and, with the fc folded into V' by the host (softmax rows sum to 1), Attention.forward :71-92.
// The [Pq x P'] attention matrix never exists in memory.
//
// Work item = (image, 128-query tile, DVT-channel slice of d_v, DVT = 256 when d_v % 256 == 0 else 128);
// persistent CTAs loop over items.
// Per item, two passes over the P' keys in tiles of 64:
//   pass 1  S~ = Qhi.Khi^T (one fp16 MMA per K step) -> running row maximum m.  The maximum is only a
//           stabiliser, so the cheap single-product S~ is enough (it is within ~1e-3*|S| of S).
//   pass 2  S  = Q.K^T in exact mode (hi*lo + lo*hi + hi*hi, fp32 TMEM accumulator);
//           p  = exp(S/sqrt(d_k) - m) (fp32, per query row in registers, row sum l accumulated);
//           P  = p * 2^10 split to fp16 hi/lo, written to shared memory in the UMMA K-major 128B-swizzled
//                layout; O += P.V'^T in exact mode (fp32 TMEM accumulator, DVT columns, N=128 MMAs).
//   epilogue out = O / (l * 2^10) + residual, re-split to hi/lo (and/or fp32).
// Because m is fixed before pass 2 there is no running rescale of O: the MMA warp never waits on a
// correction step and the result does not depend on tile order.
//
// Warp roles (352 threads): warp 0 TMA producer (Q tile, K ring, V ring); warp 1 issues the S = Q.K^T MMAs
// and warp 10 the O += P.V'^T MMAs (two independent issue streams: a single issuing thread that also has to
// poll five barriers per key tile was the measured bottleneck); warps 2-9 softmax + epilogue, two warps
// per TMEM lane quarter, each owning half of the key columns of a tile and half of the output channels
// (row max / sum exchanged through shared memory).
// TMEM: S double-buffered 2 x 64 columns, O DVT columns.  Shared memory: Q 32 KB, K ring 2 x 16 KB,
// V ring 3 x 32 KB (128-row halves of the V'^T tile), P double buffer 2 x 32 KB (hi+lo planes each).
#include "tc_attn.cuh"

#include <string.h>

namespace tdn {

struct AttnBars {
  uint64_t q_full, q_empty;
  uint64_t k_full[AT_KSTAGES], k_empty[AT_KSTAGES];
  uint64_t v_full[AT_VSTAGES], v_empty[AT_VSTAGES];
  uint64_t s_full[2], s_empty[2];
  uint64_t p_full[2], p_empty[2];
  uint64_t o_full, o_empty;
  uint32_t tmem_ptr;
  float xch[2][AT_BQ];     // row max / row sum exchange between the two softmax warp groups
};

constexpr int AT_SMEM_BYTES = AT_SMEM_DATA + 1024 /*alignment slack*/ + ((int)sizeof(AttnBars) + 127) / 128 * 128;
static_assert(AT_SMEM_BYTES <= 232448, "attention kernel exceeds the 227 KB shared-memory limit");

template <int DVT>   // d_v slice per work item: 128 or 256 (one or two 128-row V'^T halves per key tile)
__global__ void __launch_bounds__(AT_THREADS, 1)
tc_attn_kernel(const __grid_constant__ CUtensorMap tmQ_hi, const __grid_constant__ CUtensorMap tmQ_lo,
               const __grid_constant__ CUtensorMap tmK_hi, const __grid_constant__ CUtensorMap tmK_lo,
               const __grid_constant__ CUtensorMap tmV_hi, const __grid_constant__ CUtensorMap tmV_lo,
               const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sQ = smem;                                            // hi | lo
  uint8_t* sK = sQ + 2 * AT_Q_PLANE;                             // stages x (hi | lo)
  uint8_t* sV = sK + AT_KSTAGES * 2 * AT_K_PLANE;
  uint8_t* sP = sV + AT_VSTAGES * 2 * AT_V_PLANE;                // 2 buffers x (hi | lo)
  AttnBars* bars = reinterpret_cast<AttnBars*>(sP + 2 * 2 * AT_P_PLANE);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

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
      mbar_init(&bars->s_empty[s], AT_SOFTMAX_WARPS);   // one arrival per softmax warp (lane 0 after __syncwarp):
      mbar_init(&bars->p_full[s], AT_SOFTMAX_WARPS);    // 8 instead of 256 serialised shared-memory atomics per tile
      mbar_init(&bars->p_empty[s], 1);
    }
    mbar_init(&bars->o_full, 1);
    mbar_init(&bars->o_empty, AT_SOFTMAX_WARPS);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(&bars->tmem_ptr, AT_TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_ptr;
  tc_pdl_sync();
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
      for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
        const int dvt = item % p.dv_tiles;
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
        for (int kt = 0; kt < T1; ++kt) {
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
        load_k(0);
        for (int kt = 0; kt < T; ++kt) {
          if (kt + 1 < T) load_k(kt + 1);
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
    for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
      mbar_wait(&bars->q_full, qph);
      // Pass-1 tiles are 128 keys wide: buffer 0 = the S columns, buffer 1 = the first 128 O columns, which are
      // idle until P.V starts -- once the epilogue of the previous item has read them.
      mbar_wait(&bars->o_empty, oph ^ 1);
      for (int it = 0; it < T1; ++it, ++sn) {
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
      if (T1 > 0) mbar_wait(&bars->s_empty[(sn - 1) & 1], ((sn - 1) >> 1) & 1);
      for (int it = 0; it < T; ++it, ++sn) {
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
            if (p.fast) {
              umma_f16(d, a_h, b_h, idesc_s, k != 0);
            } else {
              umma_f16(d, a_h, b_l, idesc_s, k != 0);
              umma_f16(d, a_l, b_h, idesc_s, 1);
              umma_f16(d, a_h, b_h, idesc_s, 1);
            }
          }
          umma_commit(&bars->s_full[sb]);
          umma_commit(&bars->k_empty[ks]);
          if (it == T - 1) umma_commit(&bars->q_empty);            // Q tile free once the last S has retired
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
    int vs = 0, pb = 0;
    uint32_t vph = 0, pph = 0, oph = 0;
    for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
      mbar_wait(&bars->o_empty, oph ^ 1);                        // epilogue of the previous item has read O
      for (int kt = 0; kt < T; ++kt) {
        mbar_wait(&bars->p_full[pb], pph);
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
              if (p.fast) {
                umma_f16(d, a_h, b_h, idesc_o, (kt | k) != 0);
              } else {
                umma_f16(d, a_h, b_l, idesc_o, (kt | k) != 0);
                umma_f16(d, a_l, b_h, idesc_o, 1);
                umma_f16(d, a_h, b_h, idesc_o, 1);
              }
            }
            umma_commit(&bars->v_empty[vs]);
            if (h == HALVES - 1) {
              umma_commit(&bars->p_empty[pb]);
              if (kt == T - 1) umma_commit(&bars->o_full);
            }
          }
          __syncwarp();
          if (++vs == AT_VSTAGES) { vs = 0; vph ^= 1; }
        }
        if (++pb == 2) { pb = 0; pph ^= 1; }
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
    int pb = 0;
    uint32_t pph = 0, oph = 0;
    uint32_t sn = 0;                                            // S tiles consumed so far (same counting as MMA issuer 1)
    bool out_of_range = false;
    auto group_sync = [] { asm volatile("bar.sync 1, 256;" ::: "memory"); };
    for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
      const int dvt = item % p.dv_tiles;
      int t = item / p.dv_tiles;
      const int qt = p.qt_begin + t % p.q_tiles;
      const int img = t / p.q_tiles;
      const int q_idx = qt * AT_BQ + row;
      const bool valid = q_idx < p.Pq;

      // ---- pass 1: row maximum of S~; 128-key tiles, this group's 64 key columns of each
      float m = -INFINITY;
      for (int kt = 0; kt < T1; ++kt, ++sn) {
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
      // exponent offset of pass 2: the row maximum AND log2 of the 2^10 probability scale, so that one FMA + one
      // MUFU.EX2 yield p * 2^10 directly (the row sum l is then scaled by 2^10 as well: out = O / l)
      const float m_scaled = m * p.scale_log2 - 10.f;
      static_assert(AT_P_SCALE == 1024.f, "the exponent offset above assumes a 2^10 probability scale");

      // ---- pass 2: probabilities -> shared memory (UMMA K-major, 128B swizzle), partial row sum
      float l = 0.f;
      for (int kt = 0; kt < T; ++kt, ++sn) {
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
        mbar_wait(&bars->p_empty[pb], pph ^ 1);
        uint8_t* ph = sP + pb * 2 * AT_P_PLANE + row * 128;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int phys = ((group * 4 + c) ^ (row & 7)) << 4;   // 16-byte chunk inside the 128-byte swizzled row
          *reinterpret_cast<uint4*>(ph + phys) = phv[c];
          *reinterpret_cast<uint4*>(ph + AT_P_PLANE + phys) = plv[c];
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->p_full[pb]);
        if (++pb == 2) { pb = 0; pph ^= 1; }
      }
      bars->xch[group][row] = l;
      group_sync();
      l += bars->xch[group ^ 1][row];
      group_sync();

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
    if (out_of_range && p.range_flag) *reinterpret_cast<volatile int*>(p.range_flag) = 1;   // idempotent store: the flag may live in host-mapped memory
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, AT_TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------
int encode_map_f16(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims,
                   const cuuint64_t* strides_bytes, const cuuint32_t* box, const char* what,
                   const cuuint32_t* elem_strides, int swizzle128 = 1);

// count_only != nullptr: validate, make the same kernel-selection decisions, and report how many kernels the call
// would launch instead of launching them (tdn_attention_tc_launches).
int attention_tc(const tdn_attention_desc* d, cudaStream_t stream, int* count_only) {
  TDN_REQUIRE(d->q_hi && d->q_lo && d->k_hi && d->k_lo && d->vt_hi && d->vt_lo, TDN_ERR_INVALID,
              "attention_tc: null operand");
  TDN_REQUIRE(d->d_k == AT_DK, TDN_ERR_UNSUPPORTED, "attention_tc: d_k must be 64 (got %d)", d->d_k);
  TDN_REQUIRE(d->d_v % AT_DVH == 0, TDN_ERR_UNSUPPORTED, "attention_tc: d_v=%d must be a multiple of 128", d->d_v);
  // 256-wide slices halve the QK^T / softmax recompute; small problems (the FIFO hops with P' queries) would
  // not fill the SMs with them, so they take 128-wide slices = twice as many work items.
  const int num_sms = device_sm_count();
  TDN_REQUIRE(num_sms > 0, TDN_ERR_CUDA, "attention_tc: cannot query the SM count");
  const int num_sms_cached = num_sms;
  const long long items256 = (long long)d->n * ceil_div(d->pq, AT_BQ) * (d->d_v / 256);
  const int dvt_size = (d->d_v % 256 == 0 && items256 >= num_sms_cached) ? 256 : 128;
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
  p.dv_tiles = d->d_v / dvt_size;
  p.k_tiles = ceil_div(d->pk, AT_BK);
  p.k_tiles1 = ceil_div(d->pk, AT_BK1);
  long long items = (long long)d->n * p.q_tiles * p.dv_tiles;
  TDN_REQUIRE(items < (1ll << 31), TDN_ERR_UNSUPPORTED, "attention_tc: too many work items");
  p.num_items = (int)items;
  p.items_a = p.num_items;   // no mixed-width tail unless the split below says so
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
  p.fast = (d->flags & TDN_TC_FLAG_FAST) ? 3 : 0;   // bit 0: single-product S, bit 1: single-product P.V' (the TS kernels test them separately)
  if (const char* dbg = getenv("TDNET_ATTN_DEBUG")) p.fast |= atoi(dbg) & 3;   // probes only: ablation of either product group

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
  // Kernel family: 1 (default) = tc_attn_ts.cu, probabilities handed to the P.V' MMAs through tensor memory;
  // 0 = the kernels of this file (P through shared memory), kept as the yardstick.  TDNET_ATTN_TS selects.
  // (read on every call: the probes and tests flip it inside one process)
  // TDNET_ATTN_TS = 3: the TS kernels with the Q tile in tensor memory as well (tc_attn_ts.cu, "QT").
  // TDNET_ATTN_TS = 4: tc_attn_s128.cu -- the TS kernel with 128-key S MMAs in pass 2 (one S buffer, two P slots).
  const char* ts_env = getenv("TDNET_ATTN_TS");
  const int ts_variant = ts_env ? atoi(ts_env) : ATTN_DEFAULT_VARIANT;
  const bool use_ts = ts_variant != 0;
  {
    static PerDeviceFlag attr_set;                   // the > 48 KB shared-memory opt-in is per device
    const int slot = current_device_slot();
    if (!attr_set.is_set(slot)) {
      TDN_CUDA_OK(cudaFuncSetAttribute(tc_attn_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM_BYTES));
      TDN_CUDA_OK(cudaFuncSetAttribute(tc_attn_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM_BYTES));
      attr_set.set(slot);
    }
  }
  auto launch = [&](int dvt, int grid, const AttnParams& pp) -> cudaError_t {
    if (count_only) { ++*count_only; return cudaSuccess; }
    const bool short_launch = pp.num_items <= 2 * grid;
    if (use_ts) {
      AttnParams q = pp;
      q.per_cta = ceil_div(q.items_a, grid);     // contiguous blocks: the d_v slices of a query tile meet on one CTA
      if (ts_variant == 4) return attention_s128_launch(dvt, grid, stream, short_launch, mq_h, mq_l, mk_h, mk_l, mv_h, mv_l, q);
      return attention_ts_launch(dvt, ts_variant == 3, grid, stream, short_launch, mq_h, mq_l, mk_h, mk_l, mv_h, mv_l, q);
    }
    if (dvt == 256)
      return tc_launch(tc_attn_kernel<256>, grid, AT_THREADS, AT_SMEM_BYTES, stream, short_launch, mq_h, mq_l, mk_h, mk_l, mv_h, mv_l, pp);
    return tc_launch(tc_attn_kernel<128>, grid, AT_THREADS, AT_SMEM_BYTES, stream, short_launch, mq_h, mq_l, mk_h, mk_l, mv_h, mv_l, pp);
  };
  // Wave quantisation: the persistent grid walks the items in rounds of one per SM, and a ragged last round
  // costs a full item time (big hop at 1024x2048: 512 items on 148 SMs = 3.46 rounds -> 4).  When the item count
  // is not a multiple of the SM count, the query tiles of the full rounds run as 256-wide items and the
  // remaining query tiles as 128-wide ones (twice as many items: QK^T and the softmax are recomputed per slice,
  // P.V halves; measured 0.79 of a 256-wide round including the second launch) in a second launch -- 3.79 rounds
  // instead of 4 on the big hop, 0.2915 -> 0.2765 ms on B200, bit-identical output.  Both kernels
  // compute every output element with the same products in the same order, so results do not depend on the
  // split.  TDNET_ATTN_TAIL=0 disables it.
  static int tail_env = -1;
  if (tail_env < 0) {
    const char* e = getenv("TDNET_ATTN_TAIL");
    tail_env = e ? atoi(e) : 1;
  }
  if (dvt_size == 256 && tail_env && p.num_items > num_sms && p.num_items % num_sms != 0) {
    const int per_qt = d->n * p.dv_tiles;                                   // 256-wide items per query tile
    const int q1 = (p.num_items / num_sms) * num_sms / per_qt;              // query tiles of the full rounds
    const int q2 = p.q_tiles - q1;
    const long long items1 = (long long)q1 * per_qt, items2 = (long long)q2 * d->n * (d->d_v / 128);
    const double cost_plain = (double)ceil_div(p.num_items, num_sms);
    const double cost_split = (double)ceil_div(items1, num_sms) + 0.8 * (double)ceil_div(items2, num_sms);
    if (q1 > 0 && q2 > 0 && cost_split < cost_plain) {
      AttnParams p1 = p, p2 = p;
      p1.q_tiles = q1; p1.num_items = (int)items1;
      p2.qt_begin = q1; p2.q_tiles = q2; p2.dv_tiles = d->d_v / 128; p2.num_items = (int)items2;
      if (use_ts) {
        // one launch: every CTA walks its 256-wide items and then (at most) one 128-wide tail item -- no kernel boundary,
        // and the tail item's pass 1 / prologue overlap the previous item's epilogue like any other item
        AttnParams pm = p1;
        pm.items_a = p1.num_items;
        pm.num_items = p1.num_items + p2.num_items;
        pm.dv_tiles_b = p2.dv_tiles; pm.qt_begin_b = p2.qt_begin; pm.q_tiles_b = p2.q_tiles;
        TDN_CUDA_OK(launch(256, num_sms, pm));
        return TDN_OK;
      }
      p1.items_a = p1.num_items; p2.items_a = p2.num_items;
      const int g1 = p1.num_items < num_sms ? p1.num_items : num_sms;
      const int g2 = p2.num_items < num_sms ? p2.num_items : num_sms;
      TDN_CUDA_OK(launch(256, g1, p1));
      TDN_CUDA_OK(launch(128, g2, p2));
      return TDN_OK;
    }
  }
  int grid = p.num_items < num_sms ? p.num_items : num_sms;
  TDN_CUDA_OK(launch(dvt_size, grid, p));
  return TDN_OK;
}

}  // namespace tdn
