// Fused attention-propagation kernel, "S128" variant of tc_attn_ts.cu: the pass-2 S = Q.K^T MMAs cover 128 keys.
//
//   out[q, :] = softmax_k( Q[q,:] . K[k,:] / sqrt(d_k) ) @ V'[k, :]  (+ residual[q, :])
//
// Same operator, work items, two passes and arithmetic (products, probabilities, accumulation order per output
// element and per row sum) as tc_attn_ts.cu / tc_attn.cu -- replaces transformer.py:126-139 and, with fc folded into
// V', Attention.forward :71-92 -- and bit-identical to both.  What differs is the shape of the exact-mode S MMAs of
// pass 2.  In tc_attn_ts.cu they are M128 x N64 x K16: 32 tensor cycles, but 4 KB of Q plus 2 KB of K from shared
// memory = 48 cycles at 128 B/clk (measured 52), so 29 % of a key tile's tensor time runs at 0.6 of the MMA rate
// (624 of 2160 cycles per 64 keys).  With N = 128 the same 4 KB of Q meet 4 KB of K for 64 tensor cycles: operand
// reads and tensor time balance and 64 keys cost 384 cycles.
//
// Tensor memory is full (O: 256 columns, S/P: 256 columns), so the 4 x 64-column in-place S/P ring of tc_attn_ts.cu
// cannot simply be paired up: an S tile of 128 keys would have to wait for TWO P.V' tiles to retire and would queue
// behind the next two in the tensor pipe.  Instead S and P are decoupled:
//   * ONE S buffer of 128 columns.  The softmax warps pull their 2 x 32 columns into registers and hand the buffer
//     back at once (s_free), so S(j+1) is issued ~a tcgen05.ld after S(j) retires and sits in the tensor queue right
//     behind the P.V' MMAs of tile pair j-1 -- where it belongs.
//   * TWO P slots of 64 columns (64 keys: hi pairs and lo pairs per 32-key group, the layout tc_attn_ts.cu writes in
//     place).  The probabilities wait in registers until the slot's previous P.V' MMAs have retired (p_free).
// Tensor-pipe order in steady state: S(j) | PV(2j-2) PV(2j-1) | S(j+1) | PV(2j) PV(2j+1) | ...
//
// Warp roles (512 threads) as in tc_attn_ts.cu: warp 0 TMA producer for Q and K; warp 3 TMA producer for V'^T; warp 1
// issues S (both passes); warp 2 issues O += P.V'^T (A operand from tensor memory); warps 4-11 softmax, two per TMEM
// lane quarter (group g owns key columns [32g, 32g+32) of every 64-key half tile, as before: the row sums are added
// in the same order); warps 12-15 epilogue.
// TMEM (512 columns): O = [0, DVT); S = [256, 384); P slot h = [384 + 64 h, 448 + 64 h).  Pass 1 (row maxima of
// S~ = Qhi.Khi^T, 128-key tiles) uses [256, 384) and [384, 512) as its two tiles.
// Shared memory: Q 32 KB, K ring 3 x 16 KB (one PLANE of 128 keys per slot: a pass-2 tile takes two consecutive slots, hi
// then lo, a pass-1 tile one), V'^T ring 4 x 32 KB (128-row halves of a 64-key tile), epilogue blocks 4 x 4 KB -- the
// footprint of tc_attn_ts.cu.
#include "tc_attn_epilogue.cuh"

namespace tdn {

constexpr int AS_THREADS = 512;
constexpr int AS_PV_WARP = 2;
constexpr int AS_V_WARP = 3;
constexpr int AS_SOFTMAX_WARP0 = 4;      // warps 4-11
constexpr int AS_EPI_WARP0 = 12;         // warps 12-15
constexpr int AS_EPI_WARPS = 4;
constexpr int AS_KSTAGES = 3, AS_VSTAGES = 4;
constexpr int AS_K_STAGE = 2 * AT_K_PLANE;         // one PLANE of 128 keys (two 64-key boxes back to back = one 128-row swizzled tile)
constexpr int AS_SMEM_DATA = 2 * AT_Q_PLANE + AS_KSTAGES * AS_K_STAGE + AS_VSTAGES * 2 * AT_V_PLANE +
                             AS_EPI_WARPS * ATS_EPI_STAGE;

struct AttnS128Bars {
  uint64_t q_full, q_empty;
  uint64_t k_full[AS_KSTAGES], k_empty[AS_KSTAGES];
  uint64_t v_full[AS_VSTAGES], v_empty[AS_VSTAGES];
  uint64_t s1_full[2], s1_empty[2];                // pass 1: S~ tile ready / read by the softmax warps
  uint64_t s_full, s_free;                         // pass 2: S tile (128 keys) ready / pulled into registers
  uint64_t p_full[2], p_free[2];                   // pass 2: P slot written / consumed by P.V'
  uint64_t o_full, o_empty;
  uint64_t l_full, l_empty;                        // row sums of an item written / read by the epilogue warps
  uint32_t tmem_ptr;
  float xch[2][AT_BQ];     // row max exchange between the two softmax warp groups, then their partial row sums
                           // for the epilogue warps (rewritten only after l_empty of the previous item)
};

constexpr int AS_SMEM_BYTES = AS_SMEM_DATA + 1024 /*alignment slack*/ + ((int)sizeof(AttnS128Bars) + 127) / 128 * 128;
static_assert(AS_SMEM_BYTES <= 232448, "attention kernel exceeds the 227 KB shared-memory limit");

template <int DVT>   // d_v slice per work item: 128 or 256 (one or two 128-row V'^T halves per key tile)
__global__ void __launch_bounds__(AS_THREADS, 1)
tc_attn_s128_kernel(const __grid_constant__ CUtensorMap tmQ_hi, const __grid_constant__ CUtensorMap tmQ_lo,
                    const __grid_constant__ CUtensorMap tmK_hi, const __grid_constant__ CUtensorMap tmK_lo,
                    const __grid_constant__ CUtensorMap tmV_hi, const __grid_constant__ CUtensorMap tmV_lo,
                    const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sQ = smem;                                            // hi | lo
  uint8_t* sK = sQ + 2 * AT_Q_PLANE;                             // ring of 128-key planes
  uint8_t* sV = sK + AS_KSTAGES * AS_K_STAGE;                    // stages x (hi | lo)
  uint8_t* sE = sV + AS_VSTAGES * 2 * AT_V_PLANE;                // epilogue turn-around blocks
  AttnS128Bars* bars = reinterpret_cast<AttnS128Bars*>(sE + AS_EPI_WARPS * ATS_EPI_STAGE);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tmQ_hi); prefetch_tensormap(&tmQ_lo);
    prefetch_tensormap(&tmK_hi); prefetch_tensormap(&tmK_lo);
    prefetch_tensormap(&tmV_hi); prefetch_tensormap(&tmV_lo);
    mbar_init(&bars->q_full, 1);
    mbar_init(&bars->q_empty, 1);
    for (int s = 0; s < AS_KSTAGES; ++s) { mbar_init(&bars->k_full[s], 1); mbar_init(&bars->k_empty[s], 1); }
    for (int s = 0; s < AS_VSTAGES; ++s) { mbar_init(&bars->v_full[s], 1); mbar_init(&bars->v_empty[s], 1); }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&bars->s1_full[s], 1);
      mbar_init(&bars->s1_empty[s], AT_SOFTMAX_WARPS);   // one arrival per softmax warp (lane 0 after __syncwarp)
      mbar_init(&bars->p_full[s], AT_SOFTMAX_WARPS);
      mbar_init(&bars->p_free[s], 1);
    }
    mbar_init(&bars->s_full, 1);
    mbar_init(&bars->s_free, AT_SOFTMAX_WARPS);
    mbar_init(&bars->o_full, 1);
    mbar_init(&bars->o_empty, AS_EPI_WARPS);
    mbar_init(&bars->l_full, AT_SOFTMAX_WARPS);
    mbar_init(&bars->l_empty, AS_EPI_WARPS);
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
  const uint32_t tmem_O = tmem_base;
  const uint32_t tmem_S = tmem_base + 256;       // pass 2: S tile; pass 1: tile 0
  const uint32_t tmem_P = tmem_base + 384;       // pass 2: P slot h at + 64 h; pass 1: tile 1
  const int T = p.k_tiles;                       // 64-key tiles (P.V' granularity)
  const int T1 = p.k_tiles1;                     // 128-key tiles (S granularity, both passes)

  if (warp == 0) {
    // ================================ TMA producer: Q tile and key tiles ================================
    if (lane == 0) {
      int ks = 0;
      uint32_t kph = 0, qph = 0;
      for (int k = 0, item, prev = -1; (item = attn_walk(p, k)) >= 0; prev = item, ++k) {
        const AttnItem w = attn_item<DVT>(p, item);
        const int qt = w.qt, img = w.img;
        mbar_wait(&bars->q_empty, qph ^ 1);
        mbar_expect_tx(&bars->q_full, 2 * AT_Q_PLANE);
        tma_load_3d(sQ, &tmQ_hi, &bars->q_full, 0, qt * AT_BQ, img);
        tma_load_3d(sQ + AT_Q_PLANE, &tmQ_lo, &bars->q_full, 0, qt * AT_BQ, img);
        qph ^= 1;
        // pass 1 (skipped when this CTA has just computed the row maxima of the same query tile): the hi plane of the
        // keys only (S~ = Qhi.Khi^T); two 64-key boxes land back to back = one 128-row swizzled tile (a box past the
        // last key is zero-filled)
        const int t1 = attn_shares_rowmax(p, item, prev) ? 0 : T1;
        for (int kt = 0; kt < t1; ++kt) {
          mbar_wait(&bars->k_empty[ks], kph ^ 1);
          uint8_t* dst = sK + ks * AS_K_STAGE;
          mbar_expect_tx(&bars->k_full[ks], 2 * AT_K_PLANE);
          tma_load_3d(dst, &tmK_hi, &bars->k_full[ks], 0, kt * AT_BK1, img);
          tma_load_3d(dst + AT_K_PLANE, &tmK_hi, &bars->k_full[ks], 0, kt * AT_BK1 + AT_BK, img);
          if (++ks == AS_KSTAGES) { ks = 0; kph ^= 1; }
        }
        // pass 2: the hi plane and the lo plane of 128 keys go to consecutive ring slots
        for (int kt = 0; kt < 2 * T1; ++kt) {
          const CUtensorMap* tm = (kt & 1) ? &tmK_lo : &tmK_hi;
          const int k0 = (kt >> 1) * AT_BK1;
          mbar_wait(&bars->k_empty[ks], kph ^ 1);
          uint8_t* dk = sK + ks * AS_K_STAGE;
          mbar_expect_tx(&bars->k_full[ks], 2 * AT_K_PLANE);
          tma_load_3d(dk, tm, &bars->k_full[ks], 0, k0, img);
          tma_load_3d(dk + AT_K_PLANE, tm, &bars->k_full[ks], 0, k0 + AT_BK, img);
          if (++ks == AS_KSTAGES) { ks = 0; kph ^= 1; }
        }
      }
    }
  } else if (warp == AS_V_WARP) {
    // ================================ TMA producer: V'^T slice ================================
    if (lane == 0) {
      int vs = 0;
      uint32_t vph = 0;
      for (int k = 0, item; (item = attn_walk(p, k)) >= 0; ++k) {
        const AttnItem w = attn_item<DVT>(p, item);
        const int img = w.img;
        for (int kt = 0; kt < T; ++kt) {
          for (int h = 0; h < w.halves; ++h) {
            mbar_wait(&bars->v_empty[vs], vph ^ 1);
            uint8_t* dv = sV + vs * 2 * AT_V_PLANE;
            mbar_expect_tx(&bars->v_full[vs], 2 * AT_V_PLANE);
            tma_load_3d(dv, &tmV_hi, &bars->v_full[vs], kt * AT_BK, w.dv0 + h * AT_DVH, img);
            tma_load_3d(dv + AT_V_PLANE, &tmV_lo, &bars->v_full[vs], kt * AT_BK, w.dv0 + h * AT_DVH, img);
            if (++vs == AS_VSTAGES) { vs = 0; vph ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer 1: S = Q.K^T (both passes) ================================
    // The whole warp runs the loop and the barrier waits so that stage indices, phases and descriptors
    // stay warp-uniform (uniform registers feed tcgen05.mma directly); one elected lane issues.
    constexpr uint32_t idesc_s = umma_idesc_f16(AT_BQ, AT_BK1);   // 128 x 128 (both passes)
    int ks = 0;
    uint32_t kph = 0, qph = 0;
    uint32_t n1 = 0;      // pass-1 tiles issued so far: tile buffer n1 & 1, use n1 >> 1
    uint32_t ns = 0;      // pass-2 S tiles issued so far
    uint32_t items_done = 0;
    const uint32_t q_hi = smem_u32(sQ), q_lo = q_hi + AT_Q_PLANE;
    for (int k = 0, item, prev = -1; (item = attn_walk(p, k)) >= 0; prev = item, ++k, ++items_done) {
      const bool reuse = attn_shares_rowmax(p, item, prev);   // the row maxima of this query tile are known already
      mbar_wait(&bars->q_full, qph);
      // pass 1 writes the S buffer and both P slots, which hold probabilities of the previous item until its last P.V'
      // MMA has retired; with the row maxima reused there is no pass 1 and the s_free / p_free waits of pass 2 suffice
      const int t1 = reuse ? 0 : T1;
      // (waited for on EVERY item, also when there is no pass 1: a parity wait that skips a phase passes at once when the
      //  barrier is still two phases back -- with few key tiles per item this warp gets that far ahead -- and P.V' cannot
      //  restart before the epilogue has drained O anyway, so the wait costs nothing)
      if (items_done > 0) mbar_wait(&bars->o_full, (items_done - 1) & 1);
      for (int it = 0; it < t1; ++it, ++n1) {
        const int pair = n1 & 1;
        mbar_wait(&bars->k_full[ks], kph);
        mbar_wait(&bars->s1_empty[pair], ((n1 >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t k_hi = smem_u32(sK + ks * AS_K_STAGE);
        const uint32_t d = tmem_S + pair * AT_BK1;
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < AT_DK / 16; ++k)
            umma_f16(d, umma_desc_k_sw128(q_hi + k * 32), umma_desc_k_sw128(k_hi + k * 32), idesc_s, k != 0);
          umma_commit(&bars->s1_full[pair]);
          umma_commit(&bars->k_empty[ks]);
        }
        __syncwarp();
        if (++ks == AS_KSTAGES) { ks = 0; kph ^= 1; }
      }
      // pass 2 overwrites the pass-1 tiles: the softmax warps must have read the last two of them
      for (uint32_t j = 1; j <= 2 && j <= (uint32_t)t1; ++j) {
        const uint32_t t = n1 - j;
        mbar_wait(&bars->s1_empty[t & 1], (t >> 1) & 1);
      }
      for (int it = 0; it < T1; ++it, ++ns) {
        // two ring slots per tile: hi plane, then lo plane
        const int ks_h = ks;
        const uint32_t kph_h = kph;
        if (++ks == AS_KSTAGES) { ks = 0; kph ^= 1; }
        const int ks_l = ks;
        const uint32_t kph_l = kph;
        if (++ks == AS_KSTAGES) { ks = 0; kph ^= 1; }
        mbar_wait(&bars->k_full[ks_h], kph_h);
        mbar_wait(&bars->k_full[ks_l], kph_l);
        mbar_wait(&bars->s_free, (ns & 1) ^ 1);
        tc_fence_after();
        const uint32_t k_hi = smem_u32(sK + ks_h * AS_K_STAGE), k_lo = smem_u32(sK + ks_l * AS_K_STAGE);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < AT_DK / 16; ++k) {
            const uint64_t a_h = umma_desc_k_sw128(q_hi + k * 32), b_h = umma_desc_k_sw128(k_hi + k * 32);
            const uint64_t a_l = umma_desc_k_sw128(q_lo + k * 32), b_l = umma_desc_k_sw128(k_lo + k * 32);
            if (p.fast & 1) {
              umma_f16(tmem_S, a_h, b_h, idesc_s, k != 0);
            } else {
              umma_f16(tmem_S, a_h, b_l, idesc_s, k != 0);
              umma_f16(tmem_S, a_l, b_h, idesc_s, 1);
              umma_f16(tmem_S, a_h, b_h, idesc_s, 1);
            }
          }
          umma_commit(&bars->s_full);
          umma_commit(&bars->k_empty[ks_h]);
          umma_commit(&bars->k_empty[ks_l]);
          if (it == T1 - 1) umma_commit(&bars->q_empty);           // Q tile free once the last S has retired
        }
        __syncwarp();
      }
      qph ^= 1;
    }
  } else if (warp == AS_PV_WARP) {
    // ================================ MMA issuer 2: O += P.V'^T, P read from tensor memory ================================
    constexpr uint32_t idesc_o = umma_idesc_f16(AT_BQ, AT_DVH);  // 128 x 128
    int vs = 0;
    uint32_t vph = 0, oph = 0;
    uint32_t np0 = 0, np1 = 0;                                   // uses of each P slot so far
    for (int k = 0, item; (item = attn_walk(p, k)) >= 0; ++k) {
      const int halves = attn_item<DVT>(p, item).halves;
      mbar_wait(&bars->o_empty, oph ^ 1);                        // epilogue of the previous item has read O
      for (int kt = 0; kt < T; ++kt) {
        const int slot = kt & 1;
        mbar_wait(&bars->p_full[slot], (slot ? np1 : np0) & 1);
        if (slot) ++np1; else ++np0;
        const uint32_t p_base = tmem_P + slot * AT_BK;
        for (int h = 0; h < halves; ++h) {
          mbar_wait(&bars->v_full[vs], vph);
          tc_fence_after();
          const uint32_t v_hi = smem_u32(sV + vs * 2 * AT_V_PLANE), v_lo = v_hi + AT_V_PLANE;
          const uint32_t d = tmem_O + h * AT_DVH;
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < AT_BK / 16; ++k) {
              // keys [16k, 16k+16) were written by softmax group k >> 1: hi pairs at columns 32g + 8 (k & 1),
              // lo pairs 16 columns further
              const uint32_t a_h = p_base + (k >> 1) * 32 + (k & 1) * 8, a_l = a_h + 16;
              const uint64_t b_h = umma_desc_k_sw128(v_hi + k * 32), b_l = umma_desc_k_sw128(v_lo + k * 32);
              if (p.fast & 2) {
                umma_f16_ts(d, a_h, b_h, idesc_o, (kt | k) != 0);
              } else {
                umma_f16_ts(d, a_h, b_l, idesc_o, (kt | k) != 0);
                umma_f16_ts(d, a_l, b_h, idesc_o, 1);
                umma_f16_ts(d, a_h, b_h, idesc_o, 1);
              }
            }
            umma_commit(&bars->v_empty[vs]);
            if (h == halves - 1) {
              umma_commit(&bars->p_free[slot]);
              if (kt == T - 1) umma_commit(&bars->o_full);
            }
          }
          __syncwarp();
          if (++vs == AS_VSTAGES) { vs = 0; vph ^= 1; }
        }
      }
      oph ^= 1;
    }
  } else if (warp < AS_EPI_WARP0) {
    // ================================ softmax warps ================================
    const int quarter = warp & 3;
    const int group = (warp - AS_SOFTMAX_WARP0) >> 2;
    const int row = quarter * 32 + lane;                      // query row inside the tile = TMEM lane
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    uint32_t n1 = 0, ns = 0;                                  // same counting as MMA issuer 1
    uint32_t np[2] = {0, 0};                                  // same counting as MMA issuer 2
    uint32_t items_done = 0;
    auto group_sync = [] { asm volatile("bar.sync 1, 256;" ::: "memory"); };
    float m = -INFINITY;                                      // row maximum; survives to the next item of the same query tile
    for (int k = 0, item, prev = -1; (item = attn_walk(p, k)) >= 0; prev = item, ++k, ++items_done) {
      // ---- pass 1: row maximum of S~; 128-key tiles, this group's 64 key columns of each
      const bool reuse = attn_shares_rowmax(p, item, prev);
      if (!reuse) {
        m = -INFINITY;
        for (int kt = 0; kt < T1; ++kt, ++n1) {
          const int pair = n1 & 1;
          mbar_wait(&bars->s1_full[pair], (n1 >> 1) & 1);
          tc_fence_after();
          uint32_t r0[32], r1[32];
          const uint32_t src = tmem_S + pair * AT_BK1 + lane_addr + group * 64;
          tmem_ld_32x32(src, r0);
          tmem_ld_32x32(src + 32, r1);
          tmem_ld_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bars->s1_empty[pair]);      // the tile is in registers: hand the buffer back first
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
        }
        mbar_wait(&bars->l_empty, (items_done & 1) ^ 1);         // the row sums of the previous item have been read
        bars->xch[group][row] = m;
        group_sync();
        m = fmaxf(m, bars->xch[group ^ 1][row]);
        group_sync();                                           // xch carries the row sums next
      } else {
        mbar_wait(&bars->l_empty, (items_done & 1) ^ 1);         // (the row sums below reuse xch)
      }
      // exponent offset of pass 2: the row maximum AND log2 of the 2^10 probability scale, so that one FMA + one
      // MUFU.EX2 yield p * 2^10 directly (the row sum l is then scaled by 2^10 as well: out = O / l)
      const float m_scaled = m * p.scale_log2 - 10.f;
      static_assert(AT_P_SCALE == 1024.f, "the exponent offset above assumes a 2^10 probability scale");

      // ---- pass 2: S (128 keys) -> registers -> probabilities of the two 64-key halves -> P slots; partial row sum
      float l = 0.f;
      for (int kt = 0; kt < T1; ++kt, ++ns) {
        mbar_wait(&bars->s_full, ns & 1);
        tc_fence_after();
        uint32_t r[2][32];
        tmem_ld_32x32(tmem_S + lane_addr + group * 32, r[0]);
        tmem_ld_32x32(tmem_S + lane_addr + AT_BK + group * 32, r[1]);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->s_free);               // S(kt+1) may be issued while this tile is processed
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          if (2 * kt + h >= T) break;                            // odd number of 64-key tiles: the last half b does not exist
          const int kbase = (2 * kt + h) * AT_BK + group * 32;
          float pr[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) pr[j] = fast_exp2(fmaf(__uint_as_float(r[h][j]), p.scale_log2, -m_scaled));
          if (kbase + 32 > p.Pk) {                               // ragged last tile: keys past P' contribute nothing
#pragma unroll
            for (int j = 0; j < 32; ++j) pr[j] = (kbase + j < p.Pk) ? pr[j] : 0.f;
          }
#pragma unroll
          for (int j = 0; j < 32; ++j) l += pr[j];
          uint32_t ph[16], pl[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            __half2 hi, lo;
            split_f32x2(pr[2 * e], pr[2 * e + 1], hi, lo);       // key 2e in the low half, key 2e+1 in the high half
            ph[e] = *reinterpret_cast<const uint32_t*>(&hi);
            pl[e] = *reinterpret_cast<const uint32_t*>(&lo);
          }
          mbar_wait(&bars->p_free[h], (np[h] & 1) ^ 1);          // the slot's previous P.V' MMAs have retired
          ++np[h];
          tc_fence_after();
          const uint32_t taddr = tmem_P + h * AT_BK + lane_addr + group * 32;
          tmem_st_32x16(taddr, ph);
          tmem_st_32x16(taddr + 16, pl);
          tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bars->p_full[h]);
        }
      }
      // ---- hand the partial row sum to the epilogue warps (they add the two groups in a fixed order)
      bars->xch[group][row] = l;
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->l_full);               // release: orders the stores above
    }
  } else {
    // ================================ epilogue warps ================================
    // out = O / l + residual for the 32 query rows of this warp's TMEM lane quarter, all DVT channels
    // (attn_epilogue_item in tc_attn_epilogue.cuh).
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t tmem_o = tmem_O + ((uint32_t)(quarter * 32) << 16);
    const uint32_t stg = smem_u32(sE + (warp - AS_EPI_WARP0) * ATS_EPI_STAGE);
    const int fmt = attn_epilogue_fmt(p);
    uint32_t iph = 0;
    bool out_of_range = false;
    for (int k = 0, item; (item = attn_walk(p, k)) >= 0; ++k) {
      const AttnItem w = attn_item<DVT>(p, item);
      const int q0 = w.qt * AT_BQ + quarter * 32;              // first query row of this warp
      const int NCHUNK = w.halves * (AT_DVH / 32);
      const long long obase = (long long)w.img * p.o_bs + w.dv0;
      const long long rbase = (long long)w.img * p.r_bs + w.dv0;
      mbar_wait(&bars->l_full, iph);
      const float l = bars->xch[0][row] + bars->xch[1][row];
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->l_empty);
      const float inv = 1.f / l;                               // l carries the 2^10 scale of P
      mbar_wait(&bars->o_full, iph);
      tc_fence_after();
      iph ^= 1;
      attn_epilogue_dispatch(fmt, p, tmem_o, stg, lane, q0, rbase, obase, inv, 0, NCHUNK, out_of_range);
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

cudaError_t attention_s128_launch(int dvt, int grid, cudaStream_t stream, bool short_launch, const CUtensorMap& mq_h,
                                  const CUtensorMap& mq_l, const CUtensorMap& mk_h, const CUtensorMap& mk_l,
                                  const CUtensorMap& mv_h, const CUtensorMap& mv_l, const AttnParams& p) {
  // the > 48 KB dynamic shared-memory opt-in is a per-device function attribute: set it once per device
  static PerDeviceFlag attr_set;
  const int slot = current_device_slot();
  if (!attr_set.is_set(slot)) {
    cudaError_t e;
    if ((e = cudaFuncSetAttribute(tc_attn_s128_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, AS_SMEM_BYTES)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(tc_attn_s128_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, AS_SMEM_BYTES)) != cudaSuccess) return e;
    attr_set.set(slot);
  }
  if (dvt == 256)
    return tc_launch(tc_attn_s128_kernel<256>, grid, AS_THREADS, AS_SMEM_BYTES, stream, short_launch, mq_h, mq_l, mk_h, mk_l, mv_h, mv_l, p);
  return tc_launch(tc_attn_s128_kernel<128>, grid, AS_THREADS, AS_SMEM_BYTES, stream, short_launch, mq_h, mq_l, mk_h, mk_l, mv_h, mv_l, p);
}

}  // namespace tdn
