// Fused attention-propagation kernel, tensor-memory operand variant ("TS": A from TMEM, B from shared memory).
//
//   out[q, :] = softmax_k( Q[q,:] . K[k,:] / sqrt(d_k) ) @ V'[k, :]  (+ residual[q, :])
//
// Same operator, same work items, same two passes and the same arithmetic (products, probabilities, accumulation
// order per output element) as tc_attn.cu -- replaces transformer.py:126-139 and, with fc folded into V',
// Attention.forward :71-92 -- but the probabilities never touch shared memory:
//
//   * S = Q.K^T lands in one of FOUR 64-column TMEM buffers; the softmax warps read their 32 columns
//     (tcgen05.ld), compute p * 2^10 = 2^(S.c - m + 10), split it to fp16 hi / lo and write the packed pairs back
//     INTO THE SAME COLUMNS (tcgen05.st: 32 fp32 columns = 16 columns of hi pairs + 16 columns of lo pairs);
//   * O += P.V'^T is issued with the A operand in tensor memory (tcgen05.mma [d], [a_tmem], b_desc): a K16 step
//     reads 8 columns of a P plane.  The buffer returns to the S issuer when those MMAs retire.
//
// Why: with P in shared memory every P.V' MMA (M128 x N128 x K16, 64 tensor cycles) read 4 KB of A and 4 KB of B
// from shared memory = the full 128 B/clk of the SM, on top of the TMA writes and the P stores (ncu of tc_attn.cu:
// tensor pipe 51 % busy, 5.4 M shared-memory wavefronts per launch).  Here the MMAs read B only (64 B/clk), the
// 2 x 32 KB P double buffer becomes a fourth V'^T stage and a third K stage, and the generic->async proxy fence per
// tile is gone.  With four S/P buffers the S issuer runs up to three key tiles ahead of P.V'.
//
// Warp roles (512 threads): warp 0 TMA producer for Q and K; warp 3 TMA producer for V'^T (a full V ring must not
// hold back the key tiles S needs); warp 1 issues S = Q.K^T (both passes); warp 2 issues O += P.V'^T;
// warps 4-11 softmax, two per TMEM lane quarter (group g owns key columns [32g, 32g+32) of a tile);
// warps 12-15 epilogue, one per lane quarter (out = O / l + residual for all DVT channels of 32 query rows).
// The epilogue of item i runs NEXT TO pass 1 and the first softmax tiles of item i+1: ncu of the previous layout
// (the softmax warps also ran the epilogue) showed 22 % of an item in the epilogue and 13 % in pass 1 with the P.V'
// pipe idle in both; the row sums reach the epilogue warps through shared memory.
// TMEM (512 columns): O = columns [0, DVT); S/P buffer b = columns [256 + 64 b, 256 + 64 b + 64).  Pass 1 (row maxima of
// S~ = Qhi.Khi^T, 128-key tiles) uses the buffers pairwise as two 128-column tiles.
// Shared memory: Q 32 KB, K ring 3 x 16 KB, V'^T ring 4 x 32 KB (128-row halves of a key tile).
#include "tc_attn_epilogue.cuh"

namespace tdn {

constexpr int ATS_THREADS = 512;
constexpr int ATS_PV_WARP = 2;
constexpr int ATS_V_WARP = 3;
constexpr int ATS_SOFTMAX_WARP0 = 4;      // warps 4-11
constexpr int ATS_EPI_WARP0 = 12;         // warps 12-15
constexpr int ATS_EPI_WARPS = 4;
constexpr int ATS_KSTAGES = 3, ATS_VSTAGES = 4;
constexpr int ATS_SP = 4;                          // S/P buffers of 64 TMEM columns (three in the QT variant)
constexpr int ATS_QT_COL = 448;                    // QT variant: Q hi pairs in TMEM columns [448, 480), lo pairs [480, 512)
constexpr int ATS_SMEM_DATA = 2 * AT_Q_PLANE + ATS_KSTAGES * 2 * AT_K_PLANE + ATS_VSTAGES * 2 * AT_V_PLANE +
                              ATS_EPI_WARPS * ATS_EPI_STAGE;

struct AttnTsBars {
  uint64_t q_full, q_empty;
  uint64_t k_full[ATS_KSTAGES], k_empty[ATS_KSTAGES];
  uint64_t v_full[ATS_VSTAGES], v_empty[ATS_VSTAGES];
  uint64_t s1_full[2], s1_empty[2];                // pass 1: S~ tile ready / read by the softmax warps
  uint64_t s_full[ATS_SP], p_full[ATS_SP], sp_empty[ATS_SP];   // pass 2: S ready / P written / P consumed by P.V'
  uint64_t o_full, o_empty;
  uint64_t l_full, l_empty;                        // row sums of an item written / read by the epilogue warps
  uint64_t qt_full;                                // QT variant: the Q tile has been copied into tensor memory
  uint32_t tmem_ptr;
  float xch[2][AT_BQ];     // row max exchange between the two softmax warp groups, then their partial row sums
                           // for the epilogue warps (rewritten only after l_empty of the previous item)
};

constexpr int ATS_SMEM_BYTES = ATS_SMEM_DATA + 1024 /*alignment slack*/ + ((int)sizeof(AttnTsBars) + 127) / 128 * 128;
static_assert(ATS_SMEM_BYTES <= 232448, "attention kernel exceeds the 227 KB shared-memory limit");


// QT ("Q in tensor memory"): the pass-2 S MMAs are M128 x N64 x K16 -- 32 tensor cycles, but 4 KB of A plus 2 KB of B from
// shared memory at 128 B/clk = 48 cycles (measured 52), i.e. 29 % of a key tile's tensor time is spent at 0.6 of the MMA
// rate.  Q is the A operand of every one of them and is constant for the whole item, so the softmax warps copy the tile
// (hi and lo planes, 64 + 64 fp16 per row = 64 TMEM columns) into tensor memory once per query tile, after pass 1, and
// the S MMAs take A from there (B only from shared memory: 16 cycles).  The columns come from the S/P ring, which
// shrinks from four to three buffers; pass 1 still uses [256, 512) as two 128-column tiles (the copy follows it).
template <int DVT, bool QT>   // d_v slice per work item: 128 or 256 (one or two 128-row V'^T halves per key tile)
__global__ void __launch_bounds__(ATS_THREADS, 1)
tc_attn_ts_kernel(const __grid_constant__ CUtensorMap tmQ_hi, const __grid_constant__ CUtensorMap tmQ_lo,
                  const __grid_constant__ CUtensorMap tmK_hi, const __grid_constant__ CUtensorMap tmK_lo,
                  const __grid_constant__ CUtensorMap tmV_hi, const __grid_constant__ CUtensorMap tmV_lo,
                  const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sQ = smem;                                            // hi | lo
  uint8_t* sK = sQ + 2 * AT_Q_PLANE;                             // stages x (hi | lo)
  uint8_t* sV = sK + ATS_KSTAGES * 2 * AT_K_PLANE;               // stages x (hi | lo)
  uint8_t* sE = sV + ATS_VSTAGES * 2 * AT_V_PLANE;               // epilogue turn-around blocks
  AttnTsBars* bars = reinterpret_cast<AttnTsBars*>(sE + ATS_EPI_WARPS * ATS_EPI_STAGE);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tmQ_hi); prefetch_tensormap(&tmQ_lo);
    prefetch_tensormap(&tmK_hi); prefetch_tensormap(&tmK_lo);
    prefetch_tensormap(&tmV_hi); prefetch_tensormap(&tmV_lo);
    mbar_init(&bars->q_full, 1);
    mbar_init(&bars->q_empty, 1);
    for (int s = 0; s < ATS_KSTAGES; ++s) { mbar_init(&bars->k_full[s], 1); mbar_init(&bars->k_empty[s], 1); }
    for (int s = 0; s < ATS_VSTAGES; ++s) { mbar_init(&bars->v_full[s], 1); mbar_init(&bars->v_empty[s], 1); }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&bars->s1_full[s], 1);
      mbar_init(&bars->s1_empty[s], AT_SOFTMAX_WARPS);   // one arrival per softmax warp (lane 0 after __syncwarp)
    }
    for (int s = 0; s < ATS_SP; ++s) {
      mbar_init(&bars->s_full[s], 1);
      mbar_init(&bars->p_full[s], AT_SOFTMAX_WARPS);
      mbar_init(&bars->sp_empty[s], 1);
    }
    mbar_init(&bars->o_full, 1);
    mbar_init(&bars->o_empty, ATS_EPI_WARPS);
    mbar_init(&bars->l_full, AT_SOFTMAX_WARPS);
    mbar_init(&bars->l_empty, ATS_EPI_WARPS);
    mbar_init(&bars->qt_full, AT_SOFTMAX_WARPS);
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
  const uint32_t tmem_SP = tmem_base + 256;      // + b * 64
  constexpr uint32_t SP = QT ? 3 : ATS_SP;       // S/P ring depth
  const int T = p.k_tiles;
  const int T1 = p.k_tiles1;

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
        // keys only (S~ = Qhi.Khi^T), 128 keys per stage: two 64-key boxes land back to back = one 128-row swizzled tile
        // (a box past the last key is zero-filled)
        const int t1 = attn_shares_rowmax(p, item, prev) ? 0 : T1;
        for (int kt = 0; kt < t1; ++kt) {
          mbar_wait(&bars->k_empty[ks], kph ^ 1);
          uint8_t* dst = sK + ks * 2 * AT_K_PLANE;
          mbar_expect_tx(&bars->k_full[ks], 2 * AT_K_PLANE);
          tma_load_3d(dst, &tmK_hi, &bars->k_full[ks], 0, kt * AT_BK1, img);
          tma_load_3d(dst + AT_K_PLANE, &tmK_hi, &bars->k_full[ks], 0, kt * AT_BK1 + AT_BK, img);
          if (++ks == ATS_KSTAGES) { ks = 0; kph ^= 1; }
        }
        // pass 2: keys hi + lo, 64 per stage
        for (int kt = 0; kt < T; ++kt) {
          mbar_wait(&bars->k_empty[ks], kph ^ 1);
          uint8_t* dk = sK + ks * 2 * AT_K_PLANE;
          mbar_expect_tx(&bars->k_full[ks], 2 * AT_K_PLANE);
          tma_load_3d(dk, &tmK_hi, &bars->k_full[ks], 0, kt * AT_BK, img);
          tma_load_3d(dk + AT_K_PLANE, &tmK_lo, &bars->k_full[ks], 0, kt * AT_BK, img);
          if (++ks == ATS_KSTAGES) { ks = 0; kph ^= 1; }
        }
      }
    }
  } else if (warp == ATS_V_WARP) {
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
            if (++vs == ATS_VSTAGES) { vs = 0; vph ^= 1; }
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
    uint32_t kph = 0, qph = 0;
    uint32_t n1 = 0;      // pass-1 tiles issued so far: buffer pair n1 & 1, use n1 >> 1
    uint32_t n2 = 0;      // pass-2 tiles issued so far: buffer n2 % SP, use n2 / SP
    uint32_t nq = 0;      // QT: query tiles copied into tensor memory so far
    uint32_t items_done = 0;
    const uint32_t q_hi = smem_u32(sQ), q_lo = q_hi + AT_Q_PLANE;
    for (int k = 0, item, prev = -1; (item = attn_walk(p, k)) >= 0; prev = item, ++k, ++items_done) {
      const bool reuse = attn_shares_rowmax(p, item, prev);   // the row maxima of this query tile are known already
      mbar_wait(&bars->q_full, qph);
      // pass 1 writes whole S/P buffer pairs, which still hold probabilities of the previous item until its last P.V'
      // MMA has retired; with the row maxima reused there is no pass 1 and the per-buffer sp_empty waits of pass 2 suffice
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
        const uint32_t k_hi = smem_u32(sK + ks * 2 * AT_K_PLANE);
        const uint32_t d = tmem_SP + pair * AT_BK1;
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < AT_DK / 16; ++k)
            umma_f16(d, umma_desc_k_sw128(q_hi + k * 32), umma_desc_k_sw128(k_hi + k * 32), idesc_s1, k != 0);
          umma_commit(&bars->s1_full[pair]);
          umma_commit(&bars->k_empty[ks]);
        }
        __syncwarp();
        if (++ks == ATS_KSTAGES) { ks = 0; kph ^= 1; }
      }
      // pass 2 overwrites the pass-1 tiles: the softmax warps must have read the last two of them
      for (uint32_t j = 1; j <= 2 && j <= (uint32_t)t1; ++j) {
        const uint32_t t = n1 - j;
        mbar_wait(&bars->s1_empty[t & 1], (t >> 1) & 1);
      }
      if (QT && !reuse) {                                           // the softmax warps have put this Q tile into TMEM
        mbar_wait(&bars->qt_full, nq & 1);
        ++nq;
      }
      for (int it = 0; it < T; ++it, ++n2) {
        const int sb = n2 % SP;
        mbar_wait(&bars->k_full[ks], kph);
        mbar_wait(&bars->sp_empty[sb], ((n2 / SP) & 1) ^ 1);
        tc_fence_after();
        const uint32_t k_hi = smem_u32(sK + ks * 2 * AT_K_PLANE), k_lo = k_hi + AT_K_PLANE;
        const uint32_t d = tmem_SP + sb * AT_BK;
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < AT_DK / 16; ++k) {
            const uint64_t a_h = umma_desc_k_sw128(q_hi + k * 32), b_h = umma_desc_k_sw128(k_hi + k * 32);
            const uint64_t a_l = umma_desc_k_sw128(q_lo + k * 32), b_l = umma_desc_k_sw128(k_lo + k * 32);
            if (QT) {
              const uint32_t t_h = tmem_base + ATS_QT_COL + k * 8, t_l = t_h + 32;   // 8 columns of fp16 pairs per K16 step
              if (p.fast & 1) {
                umma_f16_ts(d, t_h, b_h, idesc_s, k != 0);
              } else {
                umma_f16_ts(d, t_h, b_l, idesc_s, k != 0);
                umma_f16_ts(d, t_l, b_h, idesc_s, 1);
                umma_f16_ts(d, t_h, b_h, idesc_s, 1);
              }
            } else if (p.fast & 1) {
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
        if (++ks == ATS_KSTAGES) { ks = 0; kph ^= 1; }
      }
      qph ^= 1;
    }
  } else if (warp == ATS_PV_WARP) {
    // ================================ MMA issuer 2: O += P.V'^T, P read from tensor memory ================================
    constexpr uint32_t idesc_o = umma_idesc_f16(AT_BQ, AT_DVH);  // 128 x 128
    int vs = 0;
    uint32_t vph = 0, oph = 0;
    uint32_t n2 = 0;
    for (int k = 0, item; (item = attn_walk(p, k)) >= 0; ++k) {
      const int halves = attn_item<DVT>(p, item).halves;
      mbar_wait(&bars->o_empty, oph ^ 1);                        // epilogue of the previous item has read O
      for (int kt = 0; kt < T; ++kt, ++n2) {
        const int sb = n2 % SP;
        mbar_wait(&bars->p_full[sb], (n2 / SP) & 1);
        const uint32_t p_base = tmem_SP + sb * AT_BK;
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
              umma_commit(&bars->sp_empty[sb]);
              if (kt == T - 1) umma_commit(&bars->o_full);
            }
          }
          __syncwarp();
          if (++vs == ATS_VSTAGES) { vs = 0; vph ^= 1; }
        }
      }
      oph ^= 1;
    }
  } else if (warp < ATS_EPI_WARP0) {
    // ================================ softmax warps ================================
    const int quarter = warp & 3;
    const int group = (warp - ATS_SOFTMAX_WARP0) >> 2;
    const int row = quarter * 32 + lane;                      // query row inside the tile = TMEM lane
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    uint32_t n1 = 0, n2 = 0;                                  // same counting as MMA issuer 1
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
        const uint32_t src = tmem_SP + pair * AT_BK1 + lane_addr + group * 64;
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
      if (QT) {
        // Q tile -> tensor memory (group 0: hi plane, group 1: lo plane; one row per thread).  Every softmax warp is past
        // its last pass-1 tile here (the syncs above), so the columns [448, 512) of the second pass-1 buffer are free.
        mbar_wait(&bars->q_full, items_done & 1);             // TMA-written shared memory: observe the barrier ourselves
        const uint32_t qrow = smem_u32(sQ) + group * AT_Q_PLANE + row * 128;
        uint32_t w0[16], w1[16];
#pragma unroll
        for (int c = 0; c < 4; ++c) {                           // 128-byte swizzle: 16-byte piece c of row r sits at c ^ (r & 7)
          const uint4 a = lds128(qrow + ((c ^ (row & 7)) << 4));
          const uint4 b2 = lds128(qrow + (((c + 4) ^ (row & 7)) << 4));
          w0[4 * c] = a.x; w0[4 * c + 1] = a.y; w0[4 * c + 2] = a.z; w0[4 * c + 3] = a.w;
          w1[4 * c] = b2.x; w1[4 * c + 1] = b2.y; w1[4 * c + 2] = b2.z; w1[4 * c + 3] = b2.w;
        }
        const uint32_t qdst = tmem_base + ATS_QT_COL + group * 32 + lane_addr;
        tmem_st_32x16(qdst, w0);
        tmem_st_32x16(qdst + 16, w1);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->qt_full);
      }
      } else {
        mbar_wait(&bars->l_empty, (items_done & 1) ^ 1);       // (the row sums below reuse xch)
      }
      // exponent offset of pass 2: the row maximum AND log2 of the 2^10 probability scale, so that one FMA + one
      // MUFU.EX2 yield p * 2^10 directly (the row sum l is then scaled by 2^10 as well: out = O / l)
      const float m_scaled = m * p.scale_log2 - 10.f;
      static_assert(AT_P_SCALE == 1024.f, "the exponent offset above assumes a 2^10 probability scale");

      // ---- pass 2: S -> probabilities, written back over S as packed fp16 hi / lo pairs; partial row sum
      float l = 0.f;
      for (int kt = 0; kt < T; ++kt, ++n2) {
        const int sb = n2 % SP;
        const uint32_t taddr = tmem_SP + sb * AT_BK + lane_addr + group * 32;
        mbar_wait(&bars->s_full[sb], (n2 / SP) & 1);
        tc_fence_after();
        float pr[32];
        {
          uint32_t r[32];
          tmem_ld_32x32(taddr, r);
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
        uint32_t ph[16], pl[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          __half2 hi, lo;
          split_f32x2(pr[2 * e], pr[2 * e + 1], hi, lo);        // key 2e in the low half, key 2e+1 in the high half
          ph[e] = *reinterpret_cast<const uint32_t*>(&hi);
          pl[e] = *reinterpret_cast<const uint32_t*>(&lo);
        }
        tmem_st_32x16(taddr, ph);
        tmem_st_32x16(taddr + 16, pl);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->p_full[sb]);
      }
      // ---- hand the partial row sum to the epilogue warps (they add the two groups in a fixed order)
      bars->xch[group][row] = l;
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->l_full);               // release: orders the stores above
    }
  } else {
    // ================================ epilogue warps ================================
    // out = O / l + residual for the 32 query rows of this warp's TMEM lane quarter, all DVT channels
    // (attn_epilogue_item above).
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t tmem_o = tmem_O + ((uint32_t)(quarter * 32) << 16);
    const uint32_t stg = smem_u32(sE + (warp - ATS_EPI_WARP0) * ATS_EPI_STAGE);
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

cudaError_t attention_ts_launch(int dvt, bool qt, int grid, cudaStream_t stream, bool short_launch, const CUtensorMap& mq_h,
                                const CUtensorMap& mq_l, const CUtensorMap& mk_h, const CUtensorMap& mk_l,
                                const CUtensorMap& mv_h, const CUtensorMap& mv_l, const AttnParams& p) {
  // the > 48 KB dynamic shared-memory opt-in is a per-device function attribute: set it once per device
  static PerDeviceFlag attr_set;
  const int slot = current_device_slot();
  if (!attr_set.is_set(slot)) {
    cudaError_t e;
#define ATS_OPT_IN(K) \
    if ((e = cudaFuncSetAttribute(K, cudaFuncAttributeMaxDynamicSharedMemorySize, ATS_SMEM_BYTES)) != cudaSuccess) return e
    ATS_OPT_IN((tc_attn_ts_kernel<128, false>));
    ATS_OPT_IN((tc_attn_ts_kernel<256, false>));
    ATS_OPT_IN((tc_attn_ts_kernel<128, true>));
    ATS_OPT_IN((tc_attn_ts_kernel<256, true>));
#undef ATS_OPT_IN
    attr_set.set(slot);
  }
#define ATS_LAUNCH(K) tc_launch(K, grid, ATS_THREADS, ATS_SMEM_BYTES, stream, short_launch, mq_h, mq_l, mk_h, mk_l, mv_h, mv_l, p)
  if (qt) return dvt == 256 ? ATS_LAUNCH((tc_attn_ts_kernel<256, true>)) : ATS_LAUNCH((tc_attn_ts_kernel<128, true>));
  return dvt == 256 ? ATS_LAUNCH((tc_attn_ts_kernel<256, false>)) : ATS_LAUNCH((tc_attn_ts_kernel<128, false>));
#undef ATS_LAUNCH
}

}  // namespace tdn
