// Epilogue of the fused attention kernels that hand P to the P.V' MMAs through tensor memory (tc_attn_ts.cu,
// tc_attn_ts2.cu): out = O / l + residual, coalesced through a per-warp shared-memory turn-around block.
#pragma once
#include "tc_attn.cuh"

namespace tdn {

constexpr int ATS_EPI_STAGE = 4096;                // per epilogue warp: 32 rows x 128 B turn-around block

// Epilogue of one work item for one warp (32 query rows = one TMEM lane quarter, 32-channel chunks [c_begin, c_end)):
//   out = O * inv + residual.
// O arrives with one query row per thread (TMEM lane = row), but a warp-wide 16-byte global access with one ROW per
// thread touches 32 cache lines (ncu: that epilogue was bound by the L1 tag stage).  So global memory is accessed
// with 4 (fp16 planes: 64-byte row pieces) or 8 (fp32: 128-byte row pieces) consecutive lanes per row -- 8 or 4 lines
// per instruction -- and a 4 KB shared-memory block per warp turns between the two arrangements: residual global ->
// registers (one chunk ahead) -> block -> row per thread; result row per thread -> block -> global.  The 16-byte
// pieces are XOR-swizzled so that both arrangements are free of bank conflicts.
// RES: 0 none, 1 SPLIT16, 2 fp32.  OUT16: SPLIT16 output, else fp32.  q0: first query row of the warp.
template <int RES, bool OUT16>
__device__ __forceinline__ void attn_epilogue_item(const AttnParams& p, uint32_t tmem_o, uint32_t stg, int lane, int q0,
                                                   long long rbase, long long obase, float inv, int c_begin,
                                                   int c_end, bool& out_of_range) {
  rbase += c_begin * 32;
  obase += c_begin * 32;
  // "piece" arrangement: fp16 plane = 32 rows x 64 B, lanes 4r..4r+3 per row; fp32 = 32 rows x 128 B, lanes 8r..8r+7
  const int rA = lane >> 2, cA = lane & 3, rB = lane >> 3, cB = lane & 7;
  const uint32_t pieceA = stg + rA * 64 + ((cA ^ ((rA >> 1) & 3)) << 4);      // row 8k + rA: + 512 k   (lo plane + 2048)
  const uint32_t pieceB0 = stg + rB * 128 + ((cB ^ rB) << 4);                  // row 4k + rB, k even: + 512 k
  const uint32_t pieceB1 = stg + rB * 128 + ((cB ^ (rB + 4)) << 4);            //               k odd:  + 512 k
  // "row" arrangement: thread = row, 16-byte piece q
  const uint32_t rowA = stg + lane * 64, xA = (lane >> 1) & 3;
  const uint32_t rowB = stg + lane * 128, xB = lane & 7;

  constexpr int NR = RES == 2 ? 8 : 4, NO = OUT16 ? 4 : 8;
  const __half* rh[4];
  const float* rf[8];
  __half* oh[4];
  float* of[8];
  uint32_t rmask = 0, omask = 0;
  const long long dres_lo = RES == 1 ? (p.res_lo - p.res_hi) : 0, dout_lo = OUT16 ? (p.out_lo - p.out_hi) : 0;
#pragma unroll
  for (int k = 0; k < NR; ++k) {
    if (RES == 1) {
      const int q = q0 + 8 * k + rA;
      rh[k] = p.res_hi + rbase + (long long)q * p.r_ld + cA * 8;
      rmask |= (q < p.Pq ? 1u : 0u) << k;
    } else if (RES == 2) {
      const int q = q0 + 4 * k + rB;
      rf[k] = p.res_f32 + rbase + (long long)q * p.r_ld + cB * 4;
      rmask |= (q < p.Pq ? 1u : 0u) << k;
    }
  }
#pragma unroll
  for (int k = 0; k < NO; ++k) {
    if (OUT16) {
      const int q = q0 + 8 * k + rA;
      oh[k] = p.out_hi + obase + (long long)q * p.o_ld + cA * 8;
      omask |= (q < p.Pq ? 1u : 0u) << k;
    } else {
      const int q = q0 + 4 * k + rB;
      of[k] = p.out_f32 + obase + (long long)q * p.o_ld + cB * 4;
      omask |= (q < p.Pq ? 1u : 0u) << k;
    }
  }
  uint4 rbuf[8] = {};            // residual of one chunk in the piece arrangement: [hi k=0..3 | lo k=0..3] or fp32 k=0..7
  auto load_res = [&]() {
    if (RES == 1) {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (rmask >> k & 1) {
          rbuf[k] = __ldg(reinterpret_cast<const uint4*>(rh[k]));
          rbuf[4 + k] = __ldg(reinterpret_cast<const uint4*>(rh[k] + dres_lo));
          rh[k] += 32;
        }
    } else if (RES == 2) {
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (rmask >> k & 1) {
          rbuf[k] = __ldg(reinterpret_cast<const uint4*>(rf[k]));
          rf[k] += 32;
        }
    }
  };
  load_res();
#pragma unroll 1
  for (int chunk = c_begin; chunk < c_end; ++chunk) {
    // residual of this chunk -> block; its registers then take the loads of the next chunk, which stay in flight
    // while this chunk is processed
    if (RES == 1) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        sts128(pieceA + 512 * k, rbuf[k]);
        sts128(pieceA + 2048 + 512 * k, rbuf[4 + k]);
      }
    } else if (RES == 2) {
#pragma unroll
      for (int k = 0; k < 8; ++k) sts128(((k & 1) ? pieceB1 : pieceB0) + 512 * k, rbuf[k]);
    }
    if (RES != 0) {
      __syncwarp();
      if (chunk + 1 < c_end) load_res();
    }
    uint32_t r[32];
    tmem_ld_32x32(tmem_o + chunk * 32, r);
    tmem_ld_wait();
    // out = (O * inv) + (hi + lo), same rounding order as the other epilogues
    auto put16 = [&](int q, const float (&v)[8]) {           // 8 channels -> SPLIT16 pieces q of this thread's row
      __half2 hi[4], lo[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        out_of_range |= fmaxf(fabsf(v[2 * e]), fabsf(v[2 * e + 1])) > 60000.f;
        split_f32x2(v[2 * e], v[2 * e + 1], hi[e], lo[e]);
      }
      sts128(rowA + ((q ^ xA) << 4), *reinterpret_cast<const uint4*>(hi));
      sts128(rowA + 2048 + ((q ^ xA) << 4), *reinterpret_cast<const uint4*>(lo));
    };
    auto put32 = [&](int q, float a, float b, float c, float d) {   // 4 channels -> fp32 piece q of this thread's row
      uint4 o4;
      o4.x = __float_as_uint(a); o4.y = __float_as_uint(b); o4.z = __float_as_uint(c); o4.w = __float_as_uint(d);
      sts128(rowB + ((q ^ xB) << 4), o4);
    };
    auto res16 = [&](int q, float (&v)[8]) {                 // v = O * inv + residual for channels 8q .. 8q+7
      const uint4 h4 = lds128(rowA + ((q ^ xA) << 4));
      const uint4 l4 = lds128(rowA + 2048 + ((q ^ xA) << 4));
      const __half2* hh = reinterpret_cast<const __half2*>(&h4);
      const __half2* ll = reinterpret_cast<const __half2*>(&l4);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 a = __half22float2(hh[e]), b2 = __half22float2(ll[e]);
        v[e * 2 + 0] = __fadd_rn(__fmul_rn(__uint_as_float(r[q * 8 + e * 2 + 0]), inv), a.x + b2.x);
        v[e * 2 + 1] = __fadd_rn(__fmul_rn(__uint_as_float(r[q * 8 + e * 2 + 1]), inv), a.y + b2.y);
      }
    };
    if (RES == 0 || (RES == 1) == OUT16) {
      // no residual, or residual and output in the same format: every thread rewrites its own row pieces in place
      if (OUT16) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float v[8];
          if (RES == 1) {
            res16(q, v);
          } else {
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = __uint_as_float(r[q * 8 + e]) * inv;
          }
          put16(q, v);
        }
      } else {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          float4 f = make_float4(0.f, 0.f, 0.f, 0.f);
          if (RES == 2) {
            const uint4 f4 = lds128(rowB + ((q ^ xB) << 4));
            f = make_float4(__uint_as_float(f4.x), __uint_as_float(f4.y), __uint_as_float(f4.z), __uint_as_float(f4.w));
          }
          put32(q, __fadd_rn(__fmul_rn(__uint_as_float(r[q * 4 + 0]), inv), f.x),
                __fadd_rn(__fmul_rn(__uint_as_float(r[q * 4 + 1]), inv), f.y),
                __fadd_rn(__fmul_rn(__uint_as_float(r[q * 4 + 2]), inv), f.z),
                __fadd_rn(__fmul_rn(__uint_as_float(r[q * 4 + 3]), inv), f.w));
        }
      }
    } else {
      // formats differ: all residual rows are read before any result row is written
      float v[32];
      if (RES == 1) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float t8[8];
          res16(q, t8);
#pragma unroll
          for (int e = 0; e < 8; ++e) v[q * 8 + e] = t8[e];
        }
      } else {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const uint4 f4 = lds128(rowB + ((q ^ xB) << 4));
          v[q * 4 + 0] = __fadd_rn(__fmul_rn(__uint_as_float(r[q * 4 + 0]), inv), __uint_as_float(f4.x));
          v[q * 4 + 1] = __fadd_rn(__fmul_rn(__uint_as_float(r[q * 4 + 1]), inv), __uint_as_float(f4.y));
          v[q * 4 + 2] = __fadd_rn(__fmul_rn(__uint_as_float(r[q * 4 + 2]), inv), __uint_as_float(f4.z));
          v[q * 4 + 3] = __fadd_rn(__fmul_rn(__uint_as_float(r[q * 4 + 3]), inv), __uint_as_float(f4.w));
        }
      }
      __syncwarp();
      if (OUT16) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float t8[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) t8[e] = v[q * 8 + e];
          put16(q, t8);
        }
      } else {
#pragma unroll
        for (int q = 0; q < 8; ++q) put32(q, v[q * 4 + 0], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
      }
    }
    __syncwarp();
    // result block -> global with 4 / 8 lanes per row
    if (OUT16) {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (omask >> k & 1) {
          *reinterpret_cast<uint4*>(oh[k]) = lds128(pieceA + 512 * k);
          *reinterpret_cast<uint4*>(oh[k] + dout_lo) = lds128(pieceA + 2048 + 512 * k);
          oh[k] += 32;
        }
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (omask >> k & 1) {
          *reinterpret_cast<uint4*>(of[k]) = lds128(((k & 1) ? pieceB1 : pieceB0) + 512 * k);
          of[k] += 32;
        }
    }
    __syncwarp();                                            // the block is rewritten by the next chunk
  }
}


// Dispatch on the residual / output formats of the launch (fmt = residual kind * 2 + (SPLIT16 output)).
__device__ __forceinline__ void attn_epilogue_dispatch(int fmt, const AttnParams& p, uint32_t tmem_o, uint32_t stg, int lane,
                                                       int q0, long long rbase, long long obase, float inv, int c_begin,
                                                       int c_end, bool& out_of_range) {
  switch (fmt) {
    case 0: attn_epilogue_item<0, false>(p, tmem_o, stg, lane, q0, rbase, obase, inv, c_begin, c_end, out_of_range); break;
    case 1: attn_epilogue_item<0, true>(p, tmem_o, stg, lane, q0, rbase, obase, inv, c_begin, c_end, out_of_range); break;
    case 2: attn_epilogue_item<1, false>(p, tmem_o, stg, lane, q0, rbase, obase, inv, c_begin, c_end, out_of_range); break;
    case 3: attn_epilogue_item<1, true>(p, tmem_o, stg, lane, q0, rbase, obase, inv, c_begin, c_end, out_of_range); break;
    case 4: attn_epilogue_item<2, false>(p, tmem_o, stg, lane, q0, rbase, obase, inv, c_begin, c_end, out_of_range); break;
    default: attn_epilogue_item<2, true>(p, tmem_o, stg, lane, q0, rbase, obase, inv, c_begin, c_end, out_of_range); break;
  }
}
__device__ __forceinline__ int attn_epilogue_fmt(const AttnParams& p) {
  return (p.res_hi ? 1 : p.res_f32 ? 2 : 0) * 2 + (p.out_hi ? 1 : 0);
}

}  // namespace tdn
