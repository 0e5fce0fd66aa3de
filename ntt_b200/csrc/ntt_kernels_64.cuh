// Pass kernels of the 64-bit transforms (include/ntt64.h; SURVEY.md par. 8f row 4, second half): the same
// self-sorting in-place passes as the 256-bit library (plan.h: a CTA owns a tile of 2^T elements picked by T
// index bits, several DIF stages per register round, exchange through swizzled shared memory, the bit
// permutation fused into the stores -- no bit-reversal kernel, cf. /root/reference/src/big-num.cu:57-65), over
// one-u64 elements.  An element is two registers, so a thread keeps SIXTEEN of them (radix-16 rounds: four
// stages between exchanges) and a tile of 2^12 elements is 32 KiB of shared memory (two 32-bit planes, same
// swizzle as the limb planes of the wide kernels), four CTAs of 256 threads per SM.
//
// Bytes: 8 B read + 8 B written per element per pass, warps touching runs of 2^g consecutive elements (g =
// plan coalesce_bits, default 4 = 128-byte lines).
#pragma once
#include <cuda_runtime.h>

#include "field64.cuh"
#include "kernel_common.cuh"
#include "ntt_pass_generic.cuh"
#include "plan.h"

namespace ntt64 {

using ntt256::PassDesc;
using ntt256::deposit32;
using ntt256::deposit64;
using ntt256::swz;

struct PassArgs64 {
  Params64 f;
  PassDesc pd;
  const u64* om_hi;  // w_N^(i << lo_bits), i < 2^hi_bits   (N = 2^max_log_n; table format of the field, field64.cuh)
  const u64* om_lo;  // w_N^i, i < 2^lo_bits
  const u64* loc;    // w_{2^loc_bits}^i: round twiddles
  int hi_bits, lo_bits, loc_bits;
  int inverse;       // 1: exponents are negated
  u64 c16[7];        // w_16^1 .. w_16^7 (or their inverses)
  u64 scale;         // 1/n (pd.scale_out)
  unsigned long long batch_stride;
  int scaled_round;       // inverse: this round of the pass reads loc_scaled = (1/n) * loc and multiplies every element
  const u64* loc_scaled;
  // element-wise companions fused into the end passes (coset shift in, coset shift / pointwise factor out)
  const u64* in_scale;
  const u64* out_scale;
  // Index tables of this pass (built on the host with the plan, capi64.cu build_pass_tables): with sixteen cheap
  // butterflies' worth of work per element the bit deposits of the wide kernels (a loop over T index bits per
  // thread and round) would be a quarter of all instructions here, so they are looked up instead.
  //   thr[0 .. t)          element offset of the thread's first element when loading  (deposit of round 0's lam0 by in_pos)
  //   thr[t .. 2t)         ... when storing (last round's lam0 by out_pos)
  //   thr[2t .. 3t)        the thread's share of j_low (passive local bits below the processed ones, last round)
  //   thr[3t + r t ...)    round r: lam0 | swz(lam0) << 16
  //   blk[b]               .x = index of the tile's first element inside one transform, .y = the block's share of j_low
  const u32* thr;
  const uint2* blk;
  // Fused exchange (multi-GPU four-step, last pass of step 1; same scheme as the wide kernels): when peer_bits > 0
  // the tile is stored straight into the receive buffers of the destination ranks over NVLink (peer-mapped
  // pointers): destination = top peer_bits of the local output index, remote index = (my rank << peer_shift) |
  // (low peer_shift bits).
  unsigned long long peer[8];
  int peer_bits, peer_shift, peer_rank;
};

// K16 != 0 (Goldilocks only): w_16 = 2^(12 K16), the constant twiddles are shifts (field64.cuh gl_sub_mul_w16)
template <bool GL, int K16, int GB, int NB>
NTT_DEV void round_compute64(u64* x, const PassArgs64& A, bool last, bool scaled, u32 lam0, int sb, int lowp) {
  using F = F64<GL>;
  const Params64& f = A.f;
#pragma unroll
  for (int k = 0; k < NB; ++k) {
    const int bit = GB - 1 - k;
#pragma unroll
    for (int u = 0; u < (1 << GB); ++u) {
      if (u & (1 << bit)) continue;
      const int v = u | (1 << bit);
      const u64 a = x[u], b = x[v];
      x[u] = F::add(a, b, f);
      const int t = (u >> (GB - NB)) & ((1 << (NB - 1 - k)) - 1);  // lower processed bits of u
      const int e16 = t << (4 - (NB - k));                         // constant twiddle as a power of w_16
      u64 dd;
      if constexpr (GL && K16 != 0) {
        dd = e16 ? gl_sub_mul_w16<K16>(a, b, e16) : F::sub(a, b, f);
      } else {
        dd = F::sub(a, b, f);
        if (e16) dd = F::mul(A.c16[e16 - 1], dd, f);
      }
      x[v] = dd;
    }
  }
  if (last || NB == 0) return;
  const int pb = sb + GB - NB;   // lowest bit butterflied in this round
  const int m = sb + GB - lowp;  // process bits from this group's top down
  const int sh = A.loc_bits - m;
  const u32 tl_mask = (1u << (pb - lowp)) - 1u;
  const u32 tl0 = (lam0 >> lowp) & tl_mask;
  const int ush = sb - lowp;
  const u64* loc_tab = scaled ? A.loc_scaled : A.loc;
  const u32 loc_mask = (1u << A.loc_bits) - 1u;
  // shift and sign of the table index once per round (on tau_lo): one constant multiply and one mask per element
  u32 tbase[1 << (GB - NB)];
#pragma unroll
  for (int rst = 0; rst < (1 << (GB - NB)); ++rst) {
    const u32 t0 = ((tl0 | ((u32)rst << ush)) & tl_mask) << sh;
    tbase[rst] = A.inverse ? (0u - t0) : t0;
  }
#pragma unroll
  for (int u = 0; u < (1 << GB); ++u) {
    const int uproc = u >> (GB - NB);
    int v = 0;
#pragma unroll
    for (int i = 0; i < NB; ++i) v |= ((uproc >> i) & 1) << (NB - 1 - i);
    const u32 rest = u & ((1 << (GB - NB)) - 1);
    if (scaled || uproc != 0) {
      const u32 e = (tbase[rest] * (u32)v) & loc_mask;
      x[u] = F::mul(__ldg(loc_tab + e), x[u], f);
    }
  }
}

template <bool GL, int K16, int GB>
NTT_DEV void ntt_pass_body64(const u64* data_in, u64* data_out, const PassArgs64& A, u32* smem) {
  using F = F64<GL>;
  constexpr int NE = 1 << GB;
  const PassDesc& pd = A.pd;
  const Params64& f = A.f;
  const int T = pd.T;
  const int d = pd.d;
  const u32 tid = threadIdx.x;
  const u32 tile_words = 1u << T;

  const uint2 bt = __ldg(A.blk + blockIdx.x);
  const unsigned long long base_t = bt.x;  // index inside one transform
  const u32 jlow_blk = bt.y;
  const unsigned long long base = base_t + (unsigned long long)blockIdx.y * A.batch_stride;
  const u32 nthr = blockDim.x;
  const u64* gbase_in = data_in + base;
  u64* gbase = data_out + base;

  u64 x[NE];
  const int lowp = T - d;

  // programmatic dependent launch (see ntt_pass_generic.cuh): nothing above touches mutable data -- the index
  // tables are written once, at plan creation
  ntt256::grid_dependency_wait();
  for (int r = 0; r < pd.nrounds; ++r) {
    const int sb = pd.round_sb[r];
    const int nb = pd.round_nb[r];
    const u32 lw = __ldg(A.thr + (3 + r) * nthr + tid);
    const u32 lam0 = lw & 0xffffu;
    const u32 w0 = lw >> 16;  // swz(lam0)
    const bool first = (r == 0);
    const bool last = (r == pd.nrounds - 1);
    // swz is linear over XOR: the swizzled offset of element u is w0 ^ (XOR of sx[i] over the set bits i of u)
    u32 sx[GB];
#pragma unroll
    for (int i = 0; i < GB; ++i) sx[i] = swz(1u << (sb + i));

    if (first) {
      const u32 g0 = __ldg(A.thr + tid);
      u32 st[GB];
#pragma unroll
      for (int i = 0; i < GB; ++i) st[i] = 1u << pd.in_pos[sb + i];
#pragma unroll
      for (int u = 0; u < NE; ++u) {
        u32 off = g0;
#pragma unroll
        for (int i = 0; i < GB; ++i) off += ((u >> i) & 1) ? st[i] : 0u;
        x[u] = gbase_in[off];
        if (A.in_scale != nullptr) x[u] = F::mul(__ldg(A.in_scale + (u32)base_t + off), x[u], f);
      }
    } else {
#pragma unroll
      for (int u = 0; u < NE; ++u) {
        u32 w = w0;
#pragma unroll
        for (int i = 0; i < GB; ++i) w ^= ((u >> i) & 1) ? sx[i] : 0u;
        x[u] = pack(smem[w], smem[tile_words + w]);
      }
    }

    {
      const bool scaled = (r == A.scaled_round);
      if (nb == 4) {
        if constexpr (GB >= 4) round_compute64<GL, K16, GB, 4>(x, A, last, scaled, lam0, sb, lowp);
      } else if (nb == 3) {
        if constexpr (GB >= 3) round_compute64<GL, K16, GB, 3>(x, A, last, scaled, lam0, sb, lowp);
      } else if (nb == 2) {
        if constexpr (GB >= 2) round_compute64<GL, K16, GB, 2>(x, A, last, scaled, lam0, sb, lowp);
      } else if (nb == 1) {
        round_compute64<GL, K16, GB, 1>(x, A, last, scaled, lam0, sb, lowp);
      }
    }

    if (!last) {
#pragma unroll
      for (int u = 0; u < NE; ++u) {
        u32 w = w0;
#pragma unroll
        for (int i = 0; i < GB; ++i) w ^= ((u >> i) & 1) ? sx[i] : 0u;
        smem[w] = (u32)x[u];
        smem[tile_words + w] = (u32)(x[u] >> 32);
      }
      __syncthreads();
    } else {
      const u32 g0 = __ldg(A.thr + nthr + tid);
      u32 st[GB];
#pragma unroll
      for (int i = 0; i < GB; ++i) st[i] = 1u << pd.out_pos[sb + i];
      if (pd.has_tw) {
        const u32 jl0 = jlow_blk + __ldg(A.thr + 2 * nthr + tid);
        u32 gw[GB];
#pragma unroll
        for (int i = 0; i < GB; ++i) gw[i] = (sb + i < lowp) ? pd.wt_loc[sb + i] : 0u;
        const u32 nmask = (A.hi_bits + A.lo_bits >= 32) ? 0xffffffffu : ((1u << (A.hi_bits + A.lo_bits)) - 1u);
        const u32 lomask = (1u << A.lo_bits) - 1u;
        const bool two = pd.tw_lshift < A.lo_bits;
#pragma unroll
        for (int u = 0; u < NE; ++u) {
          const u32 lam = lam0 | ((u32)u << sb);
          const u32 khi = d ? (__brev(lam >> lowp) >> (32 - d)) : 0u;
          u32 jl = jl0;
#pragma unroll
          for (int i = 0; i < GB; ++i) jl += ((u >> i) & 1) ? gw[i] : 0u;
          u32 e = (jl * khi) << pd.tw_lshift;
          if (A.inverse) e = (0u - e);
          e &= nmask;
          x[u] = F::mul(__ldg(A.om_hi + (e >> A.lo_bits)), x[u], f);
          if (two) x[u] = F::mul(__ldg(A.om_lo + (e & lomask)), x[u], f);
        }
      }
      if (pd.scale_out) {
#pragma unroll
        for (int u = 0; u < NE; ++u) x[u] = F::mul(A.scale, x[u], f);
      }
      if (A.out_scale != nullptr) {
#pragma unroll
        for (int u = 0; u < NE; ++u) {
          u32 off = (u32)base_t + g0;
#pragma unroll
          for (int i = 0; i < GB; ++i) off += ((u >> i) & 1) ? st[i] : 0u;
          x[u] = F::mul(__ldg(A.out_scale + off), x[u], f);
        }
      }
      // a single-round pass loads and stores global memory in the same round: every thread of the CTA must
      // have loaded before any permuted store lands on another thread's input
      if (first) __syncthreads();
      if (A.peer_bits == 0) {
#pragma unroll
        for (int u = 0; u < NE; ++u) {
          u32 off = g0;
#pragma unroll
          for (int i = 0; i < GB; ++i) off += ((u >> i) & 1) ? st[i] : 0u;
          gbase[off] = x[u];
        }
      } else {
        const u32 lowmask = (1u << A.peer_shift) - 1u;
        const u32 mine = (u32)A.peer_rank << A.peer_shift;
#pragma unroll
        for (int u = 0; u < NE; ++u) {
          u32 off = (u32)base + g0;  // full local output index: block part + in-tile part
#pragma unroll
          for (int i = 0; i < GB; ++i) off += ((u >> i) & 1) ? st[i] : 0u;
          u64* dst = reinterpret_cast<u64*>(A.peer[off >> A.peer_shift]) + (size_t)(mine | (off & lowmask));
          *dst = x[u];
        }
      }
    }
  }
  ntt256::grid_dependency_launch_dependents();
}

// Sixteen elements per thread: the throughput shape (tiles of up to 2^14 elements, <= 1024 threads).
template <bool GL, int K16>
__global__ void __launch_bounds__(1024, 1)
ntt64_pass16_kernel(const u64* data_in, u64* data_out, const __grid_constant__ PassArgs64 A) {
  extern __shared__ u32 smem[];
  ntt_pass_body64<GL, K16, 4>(data_in, data_out, A, smem);
}
// The same with 256 threads at most (tiles of <= 2^12 elements), compiled for four resident CTAs per SM.
template <bool GL, int K16>
__global__ void __launch_bounds__(256, 4)
ntt64_pass16d_kernel(const u64* data_in, u64* data_out, const __grid_constant__ PassArgs64 A) {
  extern __shared__ u32 smem[];
  ntt_pass_body64<GL, K16, 4>(data_in, data_out, A, smem);
}
// Four elements per thread: small, latency-bound calls (more, shorter threads).
template <bool GL, int K16>
__global__ void __launch_bounds__(1024, 1)
ntt64_pass4_kernel(const u64* data_in, u64* data_out, const __grid_constant__ PassArgs64 A) {
  extern __shared__ u32 smem[];
  ntt_pass_body64<GL, K16, 2>(data_in, data_out, A, smem);
}

// out[i] = prod over set bits k of i of pw[k] (pw[k] = w^(2^k) in the field's table format), i.e. w^i;
// `one` is 1 in that format.
template <bool GL>
__global__ void twiddle_powers_kernel64(u64* __restrict__ out, const u64* __restrict__ pw, int nbits, unsigned long long count,
                                        Params64 f, u64 one) {
  const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  u64 acc = one;
  for (int b = 0; b < nbits; ++b)
    if ((i >> b) & 1ull) acc = F64<GL>::mul(pw[b], acc, f);
  out[i] = acc;
}

// out[i] = c * in[i] (c in table format): scaled copies of small tables
template <bool GL>
__global__ void table_scale_kernel64(u64* __restrict__ out, const u64* __restrict__ in, unsigned long long count, Params64 f, u64 c) {
  const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  out[i] = F64<GL>::mul(c, in[i], f);
}

// out[i] = a[i] * b[i] (plain data).  Montgomery instance: redc(redc(a b) * R^2).
template <bool GL>
__global__ void vec_mul_kernel64(u64* __restrict__ out, const u64* __restrict__ a, const u64* __restrict__ b, unsigned long long n,
                                 Params64 f, u64 r2) {
  const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  u64 v = F64<GL>::mul(a[i], b[i], f);
  if constexpr (!GL) v = m64_mul(v, r2, f);
  out[i] = v;
}

// data[j] *= s^(j mod 2^log_n) with s^j = hi[j >> lo_bits] * lo[j & mask] (tables in table format)
template <bool GL>
__global__ void coset_scale_kernel64(u64* __restrict__ data, unsigned long long total, int log_n, int lo_bits,
                                     const u64* __restrict__ hi, const u64* __restrict__ lo, Params64 f) {
  const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const u32 j = (u32)(i & ((1ull << log_n) - 1ull));
  u64 v = F64<GL>::mul(__ldg(hi + (j >> lo_bits)), data[i], f);
  v = F64<GL>::mul(__ldg(lo + (j & ((1u << lo_bits) - 1u))), v, f);
  data[i] = v;
}

// Transforms of 2 or 4 points, one thread each.
struct TinyArgs64 {
  Params64 f;
  int log_n;
  u64 w4;  // w_4 (or its inverse), table format
  int do_scale;
  u64 scale;
  unsigned long long batch;
};
template <bool GL>
__global__ void ntt64_tiny_kernel(u64* data, const __grid_constant__ TinyArgs64 A) {
  using F = F64<GL>;
  const unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= A.batch) return;
  const Params64& f = A.f;
  if (A.log_n == 1) {
    u64* p = data + 2 * t;
    const u64 a = p[0], b = p[1];
    u64 s = F::add(a, b, f), dd = F::sub(a, b, f);
    if (A.do_scale) {
      s = F::mul(A.scale, s, f);
      dd = F::mul(A.scale, dd, f);
    }
    p[0] = s;
    p[1] = dd;
  } else {
    u64* p = data + 4 * t;
    const u64 x0 = p[0], x1 = p[1], x2 = p[2], x3 = p[3];
    const u64 a0 = F::add(x0, x2, f), a2 = F::sub(x0, x2, f);
    const u64 a1 = F::add(x1, x3, f);
    const u64 a3 = F::mul(A.w4, F::sub(x1, x3, f), f);
    u64 y0 = F::add(a0, a1, f), y2 = F::sub(a0, a1, f);  // X0, X2
    u64 y1 = F::add(a2, a3, f), y3 = F::sub(a2, a3, f);  // X1, X3
    if (A.do_scale) {
      y0 = F::mul(A.scale, y0, f);
      y1 = F::mul(A.scale, y1, f);
      y2 = F::mul(A.scale, y2, f);
      y3 = F::mul(A.scale, y3, f);
    }
    p[0] = y0;
    p[1] = y1;
    p[2] = y2;
    p[3] = y3;
  }
}

}  // namespace ntt64
