// Pass kernel for NL x 32-bit limb fields (NL = 12: 384-bit containers), SURVEY.md par. 8f row 4.
// Same pass / round / exchange structure as ntt_pass_body in ntt_kernels.cuh, driven by the same
// planner (plan.h), with the field layer of field_n.cuh.  Kept to what a first wider field needs:
// in-place passes, both twiddle table paths (one or two lookups), inverse scaling in the last pass;
// no peer stores, no streamed twiddles.  Four elements per thread (48 data registers at NL = 12).
// An element is NL/4 consecutive uint4 (element-major, little-endian limbs), like cgbn_mem_t<bits>
// (/root/reference/include/cgbn/cgbn_cuda.h:51-55) for bits = 32*NL.
#pragma once
#include <cuda_runtime.h>

#include "field_n.cuh"
#include "kernel_common.cuh"
#include "plan.h"

namespace ntt256 {

template <int NL>
struct PassArgsN {
  FieldParamsN<NL> f;
  PassDesc pd;
  const uint4* om_hi;  // w_N^(i << lo_bits), Montgomery form
  const uint4* om_lo;  // w_N^i, i < 2^lo_bits
  const uint4* loc;    // w_{2^loc_bits}^i (round twiddles)
  int hi_bits, lo_bits, loc_bits;
  int inverse;
  u32 c8[3][NL];  // w_8^1..3 (or inverses), Montgomery form
  u32 scale[NL];  // 1/n, Montgomery form (pd.scale_out)
  unsigned long long batch_stride;
  // inverse: round `scaled_round` of this pass reads loc_scaled = (1/n) * loc and multiplies every
  // element (also the one whose twiddle is 1), which applies the 1/n without an extra product
  int scaled_round;
  const uint4* loc_scaled;
};

template <int NL>
NTT_DEV void ld_el(u32* v, const uint4* p) {
#pragma unroll
  for (int q = 0; q < NL / 4; ++q) {
    const uint4 t = p[q];
    v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
  }
}
template <int NL>
NTT_DEV void st_el(uint4* p, const u32* v) {
#pragma unroll
  for (int q = 0; q < NL / 4; ++q) p[q] = make_uint4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
}
template <int NL>
NTT_DEV void ld_table_n(u32* v, const uint4* tab, u32 idx) {
#pragma unroll
  for (int q = 0; q < NL / 4; ++q) {
    const uint4 t = __ldg(tab + (NL / 4) * (size_t)idx + q);
    v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
  }
}

template <int NL>
NTT_DEV void mul_by_n(u32* x, const u32* w, const FieldParamsN<NL>& f) {
  u32 r[NL];
  mont_mul_n<NL>(r, w, x, f);
#pragma unroll
  for (int k = 0; k < NL; ++k) x[k] = r[k];
}

template <int NL>
NTT_DEV void bfly_n(u32* a, u32* b, const FieldParamsN<NL>& f) {
  u32 s[NL], d[NL];
  add_lazy_n<NL>(s, a, b, f);
  sub_lazy_n<NL>(d, a, b, f);
#pragma unroll
  for (int k = 0; k < NL; ++k) {
    a[k] = s[k];
    b[k] = d[k];
  }
}

// NB DIF stages on the top NB bits of the GB-bit register index + the round twiddle (cf. round_compute).
template <int NL, int GB, int NB>
NTT_DEV void round_compute_n(u32 (*x)[NL], const PassArgsN<NL>& A, bool last, bool scaled, u32 lam0, int sb, int lowp) {
  const FieldParamsN<NL>& f = A.f;
#pragma unroll
  for (int k = 0; k < NB; ++k) {
    const int bit = GB - 1 - k;
#pragma unroll
    for (int u = 0; u < (1 << GB); ++u) {
      if (u & (1 << bit)) continue;
      bfly_n<NL>(x[u], x[u | (1 << bit)], f);
      const int t = (u >> (GB - NB)) & ((1 << (NB - 1 - k)) - 1);
      const int e8 = t << (3 - (NB - k));
      if (e8) mul_by_n<NL>(x[u | (1 << bit)], A.c8[e8 - 1], f);
    }
  }
  if (last) return;
  const int pb = sb + GB - NB;
  const int m = sb + GB - lowp;
  const int sh = A.loc_bits - m;
  const u32 tl_mask = (1u << (pb - lowp)) - 1u;
  const u32 tl0 = (lam0 >> lowp) & tl_mask;
  const int ush = sb - lowp;
  const uint4* loc_tab = scaled ? A.loc_scaled : A.loc;
#pragma unroll
  for (int u = 0; u < (1 << GB); ++u) {
    const int uproc = u >> (GB - NB);
    int v = 0;
#pragma unroll
    for (int i = 0; i < NB; ++i) v |= ((uproc >> i) & 1) << (NB - 1 - i);
    const u32 rest = u & ((1 << (GB - NB)) - 1);
    if (scaled || uproc != 0) {
      const u32 tl = (tl0 | (rest << ush)) & tl_mask;
      u32 e = (tl * (u32)v) << sh;
      if (A.inverse) e = (0u - e);
      e &= (1u << A.loc_bits) - 1u;
      u32 w[NL];
      ld_table_n<NL>(w, loc_tab, e);
      mul_by_n<NL>(x[u], w, f);
    }
  }
}

template <int NL, int GB>
NTT_DEV void ntt_pass_body_n(uint4* data, const PassArgsN<NL>& A, u32* smem) {
  constexpr int NE = 1 << GB;
  constexpr int Q = NL / 4;  // uint4 per element
  const PassDesc& pd = A.pd;
  const FieldParamsN<NL>& f = A.f;
  const int T = pd.T;
  const int d = pd.d;
  const u32 tid = threadIdx.x;
  const u32 tile_words = 1u << T;

  const unsigned long long blk = blockIdx.x;
  unsigned long long base = deposit64(blk, pd.blk_pos, pd.nblk_bits);
  u32 jlow_blk = pd.jlow_const;
  for (int i = 0; i < pd.nblk_bits; ++i) jlow_blk += (u32)((blk >> i) & 1ull) * pd.wt_blk[i];
  base += (unsigned long long)blockIdx.y * A.batch_stride;
  uint4* gbase = data + Q * base;

  u32 x[NE][NL];
  const int lowp = T - d;

  for (int r = 0; r < pd.nrounds; ++r) {
    const int sb = pd.round_sb[r];
    const int nb = pd.round_nb[r];
    const u32 lam0 = deposit32(tid, pd.tmap[r], T - GB);
    const bool first = (r == 0);
    const bool last = (r == pd.nrounds - 1);

    if (first) {
      const u32 g0 = deposit32(lam0, pd.in_pos, T);
      u32 st[GB];
#pragma unroll
      for (int i = 0; i < GB; ++i) st[i] = 1u << pd.in_pos[sb + i];
#pragma unroll
      for (int u = 0; u < NE; ++u) {
        u32 off = g0;
#pragma unroll
        for (int i = 0; i < GB; ++i) off += ((u >> i) & 1) ? st[i] : 0u;
        ld_el<NL>(x[u], gbase + Q * (size_t)off);
      }
    } else {
      const u32 w0 = swz(lam0);
#pragma unroll
      for (int u = 0; u < NE; ++u) {
        const u32 w = w0 ^ swz((u32)u << sb);
#pragma unroll
        for (int k = 0; k < NL; ++k) x[u][k] = smem[k * tile_words + w];
      }
    }

    const bool scaled = (r == A.scaled_round);
    if (nb == 2) {
      if constexpr (GB >= 2) round_compute_n<NL, GB, 2>(x, A, last, scaled, lam0, sb, lowp);
    } else if (nb == 1) {
      round_compute_n<NL, GB, 1>(x, A, last, scaled, lam0, sb, lowp);
    }

    if (!last) {
      const u32 w0 = swz(lam0);
#pragma unroll
      for (int u = 0; u < NE; ++u) {
        const u32 w = w0 ^ swz((u32)u << sb);
#pragma unroll
        for (int k = 0; k < NL; ++k) smem[k * tile_words + w] = x[u][k];
      }
      __syncthreads();
    } else {
      const u32 g0 = deposit32(lam0, pd.out_pos, T);
      u32 st[GB];
#pragma unroll
      for (int i = 0; i < GB; ++i) st[i] = 1u << pd.out_pos[sb + i];
      if (pd.has_tw) {
        u32 jl0 = jlow_blk;
        for (int i = 0; i < lowp; ++i) jl0 += ((lam0 >> i) & 1u) * pd.wt_loc[i];
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
          u32 w[NL];
          ld_table_n<NL>(w, A.om_hi, e >> A.lo_bits);
          mul_by_n<NL>(x[u], w, f);
          if (two) {
            ld_table_n<NL>(w, A.om_lo, e & lomask);
            mul_by_n<NL>(x[u], w, f);
          }
        }
      }
      if (pd.scale_out) {
#pragma unroll
        for (int u = 0; u < NE; ++u) mul_by_n<NL>(x[u], A.scale, f);
      }
      if (pd.is_final) {
#pragma unroll
        for (int u = 0; u < NE; ++u) canonicalize_n<NL>(x[u], f);
      }
      if (first) __syncthreads();  // single-round pass: all loads before any permuted store
#pragma unroll
      for (int u = 0; u < NE; ++u) {
        u32 off = g0;
#pragma unroll
        for (int i = 0; i < GB; ++i) off += ((u >> i) & 1) ? st[i] : 0u;
        st_el<NL>(gbase + Q * (size_t)off, x[u]);
      }
    }
  }
}

__global__ void __launch_bounds__(512, 1)
ntt384_pass_kernel(uint4* data, const __grid_constant__ PassArgsN<12> A) {
  extern __shared__ u32 smem[];
  ntt_pass_body_n<12, 2>(data, A, smem);
}

// out[i] = prod over set bits k of i of pw[k] (pw[k] = w^(2^k), Montgomery form), i.e. w^i.
template <int NL>
__global__ void twiddle_powers_kernel_n(uint4* __restrict__ out, const uint4* __restrict__ pw, int nbits,
                                        unsigned long long count, const __grid_constant__ FieldParamsN<NL> f,
                                        const __grid_constant__ FeN<NL> one) {
  const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  u32 acc[NL];
#pragma unroll
  for (int k = 0; k < NL; ++k) acc[k] = one.v[k];
  for (int b = 0; b < nbits; ++b) {
    if ((i >> b) & 1ull) {
      u32 w[NL];
      ld_table_n<NL>(w, pw, (u32)b);
      mul_by_n<NL>(acc, w, f);
    }
  }
  canonicalize_n<NL>(acc, f);
  st_el<NL>(out + (NL / 4) * i, acc);
}

// out[i] = c * in[i] (c a Montgomery-form constant): scaled copies of small tables.
template <int NL>
__global__ void table_scale_kernel_n(uint4* __restrict__ out, const uint4* __restrict__ in, unsigned long long count,
                                     const __grid_constant__ FieldParamsN<NL> f, const __grid_constant__ FeN<NL> c) {
  const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  u32 x[NL], t[NL];
  ld_el<NL>(x, in + (NL / 4) * i);
#pragma unroll
  for (int k = 0; k < NL; ++k) t[k] = c.v[k];
  mul_by_n<NL>(x, t, f);
  canonicalize_n<NL>(x, f);
  st_el<NL>(out + (NL / 4) * i, x);
}

// Element-wise product: mont_form = 1: out = a*b/R (Montgomery-form data in and out);
// mont_form = 0: canonical data, out = a*b (second product by R^2).
template <int NL>
__global__ void vec_mul_kernel_n(uint4* __restrict__ out, const uint4* __restrict__ a, const uint4* __restrict__ b,
                                 unsigned long long n, int mont_form, const __grid_constant__ FieldParamsN<NL> f,
                                 const __grid_constant__ FeN<NL> r2) {
  const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  u32 x[NL], y[NL];
  ld_el<NL>(x, a + (NL / 4) * i);
  ld_el<NL>(y, b + (NL / 4) * i);
  mul_by_n<NL>(x, y, f);  // needs y < p: canonical or Montgomery-form operands are
  if (!mont_form) {
    canonicalize_n<NL>(x, f);
    u32 t[NL];
#pragma unroll
    for (int k = 0; k < NL; ++k) t[k] = r2.v[k];
    mul_by_n<NL>(x, t, f);
  }
  canonicalize_n<NL>(x, f);
  st_el<NL>(out + (NL / 4) * i, x);
}

// Transforms of 2 or 4 points, one thread each (exact radix-2 DIF + bit reversal).
template <int NL>
struct TinyArgsN {
  FieldParamsN<NL> f;
  int log_n;
  u32 w4[NL];  // w_4 (or its inverse), Montgomery form
  int do_scale;
  u32 scale[NL];
  unsigned long long batch;
};
template <int NL>
__global__ void ntt_tiny_kernel_n(uint4* data, const __grid_constant__ TinyArgsN<NL> A) {
  const unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= A.batch) return;
  constexpr int Q = NL / 4;
  const FieldParamsN<NL>& f = A.f;
  if (A.log_n == 1) {
    u32 a[NL], b[NL];
    uint4* p = data + Q * 2 * t;
    ld_el<NL>(a, p);
    ld_el<NL>(b, p + Q);
    bfly_n<NL>(a, b, f);
    if (A.do_scale) {
      mul_by_n<NL>(a, A.scale, f);
      mul_by_n<NL>(b, A.scale, f);
    }
    canonicalize_n<NL>(a, f);
    canonicalize_n<NL>(b, f);
    st_el<NL>(p, a);
    st_el<NL>(p + Q, b);
  } else {
    u32 x[4][NL];
    uint4* p = data + Q * 4 * t;
#pragma unroll
    for (int i = 0; i < 4; ++i) ld_el<NL>(x[i], p + Q * i);
    bfly_n<NL>(x[0], x[2], f);
    bfly_n<NL>(x[1], x[3], f);
    mul_by_n<NL>(x[3], A.w4, f);
    bfly_n<NL>(x[0], x[1], f);
    bfly_n<NL>(x[2], x[3], f);
    // outputs in bit-reversed order: X0 = x0, X2 = x1, X1 = x2, X3 = x3
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (A.do_scale) mul_by_n<NL>(x[i], A.scale, f);
      canonicalize_n<NL>(x[i], f);
    }
    st_el<NL>(p, x[0]);
    st_el<NL>(p + Q, x[2]);
    st_el<NL>(p + 2 * Q, x[1]);
    st_el<NL>(p + 3 * Q, x[3]);
  }
}

}  // namespace ntt256
