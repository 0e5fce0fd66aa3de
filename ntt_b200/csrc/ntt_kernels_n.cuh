// Pass kernel for NL x 32-bit limb fields (NL = 12: 384-bit containers), SURVEY.md par. 8f row 4: the N-limb
// instance of the one pass body (ntt_pass_generic.cuh, the same code the 256-bit kernels are instances of), with
// the field layer of field_n.cuh -- so peer stores, rank synchronisation, streamed twiddles, out-of-place ends
// and the fused companions exist for this width too.  Four elements per thread (48 data registers at NL = 12).
// An element is NL/4 consecutive uint4 (element-major, little-endian limbs), like cgbn_mem_t<bits>
// (/root/reference/include/cgbn/cgbn_cuda.h:51-55) for bits = 32*NL.
#pragma once
#include <cuda_runtime.h>

#include "field_n.cuh"
#include "kernel_common.cuh"
#include "ntt_pass_generic.cuh"
#include "plan.h"

namespace ntt256 {

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

// The N-limb field policy of the pass body (ntt_pass_generic.cuh): Montgomery-form twiddles, values in [0, 2p)
// between butterflies, every correction exact.
template <int NLIMBS>
struct FN {
  static constexpr int NL = NLIMBS;
  static constexpr int TW = NLIMBS;
  using Params = FieldParamsN<NLIMBS>;
  using Args = PassArgsT<Params, NLIMBS, NLIMBS>;
  static NTT_DEV const Params& params(const Args& A) { return A.f; }
  static NTT_DEV void ld(u32* v, const uint4* p) { ld_el<NL>(v, p); }
  static NTT_DEV void st(uint4* p, const u32* v) { st_el<NL>(p, v); }
  static NTT_DEV void ld_tw(u32* w, const uint4* tab, u32 idx) { ld_table_n<NL>(w, tab, idx); }
  static NTT_DEV void mul_tw(u32* x, const u32* w, const Params& f) { mul_by_n<NL>(x, w, f); }
  static NTT_DEV void bfly(u32* a, u32* b, const Params& f, bool) { bfly_n<NL>(a, b, f); }
  static NTT_DEV void canon(u32* x, const Params& f) { canonicalize_n<NL>(x, f); }
  static NTT_DEV void canon_exact(u32* x, const Params& f) { canonicalize_n<NL>(x, f); }
  static NTT_DEV void st_tw(uint4* tab, u32 idx, const u32* w, const Args&) { st_el<NL>(tab + (NL / 4) * (size_t)idx, w); }
  // x <- x * v[idx] / R for a caller-supplied Montgomery-form vector (= the plain product)
  static NTT_DEV void mul_var(u32* x, const uint4* vec, u32 idx, const Params& f) {
    u32 v[NL], r[NL];
    ld_table_n<NL>(v, vec, idx);
    canonicalize_n<NL>(x, f);  // the product wants its left operand below p
    mont_mul_n<NL>(r, x, v, f);
#pragma unroll
    for (int k = 0; k < NL; ++k) x[k] = r[k];
  }
};
template <int NL>
using PassArgsN = PassArgsT<FieldParamsN<NL>, NL, NL>;

// 384-bit containers: four elements per thread (48 data registers).  FZ: the launch carries element-wise
// companions (in_scale / out_scale / out_mul).
template <bool FZ>
__global__ void __launch_bounds__(512, 1)
ntt384_pass_kernel(const uint4* data_in, uint4* data_out, const __grid_constant__ PassArgsN<12> A) {
  extern __shared__ u32 smem[];
  ntt_pass_body<2, FN<12>, FZ>(data_in, data_out, A, smem);
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
