// Prime-field arithmetic for NL x 32-bit limbs (NL even), the generalisation of field.cuh that the
// 384-bit transforms use (SURVEY.md par. 8f row 4; the reference's `bits` template parameter,
// /root/reference/src/big-num.cu:57,67,172).  Same design as the 256-bit code: one thread owns whole
// elements, even/odd CIOS Montgomery product on mad.lo.cc / madc.hi.cc chains that ptxas fuses into
// IMAD.WIDE.U32(.X), lazy [0, 2p) values, p < 2^(32*NL-1).
//
// field.cuh writes its chains as one asm statement per chain (fixed at 8 limbs); here every PTX
// instruction is its own `asm volatile` statement inside fully unrolled loops, and the carry flag
// lives across statements -- the PTX condition-code register is architectural state that only
// .cc/addc/subc/madc instructions touch, and the compiler itself never emits those.
// tests/test_field_model.py models the chains limb by limb for NL = 8 and NL = 12.
#pragma once
#include <cstdint>

#include "field.cuh"

namespace ntt256 {

template <int NL>
struct FieldParamsN {
  u32 p[NL];   // modulus
  u32 p2[NL];  // 2p (fits: p < 2^(32*NL-1))
  u32 np0;     // -p^{-1} mod 2^32
  u32 pad[3];
};

template <int NL>
struct FeN {
  u32 v[NL];
};

namespace fn {

// ---- single PTX instructions; the carry flag links consecutive calls
NTT_DEV void add_cc(u32& r, u32 a, u32 b) { asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); }
NTT_DEV void addc_cc(u32& r, u32 a, u32 b) { asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); }
NTT_DEV void addc(u32& r, u32 a, u32 b) { asm volatile("addc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); }
NTT_DEV void sub_cc(u32& r, u32 a, u32 b) { asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); }
NTT_DEV void subc_cc(u32& r, u32 a, u32 b) { asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); }
NTT_DEV void subc(u32& r, u32 a, u32 b) { asm volatile("subc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); }
NTT_DEV void mad_lo_cc(u32& r, u32 a, u32 b, u32 c) {
  asm volatile("mad.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
}
NTT_DEV void madc_lo_cc(u32& r, u32 a, u32 b, u32 c) {
  asm volatile("madc.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
}
NTT_DEV void madc_hi_cc(u32& r, u32 a, u32 b, u32 c) {
  asm volatile("madc.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
}
NTT_DEV void madc_hi(u32& r, u32 a, u32 b, u32 c) {
  asm volatile("madc.hi.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
}

}  // namespace fn

// acc[0..NL) = (a[0], a[2], ...) * b laid at limb pairs; `a` points at the first limb of the stride-2 set.
template <int NL>
NTT_DEV void mul_row_n(u32* acc, const u32* a, u32 b) {
#pragma unroll
  for (int j = 0; j < NL / 2; ++j) {
    asm("{\n\t.reg .u64 t;\n\t"
        "mul.wide.u32 t, %2, %3;\n\t"
        "mov.b64 {%0, %1}, t;\n\t}"
        : "=r"(acc[2 * j]), "=r"(acc[2 * j + 1])
        : "r"(a[2 * j]), "r"(b));
  }
}

// acc[0..NL) += (a[0], a[2], ...) * b; returns the carry out of the top limb.
template <int NL>
NTT_DEV u32 mad_row_n(u32* acc, const u32* a, u32 b) {
  u32 c;
  fn::mad_lo_cc(acc[0], a[0], b, acc[0]);
  fn::madc_hi_cc(acc[1], a[0], b, acc[1]);
#pragma unroll
  for (int j = 1; j < NL / 2; ++j) {
    fn::madc_lo_cc(acc[2 * j], a[2 * j], b, acc[2 * j]);
    fn::madc_hi_cc(acc[2 * j + 1], a[2 * j], b, acc[2 * j + 1]);
  }
  fn::addc(c, 0u, 0u);
  return c;
}

// One CIOS row after the accumulator moved down one limb (see shift_mad_row in field.cuh):
//   X <- Y + a_even * b (+ X[1]),  Y <- (X >> 64) + a_odd * b, even chain's carry folded into Y's top limb.
template <int NL>
NTT_DEV void shift_mad_row_n(u32* X, u32* Y, const u32* a, u32 b) {
  u32 nX[NL], nY[NL];
  fn::add_cc(nX[0], Y[0], X[1]);
#pragma unroll
  for (int j = 0; j < NL / 2 - 1; ++j) {
    fn::madc_lo_cc(nY[2 * j], a[2 * j + 1], b, X[2 * j + 2]);
    fn::madc_hi_cc(nY[2 * j + 1], a[2 * j + 1], b, X[2 * j + 3]);
  }
  fn::madc_lo_cc(nY[NL - 2], a[NL - 1], b, 0u);
  fn::madc_hi(nY[NL - 1], a[NL - 1], b, 0u);
  fn::mad_lo_cc(nX[0], a[0], b, nX[0]);
  fn::madc_hi_cc(nX[1], a[0], b, Y[1]);
#pragma unroll
  for (int j = 1; j < NL / 2; ++j) {
    fn::madc_lo_cc(nX[2 * j], a[2 * j], b, Y[2 * j]);
    fn::madc_hi_cc(nX[2 * j + 1], a[2 * j], b, Y[2 * j + 1]);
  }
  fn::addc(nY[NL - 1], nY[NL - 1], 0u);
#pragma unroll
  for (int k = 0; k < NL; ++k) {
    X[k] = nX[k];
    Y[k] = nY[k];
  }
}

template <int NL>
NTT_DEV void redc_row_n(u32* X, u32* Y, const FieldParamsN<NL>& f) {
  const u32 m = X[0] * f.np0;
  const u32 cx = mad_row_n<NL>(X, f.p, m);
  (void)mad_row_n<NL>(Y, f.p + 1, m);  // cannot carry out: p < 2^(32*NL-1)
  Y[NL - 1] += cx;
}

// r = a * b * 2^(-32*NL) mod p, result in [0, 2p).  a < p (Montgomery-form twiddle), b < 2^(32*NL).
template <int NL>
NTT_DEV void mont_mul_n(u32* r, const u32* a, const u32* b, const FieldParamsN<NL>& f) {
  u32 X[NL], Y[NL];
  mul_row_n<NL>(X, a, b[0]);
  mul_row_n<NL>(Y, a + 1, b[0]);
  redc_row_n<NL>(X, Y, f);
#pragma unroll
  for (int i = 1; i < NL; ++i) {
    shift_mad_row_n<NL>(X, Y, a, b[i]);
    redc_row_n<NL>(X, Y, f);
  }
  // result = (X + Y*2^32) >> 32 ; X[0] == 0
  fn::add_cc(r[0], Y[0], X[1]);
#pragma unroll
  for (int k = 1; k < NL - 1; ++k) fn::addc_cc(r[k], Y[k], X[k + 1]);
  fn::addc(r[NL - 1], Y[NL - 1], 0u);
}

// r = a + b, returns the carry out of the top limb.
template <int NL>
NTT_DEV u32 add_n(u32* r, const u32* a, const u32* b) {
  u32 c;
  fn::add_cc(r[0], a[0], b[0]);
#pragma unroll
  for (int k = 1; k < NL; ++k) fn::addc_cc(r[k], a[k], b[k]);
  fn::addc(c, 0u, 0u);
  return c;
}

// r = a - b, returns 0 if no borrow, 0xffffffff if a < b.
template <int NL>
NTT_DEV u32 sub_n(u32* r, const u32* a, const u32* b) {
  u32 bw;
  fn::sub_cc(r[0], a[0], b[0]);
#pragma unroll
  for (int k = 1; k < NL; ++k) fn::subc_cc(r[k], a[k], b[k]);
  fn::subc(bw, 0u, 0u);
  return bw;
}

// Lazy butterfly halves; inputs and outputs in [0, 2p).
template <int NL>
NTT_DEV void add_lazy_n(u32* s, const u32* a, const u32* b, const FieldParamsN<NL>& f) {
  u32 t[NL], d[NL];
  const u32 c = add_n<NL>(t, a, b);  // 4p may exceed 2^(32*NL): keep the carry
  const u32 bw = sub_n<NL>(d, t, f.p2);
  const u32 keep_t = bw & (c - 1u);  // all-ones iff borrow and no carry: the sum is below 2p
#pragma unroll
  for (int k = 0; k < NL; ++k) s[k] = (t[k] & keep_t) | (d[k] & ~keep_t);
}

template <int NL>
NTT_DEV void sub_lazy_n(u32* d, const u32* a, const u32* b, const FieldParamsN<NL>& f) {
  u32 t[NL], u[NL];
  const u32 bw = sub_n<NL>(t, a, b);
  (void)add_n<NL>(u, t, f.p2);
#pragma unroll
  for (int k = 0; k < NL; ++k) d[k] = (u[k] & bw) | (t[k] & ~bw);
}

// [0, 2p) -> [0, p)
template <int NL>
NTT_DEV void canonicalize_n(u32* r, const FieldParamsN<NL>& f) {
  u32 d[NL];
  const u32 bw = sub_n<NL>(d, r, f.p);
#pragma unroll
  for (int k = 0; k < NL; ++k) r[k] = (r[k] & bw) | (d[k] & ~bw);
}

}  // namespace ntt256
