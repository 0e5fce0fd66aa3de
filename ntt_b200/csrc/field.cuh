// 256-bit prime-field arithmetic for sm_100a: one thread owns whole elements
// (8 x u32 little-endian limbs in registers), no inter-thread traffic.
//
// Replaces the reference's 8-threads-per-element CGBN path
// (/root/reference/src/big-num.cu:88-103 -> include/cgbn/core/core_mont.cu:29-77,
//  core_add_sub.cu:27-49) with single-thread carry chains:
//   * Montgomery product on mad.lo.cc / madc.hi.cc chains laid out "even/odd"
//     so that ptxas fuses every lo/hi pair into one IMAD.WIDE.U32(.X) -- 128 wide
//     multiply-accumulates + 8 low multiplies per product, zero shuffles (measured on
//     B200: 66 G products/s; the carry-predicate forms issue at half the plain IMAD.WIDE
//     rate, see profiles/r01_microbench_notes.md);
//   * data stay canonical (non-Montgomery) in memory, twiddles are kept in
//     Montgomery form, so  montmul(w*R, x) = w*x mod p  needs no bn2mont/mont2bn
//     (the reference converts 2+1 times per butterfly, big-num.cu:88-92);
//   * lazy reduction: values live in [0, 2p) between butterflies and are made
//     canonical only on the way out.  Requires p < 2^255 (true for BN254 Fr and
//     BLS12-381 Fr, and enforced by ntt256_ctx_create).
//
// tests/test_field_model.py holds a limb-exact Python model of mont_mul() below and
// fuzzes the "carry out of the odd chain is always zero" claim it relies on.
#pragma once
#include <cstdint>

namespace ntt256 {

typedef uint32_t u32;
typedef uint64_t u64;

// Field constants travel as a __grid_constant__ kernel parameter, i.e. they sit in
// the constant bank and are consumed directly as IMAD/IADD3 c[][] operands.
struct FieldParams {
  u32 p[8];     // modulus
  u32 p2[8];    // 2p (fits: p < 2^255)
  u32 pbar[8];  // 2^256 - p (shoup_mul, field_shoup.cuh)
  u32 p2bar[8]; // 2^256 - 2p (field_wl.cuh)
  u32 kp[8];    // 3p: added to a difference so that it cannot go negative (field_wl.cuh)
  u32 np0;      // -p^{-1} mod 2^32
  u32 pad[3];
};

struct Fe {
  u32 v[8];
};

#define NTT_DEV __device__ __forceinline__

// ---------------------------------------------------------------- carry chains
// acc[0..7] = (a0,a2,a4,a6)*b laid at limbs (0,1),(2,3),(4,5),(6,7).
// `a` points at the first of four limbs taken with stride 2.
NTT_DEV void mul_row(u32* acc, const u32* a, u32 b) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    // one plain IMAD.WIDE.U32 (full rate); mul.lo + mul.hi would be emitted as IMAD + IMAD.HI
    asm("{\n\t.reg .u64 t;\n\t"
        "mul.wide.u32 t, %2, %3;\n\t"
        "mov.b64 {%0, %1}, t;\n\t}"
        : "=r"(acc[2 * j]), "=r"(acc[2 * j + 1])
        : "r"(a[2 * j]), "r"(b));
  }
}

// acc[0..7] += (a0,a2,a4,a6)*b; returns the carry out of limb 7.
NTT_DEV u32 mad_row(u32* acc, const u32* a, u32 b) {
  u32 c;
  asm("mad.lo.cc.u32   %0, %9,  %13, %0;\n\t"
      "madc.hi.cc.u32  %1, %9,  %13, %1;\n\t"
      "madc.lo.cc.u32  %2, %10, %13, %2;\n\t"
      "madc.hi.cc.u32  %3, %10, %13, %3;\n\t"
      "madc.lo.cc.u32  %4, %11, %13, %4;\n\t"
      "madc.hi.cc.u32  %5, %11, %13, %5;\n\t"
      "madc.lo.cc.u32  %6, %12, %13, %6;\n\t"
      "madc.hi.cc.u32  %7, %12, %13, %7;\n\t"
      "addc.u32        %8, 0, 0;"
      : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]),
        "+r"(acc[6]), "+r"(acc[7]), "=r"(c)
      : "r"(a[0]), "r"(a[2]), "r"(a[4]), "r"(a[6]), "r"(b));
  return c;
}

// One CIOS row after the accumulator has been shifted right by one limb.
// On entry X (aligned to limb 0, X[0]==0 after reduction) and Y (aligned to limb 1)
// hold the previous row.  On exit
//     X <- Y + (a0,a2,a4,a6)*b   (+ X[1], the limb that has no slot in the new Y)
//     Y <- (X >> 64) + (a1,a3,a5,a7)*b
// and the carry out of the new X is folded into Y[7].
NTT_DEV void shift_mad_row(u32* X, u32* Y, const u32* a, u32 b) {
  u32 nX[8], nY[8];
  asm("add.cc.u32      %0, %9,  %10;\n\t"        // nX0 = Y0 + X1, carry -> limb 1 = nY0
      "madc.lo.cc.u32  %1, %11, %15, %16;\n\t"   // nY0 = lo(a1*b) + X2
      "madc.hi.cc.u32  %2, %11, %15, %17;\n\t"   // nY1 = hi(a1*b) + X3
      "madc.lo.cc.u32  %3, %12, %15, %18;\n\t"   // nY2 = lo(a3*b) + X4
      "madc.hi.cc.u32  %4, %12, %15, %19;\n\t"
      "madc.lo.cc.u32  %5, %13, %15, %20;\n\t"   // nY4 = lo(a5*b) + X6
      "madc.hi.cc.u32  %6, %13, %15, %21;\n\t"
      "madc.lo.cc.u32  %7, %14, %15, 0;\n\t"     // nY6 = lo(a7*b)
      "madc.hi.u32     %8, %14, %15, 0;"         // nY7 = hi(a7*b)
      : "=&r"(nX[0]), "=&r"(nY[0]), "=&r"(nY[1]), "=&r"(nY[2]), "=&r"(nY[3]), "=&r"(nY[4]),
        "=&r"(nY[5]), "=&r"(nY[6]), "=&r"(nY[7])
      : "r"(Y[0]), "r"(X[1]), "r"(a[1]), "r"(a[3]), "r"(a[5]), "r"(a[7]), "r"(b), "r"(X[2]),
        "r"(X[3]), "r"(X[4]), "r"(X[5]), "r"(X[6]), "r"(X[7]));
  asm("mad.lo.cc.u32   %0, %9,  %13, %0;\n\t"    // nX0 += lo(a0*b)
      "madc.hi.cc.u32  %1, %9,  %13, %14;\n\t"   // nX1 = hi(a0*b) + Y1
      "madc.lo.cc.u32  %2, %10, %13, %15;\n\t"
      "madc.hi.cc.u32  %3, %10, %13, %16;\n\t"
      "madc.lo.cc.u32  %4, %11, %13, %17;\n\t"
      "madc.hi.cc.u32  %5, %11, %13, %18;\n\t"
      "madc.lo.cc.u32  %6, %12, %13, %19;\n\t"
      "madc.hi.cc.u32  %7, %12, %13, %20;\n\t"
      "addc.u32        %8, %8, 0;"               // carry of limb 7 has the weight of nY7
      : "+r"(nX[0]), "=&r"(nX[1]), "=&r"(nX[2]), "=&r"(nX[3]), "=&r"(nX[4]), "=&r"(nX[5]),
        "=&r"(nX[6]), "=&r"(nX[7]), "+r"(nY[7])
      : "r"(a[0]), "r"(a[2]), "r"(a[4]), "r"(a[6]), "r"(b), "r"(Y[1]), "r"(Y[2]), "r"(Y[3]),
        "r"(Y[4]), "r"(Y[5]), "r"(Y[6]), "r"(Y[7]));
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    X[k] = nX[k];
    Y[k] = nY[k];
  }
}

// Montgomery reduction step of one row: m = X0 * np0; X += p_even*m (X0 -> 0);
// Y += p_odd*m.  The odd chain cannot carry out (a < p < 2^255, see the model test).
NTT_DEV void redc_row(u32* X, u32* Y, const FieldParams& f) {
  u32 m = X[0] * f.np0;
  u32 cx = mad_row(X, f.p, m);
  (void)mad_row(Y, f.p + 1, m);
  Y[7] += cx;
}

// r = a * b * 2^-256 mod p, result in [0, 2p).
// Preconditions: a < p (the twiddle, Montgomery form), b < 2^256 with a*b < p*2^256
// (b < 2p is what the butterflies feed).  Even/odd CIOS: 128 IMAD.WIDE + 8 IMAD.
NTT_DEV void mont_mul(u32* r, const u32* a, const u32* b, const FieldParams& f) {
  u32 X[8], Y[8];
  mul_row(X, a, b[0]);
  mul_row(Y, a + 1, b[0]);
  redc_row(X, Y, f);
#pragma unroll
  for (int i = 1; i < 8; ++i) {
    shift_mad_row(X, Y, a, b[i]);
    redc_row(X, Y, f);
  }
  // result = (X + Y*2^32) >> 32 ; X[0] == 0
  asm("add.cc.u32   %0, %8,  %16;\n\t"
      "addc.cc.u32  %1, %9,  %17;\n\t"
      "addc.cc.u32  %2, %10, %18;\n\t"
      "addc.cc.u32  %3, %11, %19;\n\t"
      "addc.cc.u32  %4, %12, %20;\n\t"
      "addc.cc.u32  %5, %13, %21;\n\t"
      "addc.cc.u32  %6, %14, %22;\n\t"
      "addc.u32     %7, %15, 0;"
      : "=&r"(r[0]), "=&r"(r[1]), "=&r"(r[2]), "=&r"(r[3]), "=&r"(r[4]), "=&r"(r[5]),
        "=&r"(r[6]), "=&r"(r[7])
      : "r"(Y[0]), "r"(Y[1]), "r"(Y[2]), "r"(Y[3]), "r"(Y[4]), "r"(Y[5]), "r"(Y[6]), "r"(Y[7]),
        "r"(X[1]), "r"(X[2]), "r"(X[3]), "r"(X[4]), "r"(X[5]), "r"(X[6]), "r"(X[7]));
}

// ---------------------------------------------------------------- add / sub
// r = a + b, returns carry out of limb 7.
NTT_DEV u32 add256(u32* r, const u32* a, const u32* b) {
  u32 c;
  asm("add.cc.u32   %0, %9,  %17;\n\t"
      "addc.cc.u32  %1, %10, %18;\n\t"
      "addc.cc.u32  %2, %11, %19;\n\t"
      "addc.cc.u32  %3, %12, %20;\n\t"
      "addc.cc.u32  %4, %13, %21;\n\t"
      "addc.cc.u32  %5, %14, %22;\n\t"
      "addc.cc.u32  %6, %15, %23;\n\t"
      "addc.cc.u32  %7, %16, %24;\n\t"
      "addc.u32     %8, 0, 0;"
      : "=&r"(r[0]), "=&r"(r[1]), "=&r"(r[2]), "=&r"(r[3]), "=&r"(r[4]), "=&r"(r[5]),
        "=&r"(r[6]), "=&r"(r[7]), "=r"(c)
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]),
        "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]));
  return c;
}

// r = a - b, returns 0 if no borrow, 0xffffffff if a < b.
NTT_DEV u32 sub256(u32* r, const u32* a, const u32* b) {
  u32 bw;
  asm("sub.cc.u32   %0, %9,  %17;\n\t"
      "subc.cc.u32  %1, %10, %18;\n\t"
      "subc.cc.u32  %2, %11, %19;\n\t"
      "subc.cc.u32  %3, %12, %20;\n\t"
      "subc.cc.u32  %4, %13, %21;\n\t"
      "subc.cc.u32  %5, %14, %22;\n\t"
      "subc.cc.u32  %6, %15, %23;\n\t"
      "subc.cc.u32  %7, %16, %24;\n\t"
      "subc.u32     %8, 0, 0;"
      : "=&r"(r[0]), "=&r"(r[1]), "=&r"(r[2]), "=&r"(r[3]), "=&r"(r[4]), "=&r"(r[5]),
        "=&r"(r[6]), "=&r"(r[7]), "=r"(bw)
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]),
        "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]));
  return bw;
}

// r = a + (b & mask)
NTT_DEV void add256_masked(u32* r, const u32* a, const u32* b, u32 mask) {
  u32 t[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) t[k] = b[k] & mask;
  (void)add256(r, a, t);
}

// Lazy butterfly pieces; inputs and outputs in [0, 2p).
// s = a + b  (mod 2p window): subtract 2p when the sum reached it.
NTT_DEV void add_lazy(u32* s, const u32* a, const u32* b, const FieldParams& f) {
  u32 t[8], d[8];
  u32 c = add256(t, a, b);            // 4p may exceed 2^256 (BLS12-381 Fr): keep the carry
  u32 bw = sub256(d, t, f.p2);        // d = t - 2p
  // take d when the true sum >= 2p, i.e. carry set or no borrow
  u32 keep_t = bw & (c - 1u);         // all-ones iff borrow and no carry
#pragma unroll
  for (int k = 0; k < 8; ++k) s[k] = (t[k] & keep_t) | (d[k] & ~keep_t);
}

// d = a - b  (mod 2p window): add 2p back on borrow.  The add-back is a predicated carry chain
// (8 instructions) instead of mask-then-add (16): all integer instructions share one issue port.
NTT_DEV void sub_lazy(u32* d, const u32* a, const u32* b, const FieldParams& f) {
  u32 bw = sub256(d, a, b);
  asm("{\n\t.reg .pred q;\n\t"
      "setp.ne.u32 q, %8, 0;\n\t"
      "@q add.cc.u32   %0, %0, %9;\n\t"
      "@q addc.cc.u32  %1, %1, %10;\n\t"
      "@q addc.cc.u32  %2, %2, %11;\n\t"
      "@q addc.cc.u32  %3, %3, %12;\n\t"
      "@q addc.cc.u32  %4, %4, %13;\n\t"
      "@q addc.cc.u32  %5, %5, %14;\n\t"
      "@q addc.cc.u32  %6, %6, %15;\n\t"
      "@q addc.u32     %7, %7, %16;\n\t}"
      : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3]), "+r"(d[4]), "+r"(d[5]), "+r"(d[6]), "+r"(d[7])
      : "r"(bw), "r"(f.p2[0]), "r"(f.p2[1]), "r"(f.p2[2]), "r"(f.p2[3]), "r"(f.p2[4]), "r"(f.p2[5]), "r"(f.p2[6]),
        "r"(f.p2[7]));
}

// [0,2p) -> [0,p)
NTT_DEV void canonicalize(u32* r, const FieldParams& f) {
  u32 d[8];
  u32 bw = sub256(d, r, f.p);
#pragma unroll
  for (int k = 0; k < 8; ++k) r[k] = (r[k] & bw) | (d[k] & ~bw);
}

}  // namespace ntt256
