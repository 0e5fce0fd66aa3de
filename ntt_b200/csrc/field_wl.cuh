// "Wide-lazy" add / sub for fields with 2^229 < p and 5p + 2^226 < 2^256 (BN254 Fr: 2^256 / p = 5.29), the policy of
// the kernels instantiated with Shoup twiddles (SURVEY.md par. 7: "BN254 has 4p < 2^256, BLS12-381 only 2p").
//
// The reference reduces fully after every operation (CGBN add/sub + compare, /root/reference/src/big-num.cu:94-103);
// field.cuh keeps values in [0, 2p) with an EXACT conditional correction per operation (trial subtraction +
// select: 27 instructions per sum, 18 per difference, 17 per product).  Here the spare range above 4p and the
// fact that shoup_mul accepts ANY 256-bit operand are spent on cheaper corrections:
//
//   range      every value handed from one butterfly to the next lies in [0, 2p + e) with e <= 2^224;
//   compare    a correction only looks at the top 64 bits:  hi64(x) > hi64(2p)  implies x >= 2p, and its
//              negation implies x < 2p + 2^192 -- two ISETP instead of an 8-limb trial subtraction;
//   sum        s = a + b (< 4p + 2e, no overflow), then "if hi64(s) > hi64(2p): s -= 2p".  The excess over
//              2p of the result is at most max(2^192, e_a + e_b): it can double per butterfly stage and
//              restarts at 2^192 after every product, so e <= 2^192 * 2^32 over any transform (18 instr.);
//   difference feeding a product: d = a - b + 3p, unconditional: positive because b < 2p + 2^224 < 3p, and
//              d < 5p + 2^224 fits 256 bits; shoup_mul reduces it (16 instr.);
//   difference NOT feeding a product (trivial twiddle inside a round): d = a - b, "+ 3p on borrow", then the
//              top-64-bit correction: when 3p was added d lies in (p - 2^224, 3p) and the correction brings
//              it below p or leaves it below 2p + 2^192 (28 instr.);
//   product    shoup_mul ends with the same top-64-bit test (field_shoup.cuh): r in [0, 2p + 2^192);
//   way out    canonicalize_wl: one top-64-bit correction, then exact "- 2p if >= 2p" and "- p if >= p".
//
// tests/test_field_model.py holds the limb-exact Python model of these functions, drives it with the extreme
// values of every range and runs whole transforms through it with the excess tracked.
#pragma once
#include "field.cuh"

namespace ntt256 {

// x += c (mod 2^256) when the top 64 bits of x exceed those of 2p; c = 2^256 - 2p (or 2^256 - p).
NTT_DEV void add_if_above_2p(u32* x, const u32* c, const FieldParams& f) {
  asm("{\n\t.reg .pred q;\n\t.reg .u64 a, b;\n\t"
      "mov.b64 a, {%6, %7};\n\t"
      "mov.b64 b, {%16, %17};\n\t"
      "setp.gt.u64 q, a, b;\n\t"
      "@q add.cc.u32   %0, %0, %8;\n\t"
      "@q addc.cc.u32  %1, %1, %9;\n\t"
      "@q addc.cc.u32  %2, %2, %10;\n\t"
      "@q addc.cc.u32  %3, %3, %11;\n\t"
      "@q addc.cc.u32  %4, %4, %12;\n\t"
      "@q addc.cc.u32  %5, %5, %13;\n\t"
      "@q addc.cc.u32  %6, %6, %14;\n\t"
      "@q addc.u32     %7, %7, %15;\n\t}"
      : "+r"(x[0]), "+r"(x[1]), "+r"(x[2]), "+r"(x[3]), "+r"(x[4]), "+r"(x[5]), "+r"(x[6]), "+r"(x[7])
      : "r"(c[0]), "r"(c[1]), "r"(c[2]), "r"(c[3]), "r"(c[4]), "r"(c[5]), "r"(c[6]), "r"(c[7]), "r"(f.p2[6]),
        "r"(f.p2[7]));
}

// s = a + b, brought back below 2p + max(2^192, e_a + e_b)
NTT_DEV void add_wl(u32* s, const u32* a, const u32* b, const FieldParams& f) {
  (void)add256(s, a, b);
  add_if_above_2p(s, f.p2bar, f);
}

// d = a - b + 3p: for differences that a product reduces next
NTT_DEV void sub_wl_mul(u32* d, const u32* a, const u32* b, const FieldParams& f) {
  u32 t[8];
  (void)sub256(t, a, b);  // may wrap; the sum below is exact modulo 2^256 and the true value is in (0, 2^256)
  (void)add256(d, t, f.kp);
}

// d = a - b brought into [0, 2p + max(2^192, e_a)) without a product
NTT_DEV void sub_wl(u32* d, const u32* a, const u32* b, const FieldParams& f) {
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
      : "r"(bw), "r"(f.kp[0]), "r"(f.kp[1]), "r"(f.kp[2]), "r"(f.kp[3]), "r"(f.kp[4]), "r"(f.kp[5]), "r"(f.kp[6]),
        "r"(f.kp[7]));
  add_if_above_2p(d, f.p2bar, f);
}

// any value below 6p -> [0, p)
NTT_DEV void canonicalize_wl(u32* r, const FieldParams& f) {
  add_if_above_2p(r, f.p2bar, f);  // now below 4p
  u32 t[8];
  u32 c = add256(t, r, f.p2bar);   // carry <=> r >= 2p
  u32 m = 0u - c;
#pragma unroll
  for (int k = 0; k < 8; ++k) r[k] = (t[k] & m) | (r[k] & ~m);
  c = add256(t, r, f.pbar);        // carry <=> r >= p
  m = 0u - c;
#pragma unroll
  for (int k = 0; k < 8; ++k) r[k] = (t[k] & m) | (r[k] & ~m);
}

}  // namespace ntt256
