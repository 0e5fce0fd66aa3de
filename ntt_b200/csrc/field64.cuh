// Field layer of the 64-bit transforms (include/ntt64.h): one element = one u64, canonical in [0, p).
//
// The reference's kernels are templates over the container width (`template <uint tpi, uint bits>`,
// /root/reference/src/big-num.cu:57,67,172,260) and its CPU NTT runs in `long long` containers
// (big-num.cu:37-55, p = 469762049).  Two instances of that width live here:
//
//   GL = true   the Goldilocks prime p = 2^64 - 2^32 + 1: the product is one 64x64->128 multiply (four
//               IMAD.WIDE) and a reduction by 2^64 = 2^32 - 1, 2^96 = -1 (mod p): about a dozen adds, no second
//               multiply.  Twiddles are plain canonical values.
//   GL = false  any odd p < 2^63 (e.g. the reference's own 469762049): Montgomery product with R = 2^64,
//               twiddles in Montgomery form, data plain -- redc(wR * x) = w * x, no conversions.
//
// Every function takes and returns canonical values; with one u64 per element the conditional corrections
// cost two predicated instructions, which is less than tracking lazy ranges would save.
#pragma once
#include <cstdint>

#ifndef NTT_DEV
#define NTT_DEV __device__ __forceinline__
#endif

namespace ntt64 {

using u32 = unsigned int;
using u64 = unsigned long long;

constexpr u64 kGoldilocksP = 0xFFFFFFFF00000001ull;
constexpr u32 kEps = 0xFFFFFFFFu;  // 2^64 mod p for Goldilocks

struct Params64 {
  u64 p;
  u64 np;  // -p^-1 mod 2^64 (Montgomery instance only)
};

NTT_DEV u64 pack(u32 lo, u32 hi) { return ((u64)hi << 32) | lo; }

// ---------------------------------------------------------------- Goldilocks
// d = a - b mod p (a, b canonical): subtract, add p back on borrow (p = -EPS mod 2^64, so "subtract EPS")
NTT_DEV u64 gl_sub(u64 a, u64 b) {
  u32 d0, d1, m;
  asm("{\n\t"
      "sub.cc.u32  %0, %3, %5;\n\t"
      "subc.cc.u32 %1, %4, %6;\n\t"
      "subc.u32    %2, 0, 0;\n\t"   // 0 or 0xffffffff (= EPS)
      "sub.cc.u32  %0, %0, %2;\n\t"
      "subc.u32    %1, %1, 0;\n\t}"
      : "=&r"(d0), "=&r"(d1), "=&r"(m)
      : "r"((u32)a), "r"((u32)(a >> 32)), "r"((u32)b), "r"((u32)(b >> 32)));
  return pack(d0, d1);
}
// a + b mod p = a - (p - b); p - 0 = p is not canonical but a - p borrows for every a < p and comes back as a
NTT_DEV u64 gl_add(u64 a, u64 b) { return gl_sub(a, kGoldilocksP - b); }

// t + x2 * EPS for any 64-bit t = (t1:t0), made canonical: one wide multiply-add; on carry add EPS (= subtract
// p), which cannot carry again because x2 * EPS <= 2^64 - 2^33 + 1
NTT_DEV u64 gl_reduce96(u32 t0, u32 t1, u32 x2) {
  u32 r0, r1, c;
  asm("{\n\t"
      "mad.lo.cc.u32  %0, %5, %6, %3;\n\t"
      "madc.hi.cc.u32 %1, %5, %6, %4;\n\t"
      "addc.u32       %2, 0, 0;\n\t}"
      : "=&r"(r0), "=&r"(r1), "=&r"(c)
      : "r"(t0), "r"(t1), "r"(x2), "r"(kEps));
  const u32 e = 0u - c;  // EPS when the sum carried
  asm("{\n\t"
      "add.cc.u32 %0, %0, %2;\n\t"
      "addc.u32   %1, %1, 0;\n\t}"
      : "+r"(r0), "+r"(r1)
      : "r"(e));
  // canonical: r >= p  <=>  r1 = 0xffffffff and r0 != 0, and then r - p = (r0 - 1, 0)
  const bool ge = (r1 == kEps) && (r0 != 0u);
  r0 = ge ? r0 - 1u : r0;
  r1 = ge ? 0u : r1;
  return pack(r0, r1);
}

// x0 + x1 2^32 + x2 2^64 + x3 2^96 = (x1:x0) - x3 + x2 * EPS  (mod p), canonical
NTT_DEV u64 gl_reduce128(u32 x0, u32 x1, u32 x2, u32 x3) {
  u32 t0, t1, m;
  // t = (x1:x0) - x3, plus p on borrow: any u64, congruent
  asm("{\n\t"
      "sub.cc.u32  %0, %3, %5;\n\t"
      "subc.cc.u32 %1, %4, 0;\n\t"
      "subc.u32    %2, 0, 0;\n\t"
      "sub.cc.u32  %0, %0, %2;\n\t"
      "subc.u32    %1, %1, 0;\n\t}"
      : "=&r"(t0), "=&r"(t1), "=&r"(m)
      : "r"(x0), "r"(x1), "r"(x3));
  return gl_reduce96(t0, t1, x2);
}

NTT_DEV u64 gl_mul(u64 a, u64 b) {
  // the 128-bit type gives the shortest product ptxas makes of it (three IMAD.WIDE, one IMAD.WIDE.X, one IMAD.HI)
  const unsigned __int128 prod = (unsigned __int128)a * b;
  const u64 lo = (u64)prod, hi = (u64)(prod >> 64);
  return gl_reduce128((u32)lo, (u32)(lo >> 32), (u32)hi, (u32)(hi >> 32));
}

// x * 2^S mod p for a canonical x, 0 < S < 96, S not a multiple of 32: shifts and the reduction, no multiplier.
// Every 64th root of unity of this field is a power of two (2 has order 192), so the constant twiddles inside a
// register round -- powers of w_16 -- are such products (up to the sign 2^96 = -1).
template <int S>
NTT_DEV u64 gl_mul_pow2(u64 x) {
  static_assert(S > 0 && S < 96 && S % 32 != 0, "shift out of range");
  constexpr int q = S / 32, r = S % 32;
  const u32 x0 = (u32)x, x1 = (u32)(x >> 32);
  const u32 z0 = x0 << r, z1 = __funnelshift_l(x0, x1, r), z2 = x1 >> (32 - r);  // x * 2^r = z0 + z1 2^32 + z2 2^64
  if constexpr (q == 0) {
    return gl_reduce96(z0, z1, z2);
  } else if constexpr (q == 1) {
    return gl_reduce128(0u, z0, z1, z2);
  } else {
    // z0 2^64 + z1 2^96 + z2 2^128 = z0 * EPS - (z2:z1); both canonical (z0 * EPS <= 2^64 - 2^33 + 1, z2 < 2^31)
    const u64 ze = ((u64)z0 << 32) - z0;
    return gl_sub(ze, pack(z1, z2));
  }
}
// (a - b) * 2^S192 for an exponent modulo 192: 2^96 = -1 turns the upper half into (b - a) * 2^(S192 - 96)
template <int S192>
NTT_DEV u64 gl_sub_mul_pow2(u64 a, u64 b) {
  constexpr bool neg = S192 >= 96;
  const u64 d = neg ? gl_sub(b, a) : gl_sub(a, b);
  return gl_mul_pow2<S192 % 96>(d);
}
// (a - b) * w_16^e, e = 1 .. 7, when w_16 = 2^(12 K16) (K16 odd; 13 for the usual generator 7, 3 for its inverse)
template <int K16>
NTT_DEV u64 gl_sub_mul_w16(u64 a, u64 b, int e) {
  switch (e) {
    case 1: return gl_sub_mul_pow2<(12 * K16 * 1) % 192>(a, b);
    case 2: return gl_sub_mul_pow2<(12 * K16 * 2) % 192>(a, b);
    case 3: return gl_sub_mul_pow2<(12 * K16 * 3) % 192>(a, b);
    case 4: return gl_sub_mul_pow2<(12 * K16 * 4) % 192>(a, b);
    case 5: return gl_sub_mul_pow2<(12 * K16 * 5) % 192>(a, b);
    case 6: return gl_sub_mul_pow2<(12 * K16 * 6) % 192>(a, b);
    default: return gl_sub_mul_pow2<(12 * K16 * 7) % 192>(a, b);
  }
}

// ---------------------------------------------------------------- generic Montgomery, p < 2^63
NTT_DEV u64 m64_add(u64 a, u64 b, const Params64& P) {
  const u64 s = a + b;  // < 2^64
  return s >= P.p ? s - P.p : s;
}
NTT_DEV u64 m64_sub(u64 a, u64 b, const Params64& P) {
  const u64 d = a - b;
  return a < b ? d + P.p : d;
}
// a * b / 2^64 mod p
NTT_DEV u64 m64_mul(u64 a, u64 b, const Params64& P) {
  const u64 lo = a * b;
  const u64 hi = __umul64hi(a, b);
  const u64 q = lo * P.np;
  const u64 h2 = __umul64hi(q, P.p);
  // lo + low64(q * p) = 0 mod 2^64: it carries exactly when lo != 0
  const u64 t = hi + h2 + (lo != 0ull ? 1ull : 0ull);
  return t >= P.p ? t - P.p : t;
}

// ---------------------------------------------------------------- the policy the kernels are instantiated with
template <bool GL>
struct F64 {
  static NTT_DEV u64 add(u64 a, u64 b, const Params64& P) {
    if constexpr (GL) return gl_add(a, b); else return m64_add(a, b, P);
  }
  static NTT_DEV u64 sub(u64 a, u64 b, const Params64& P) {
    if constexpr (GL) return gl_sub(a, b); else return m64_sub(a, b, P);
  }
  // twiddle product: w from a table (plain for GL, Montgomery form otherwise), x plain
  static NTT_DEV u64 mul(u64 w, u64 x, const Params64& P) {
    if constexpr (GL) return gl_mul(w, x); else return m64_mul(w, x, P);
  }
};

}  // namespace ntt64
