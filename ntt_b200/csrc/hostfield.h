// Host-side 256-bit modular arithmetic used ONLY to derive constants (R^2, -p^-1, w_N, the
// seed powers w^(2^k) of the twiddle tables, 1/n).  It replaces the GMP host path of CGBN that
// the reference uses for the same purpose (/root/reference/include/cgbn/impl_mpz.cc:851-859 via
// big-num.cu:292-306).  It is not an NTT: no transform ever runs on the host.
#pragma once
#include <cstdint>
#include <cstring>

namespace ntt256 {

struct U256 {
  uint64_t w[4];
};

inline U256 u256_from_limbs(const uint32_t l[8]) {
  U256 r;
  for (int i = 0; i < 4; ++i) r.w[i] = (uint64_t)l[2 * i] | ((uint64_t)l[2 * i + 1] << 32);
  return r;
}
inline void u256_to_limbs(const U256& a, uint32_t l[8]) {
  for (int i = 0; i < 4; ++i) {
    l[2 * i] = (uint32_t)a.w[i];
    l[2 * i + 1] = (uint32_t)(a.w[i] >> 32);
  }
}
inline U256 u256_small(uint64_t v) { return U256{{v, 0, 0, 0}}; }
inline bool u256_is_zero(const U256& a) { return !(a.w[0] | a.w[1] | a.w[2] | a.w[3]); }
inline int u256_cmp(const U256& a, const U256& b) {
  for (int i = 3; i >= 0; --i) {
    if (a.w[i] > b.w[i]) return 1;
    if (a.w[i] < b.w[i]) return -1;
  }
  return 0;
}
inline uint64_t u256_add(U256& r, const U256& a, const U256& b) {
  unsigned __int128 c = 0;
  for (int i = 0; i < 4; ++i) {
    c += (unsigned __int128)a.w[i] + b.w[i];
    r.w[i] = (uint64_t)c;
    c >>= 64;
  }
  return (uint64_t)c;
}
inline uint64_t u256_sub(U256& r, const U256& a, const U256& b) {
  uint64_t bw = 0;
  for (int i = 0; i < 4; ++i) {
    unsigned __int128 t = (unsigned __int128)a.w[i] - b.w[i] - bw;
    r.w[i] = (uint64_t)t;
    bw = (uint64_t)(t >> 64) & 1;
  }
  return bw;
}
inline U256 u256_shr(const U256& a, int s) {
  U256 r;
  for (int i = 0; i < 4; ++i) {
    int w = i + (s >> 6);
    uint64_t lo = w < 4 ? a.w[w] : 0, hi = (w + 1) < 4 ? a.w[w + 1] : 0;
    r.w[i] = (s & 63) ? (lo >> (s & 63)) | (hi << (64 - (s & 63))) : lo;
  }
  return r;
}
inline int u256_bit(const U256& a, int i) { return (int)((a.w[i >> 6] >> (i & 63)) & 1); }
inline int u256_bitlen(const U256& a) {
  for (int i = 255; i >= 0; --i)
    if (u256_bit(a, i)) return i + 1;
  return 0;
}

// low 256 bits of a*b
inline U256 u256_mul_lo(const U256& a, const U256& b) {
  uint64_t t[4] = {0, 0, 0, 0};
  for (int i = 0; i < 4; ++i) {
    unsigned __int128 c = 0;
    for (int j = 0; i + j < 4; ++j) {
      c += (unsigned __int128)a.w[i] * b.w[j] + t[i + j];
      t[i + j] = (uint64_t)c;
      c >>= 64;
    }
  }
  return U256{{t[0], t[1], t[2], t[3]}};
}

struct HostField {
  U256 p, r2, one;
  U256 pinv;  // p^-1 mod 2^256 (Shoup companions, field_shoup.cuh)
  uint64_t np0;

  void init(const U256& prime) {
    p = prime;
    uint64_t inv = 1;
    for (int i = 0; i < 6; ++i) inv *= 2 - p.w[0] * inv;
    np0 = 0 - inv;
    // Newton: y <- y*(2 - p*y) doubles the number of correct low bits (64 -> 128 -> 256)
    pinv = u256_small(inv);
    for (int i = 0; i < 2; ++i) {
      U256 t = u256_mul_lo(p, pinv), two_minus;
      u256_sub(two_minus, u256_small(2), t);
      pinv = u256_mul_lo(pinv, two_minus);
    }
    U256 x = u256_small(1);
    for (int i = 0; i < 512; ++i) {
      x = add(x, x);
      if (i == 255) one = x;
    }
    r2 = x;
  }
  U256 add(const U256& a, const U256& b) const {
    U256 t, d;
    uint64_t c = u256_add(t, a, b);
    uint64_t bw = u256_sub(d, t, p);
    return (c || !bw) ? d : t;
  }
  U256 sub(const U256& a, const U256& b) const {
    U256 t;
    if (u256_sub(t, a, b)) u256_add(t, t, p);
    return t;
  }
  // a*b/R mod p, canonical
  U256 mul(const U256& a, const U256& b) const {
    uint64_t t[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; ++i) {
      unsigned __int128 c = 0;
      for (int j = 0; j < 4; ++j) {
        c += (unsigned __int128)a.w[j] * b.w[i] + t[j];
        t[j] = (uint64_t)c;
        c >>= 64;
      }
      c += t[4];
      t[4] = (uint64_t)c;
      t[5] = (uint64_t)(c >> 64);
      uint64_t m = t[0] * np0;
      c = (unsigned __int128)m * p.w[0] + t[0];
      c >>= 64;
      for (int j = 1; j < 4; ++j) {
        c += (unsigned __int128)m * p.w[j] + t[j];
        t[j - 1] = (uint64_t)c;
        c >>= 64;
      }
      c += t[4];
      t[3] = (uint64_t)c;
      t[4] = t[5] + (uint64_t)(c >> 64);
    }
    U256 r{{t[0], t[1], t[2], t[3]}}, d;
    uint64_t bw = u256_sub(d, r, p);
    return (t[4] || !bw) ? d : r;
  }
  U256 to_mont(const U256& a) const { return mul(a, r2); }
  // floor(w * 2^256 / p) for a plain canonical w: (w*2^256 - (w*2^256 mod p)) / p, exact, as a product with p^-1
  U256 shoup_companion(const U256& w_plain) const {
    U256 wm = to_mont(w_plain), neg;
    u256_sub(neg, u256_small(0), wm);
    return u256_mul_lo(neg, pinv);
  }
  U256 from_mont(const U256& a) const { return mul(a, u256_small(1)); }
  U256 pow(const U256& base_mont, const U256& e) const {
    U256 acc = one, b = base_mont;
    for (int i = 0; i < 256; ++i) {
      if (u256_bit(e, i)) acc = mul(acc, b);
      b = mul(b, b);
    }
    return acc;
  }
  U256 inv(const U256& a_mont) const {  // Fermat: a^(p-2), cf. inv() at big-num.cu:27-29
    U256 e;
    u256_sub(e, p, u256_small(2));
    return pow(a_mont, e);
  }
};

}  // namespace ntt256
