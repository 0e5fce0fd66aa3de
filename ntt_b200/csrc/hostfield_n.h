// Host-side modular arithmetic on W x 64-bit limbs, used ONLY to derive constants for the wider
// fields (R^2, -p^-1, w_N, seed powers of the twiddle tables, 1/n): the W-limb generalisation of
// hostfield.h (W = 6: 384-bit containers).  Not an NTT: no transform ever runs on the host.
#pragma once
#include <cstdint>
#include <cstring>

namespace ntt256 {

template <int W>
struct UIntN {
  uint64_t w[W];
};

template <int W>
inline UIntN<W> un_from_limbs(const uint32_t* l) {
  UIntN<W> r;
  for (int i = 0; i < W; ++i) r.w[i] = (uint64_t)l[2 * i] | ((uint64_t)l[2 * i + 1] << 32);
  return r;
}
template <int W>
inline void un_to_limbs(const UIntN<W>& a, uint32_t* l) {
  for (int i = 0; i < W; ++i) {
    l[2 * i] = (uint32_t)a.w[i];
    l[2 * i + 1] = (uint32_t)(a.w[i] >> 32);
  }
}
template <int W>
inline UIntN<W> un_small(uint64_t v) {
  UIntN<W> r;
  std::memset(&r, 0, sizeof(r));
  r.w[0] = v;
  return r;
}
template <int W>
inline bool un_is_zero(const UIntN<W>& a) {
  uint64_t o = 0;
  for (int i = 0; i < W; ++i) o |= a.w[i];
  return o == 0;
}
template <int W>
inline int un_cmp(const UIntN<W>& a, const UIntN<W>& b) {
  for (int i = W - 1; i >= 0; --i) {
    if (a.w[i] > b.w[i]) return 1;
    if (a.w[i] < b.w[i]) return -1;
  }
  return 0;
}
template <int W>
inline uint64_t un_add(UIntN<W>& r, const UIntN<W>& a, const UIntN<W>& b) {
  unsigned __int128 c = 0;
  for (int i = 0; i < W; ++i) {
    c += (unsigned __int128)a.w[i] + b.w[i];
    r.w[i] = (uint64_t)c;
    c >>= 64;
  }
  return (uint64_t)c;
}
template <int W>
inline uint64_t un_sub(UIntN<W>& r, const UIntN<W>& a, const UIntN<W>& b) {
  uint64_t bw = 0;
  for (int i = 0; i < W; ++i) {
    unsigned __int128 t = (unsigned __int128)a.w[i] - b.w[i] - bw;
    r.w[i] = (uint64_t)t;
    bw = (uint64_t)(t >> 64) & 1;
  }
  return bw;
}
template <int W>
inline UIntN<W> un_shr(const UIntN<W>& a, int s) {
  UIntN<W> r;
  for (int i = 0; i < W; ++i) {
    int w = i + (s >> 6);
    uint64_t lo = w < W ? a.w[w] : 0, hi = (w + 1) < W ? a.w[w + 1] : 0;
    r.w[i] = (s & 63) ? (lo >> (s & 63)) | (hi << (64 - (s & 63))) : lo;
  }
  return r;
}
template <int W>
inline int un_bit(const UIntN<W>& a, int i) {
  return (int)((a.w[i >> 6] >> (i & 63)) & 1);
}
template <int W>
inline int un_bitlen(const UIntN<W>& a) {
  for (int i = 64 * W - 1; i >= 0; --i)
    if (un_bit(a, i)) return i + 1;
  return 0;
}

template <int W>
struct HostFieldN {
  typedef UIntN<W> U;
  U p, r2, one;
  uint64_t np0;

  void init(const U& prime) {
    p = prime;
    uint64_t inv = 1;
    for (int i = 0; i < 6; ++i) inv *= 2 - p.w[0] * inv;
    np0 = 0 - inv;
    U x = un_small<W>(1);
    for (int i = 0; i < 128 * W; ++i) {  // 2^(64W) = R, then R^2, by doubling mod p
      x = add(x, x);
      if (i == 64 * W - 1) one = x;
    }
    r2 = x;
  }
  U add(const U& a, const U& b) const {
    U t, d;
    uint64_t c = un_add(t, a, b);
    uint64_t bw = un_sub(d, t, p);
    return (c || !bw) ? d : t;
  }
  U sub(const U& a, const U& b) const {
    U t;
    if (un_sub(t, a, b)) un_add(t, t, p);
    return t;
  }
  // a*b/R mod p, canonical (CIOS on 64-bit limbs)
  U mul(const U& a, const U& b) const {
    uint64_t t[W + 2];
    std::memset(t, 0, sizeof(t));
    for (int i = 0; i < W; ++i) {
      unsigned __int128 c = 0;
      for (int j = 0; j < W; ++j) {
        c += (unsigned __int128)a.w[j] * b.w[i] + t[j];
        t[j] = (uint64_t)c;
        c >>= 64;
      }
      c += t[W];
      t[W] = (uint64_t)c;
      t[W + 1] = (uint64_t)(c >> 64);
      uint64_t m = t[0] * np0;
      c = (unsigned __int128)m * p.w[0] + t[0];
      c >>= 64;
      for (int j = 1; j < W; ++j) {
        c += (unsigned __int128)m * p.w[j] + t[j];
        t[j - 1] = (uint64_t)c;
        c >>= 64;
      }
      c += t[W];
      t[W - 1] = (uint64_t)c;
      t[W] = t[W + 1] + (uint64_t)(c >> 64);
    }
    U r, d;
    for (int i = 0; i < W; ++i) r.w[i] = t[i];
    uint64_t bw = un_sub(d, r, p);
    return (t[W] || !bw) ? d : r;
  }
  U to_mont(const U& a) const { return mul(a, r2); }
  U from_mont(const U& a) const { return mul(a, un_small<W>(1)); }
  U pow(const U& base_mont, const U& e) const {
    U acc = one, b = base_mont;
    for (int i = 0; i < 64 * W; ++i) {
      if (un_bit(e, i)) acc = mul(acc, b);
      b = mul(b, b);
    }
    return acc;
  }
  U inv(const U& a_mont) const {  // Fermat
    U e;
    un_sub(e, p, un_small<W>(2));
    return pow(a_mont, e);
  }
};

}  // namespace ntt256
