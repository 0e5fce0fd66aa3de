// Host-only check of the constant-derivation arithmetic (ntt_b200/csrc/hostfield.h, hostfield_n.h):
// the W-limb generic field at W = 4 must agree with the 256-bit one, and satisfy the field identities
// at W = 6 (BLS12-377 Fq).  Prints one line per check; exit code 0 only if all pass.
#include <cstdio>
#include <cstdlib>

#include "../../ntt_b200/csrc/hostfield.h"
#include "../../ntt_b200/csrc/hostfield_n.h"
using namespace ntt256;

static int fails = 0;
static void check(bool ok, const char* what) {
  std::printf("%s %s\n", ok ? "ok  " : "FAIL", what);
  if (!ok) ++fails;
}

template <int W>
static UIntN<W> parse_hex(const char* hex) {  // big-endian hex of exactly 16*W digits
  UIntN<W> r;
  for (int i = 0; i < W; ++i) {
    char buf[17] = {0};
    for (int k = 0; k < 16; ++k) buf[k] = hex[16 * (W - 1 - i) + k];
    r.w[i] = std::strtoull(buf, nullptr, 16);
  }
  return r;
}

int main() {
  const uint32_t bn[8] = {0xf0000001u, 0x43e1f593u, 0x79b97091u, 0x2833e848u, 0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
  HostField h;
  h.init(u256_from_limbs(bn));
  HostFieldN<4> g;
  g.init(un_from_limbs<4>(bn));
  bool same = h.np0 == g.np0;
  for (int i = 0; i < 4; ++i) same = same && h.r2.w[i] == g.r2.w[i] && h.one.w[i] == g.one.w[i];
  check(same, "W=4 constants (np0, R, R^2) equal the 256-bit host field's");
  bool agree = true;
  uint64_t s = 88172645463325252ull;
  for (int it = 0; it < 200; ++it) {
    U256 a, b;
    UIntN<4> an, bn4;
    for (int i = 0; i < 4; ++i) {
      s ^= s << 13; s ^= s >> 7; s ^= s << 17;
      a.w[i] = an.w[i] = (i == 3) ? (s >> 4) : s;  // < 2^252 < p
      s ^= s << 13; s ^= s >> 7; s ^= s << 17;
      b.w[i] = bn4.w[i] = (i == 3) ? (s >> 4) : s;
    }
    U256 m = h.mul(a, b), ad = h.add(a, b), sb = h.sub(a, b);
    UIntN<4> mn = g.mul(an, bn4), adn = g.add(an, bn4), sbn = g.sub(an, bn4);
    for (int i = 0; i < 4; ++i) agree = agree && m.w[i] == mn.w[i] && ad.w[i] == adn.w[i] && sb.w[i] == sbn.w[i];
  }
  check(agree, "W=4 mul/add/sub agree with the 256-bit host field on 200 random pairs");

  HostFieldN<6> q;
  q.init(parse_hex<6>("01ae3a4617c510eac63b05c06ca1493b1a22d9f300f5138f1ef3622fba094800170b5d44300000008508c00000000001"));
  check(q.np0 == 0x8508bfffffffffffull, "W=6 -p^-1 mod 2^64 of BLS12-377 Fq");
  UIntN<6> x = q.to_mont(un_small<6>(987654321)), xi = q.inv(x);
  check(un_cmp(q.mul(x, xi), q.one) == 0, "W=6 x * x^-1 == 1");
  UIntN<6> back = q.from_mont(x);
  check(back.w[0] == 987654321 && back.w[1] == 0 && back.w[5] == 0, "W=6 from_mont(to_mont(v)) == v");
  // generator 15: 15^((p-1)/2) == -1, and w = 15^((p-1)/2^46) has order 2^46
  UIntN<6> pm1;
  un_sub(pm1, q.p, un_small<6>(1));
  UIntN<6> gm = q.to_mont(un_small<6>(15));
  UIntN<6> minus1 = q.sub(un_small<6>(0), q.one);
  check(un_cmp(q.pow(gm, un_shr(pm1, 1)), minus1) == 0, "W=6 15 is a quadratic non-residue of BLS12-377 Fq");
  UIntN<6> w = q.pow(gm, un_shr(pm1, 46));
  for (int i = 0; i < 45; ++i) w = q.mul(w, w);
  check(un_cmp(w, minus1) == 0, "W=6 w_{2^46}^(2^45) == -1");
  // Shoup companions (field_shoup.cuh): p * pinv == 1 mod 2^256, and wq = floor(w 2^256 / p), i.e.
  // wq*p <= w*2^256 < (wq+1)*p, checked through the low halves: w*2^256 - wq*p == (w*2^256 mod p) < p.
  {
    U256 one256 = u256_mul_lo(h.p, h.pinv);
    check(one256.w[0] == 1 && !one256.w[1] && !one256.w[2] && !one256.w[3], "p * pinv == 1 (mod 2^256)");
    bool ok = true;
    uint64_t s2 = 0x9e3779b97f4a7c15ull;
    for (int it = 0; it < 200 && ok; ++it) {
      U256 w;
      for (int i = 0; i < 4; ++i) { s2 ^= s2 << 13; s2 ^= s2 >> 7; s2 ^= s2 << 17; w.w[i] = s2; }
      w.w[3] &= 0x0fffffffffffffffull;  // < p
      if (it == 0) w = u256_small(0);
      if (it == 1) w = u256_small(1);
      if (it == 2) u256_sub(w, h.p, u256_small(1));
      U256 wq = h.shoup_companion(w);
      // low 256 bits of (0 - wq*p) must equal w*2^256 mod p = to_mont(w), which is < p
      U256 lo = u256_mul_lo(wq, h.p), rem;
      u256_sub(rem, u256_small(0), lo);
      ok = ok && u256_cmp(rem, h.to_mont(w)) == 0 && u256_cmp(rem, h.p) < 0;
    }
    check(ok, "shoup_companion(w) == floor(w 2^256 / p) on 200 values incl. 0, 1, p-1");
  }
  std::printf("%s\n", fails ? "HOSTFIELD_FAIL" : "HOSTFIELD_OK");
  return fails ? 1 : 0;
}
