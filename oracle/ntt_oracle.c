/* TEST INFRASTRUCTURE ONLY -- the CPU oracle.  Never linked into or called by the product
 * path (ntt_b200/, libntt256.so).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library.
 *
 * Plain-C restatement of the reference's CPU NTT over a 256-bit prime field:
 *   - oracle_ntt256():      NTT() at /root/reference/src/big-num.cu:37-55 (bit-reverse permute,
 *                           then radix-2 DIT, gap = omega^((P-1)/(2*stride)), w stepped by gap),
 *                           omega_n rule from big-num.cu:128-134, inverse per the commented intent
 *                           at big-num.cu:470-477 (run with inv(root), scale by inv(length)).
 *   - oracle_ntt256_ssip(): NTT_pro1 + NTT_pro2 at /root/reference/src/self-sort-in-place.cu:79-128
 *                           (self-sort-in-place ordering; second, independent statement).
 * Memory image = cgbn_mem_t<256> (include/cgbn/cgbn_cuda.h:51-55): 8 little-endian u32 limbs,
 * canonical residues.  Internally 4 x u64 limbs with unsigned __int128 Montgomery products;
 * all values written back are canonical, so results equal the reference's `%` arithmetic.
 *
 * Parity status: pinned for p = 469762049, g = 3 (the reference's only field) by
 * tests/golden/ref_i64_kat.json and oracle/_ref; BN254 Fr / BLS12-381 Fr are pinned by the outputs
 * of the reference's own GPU drivers run on those fields (tests/golden/ref256_gpu_kat.json, made by
 * scripts/gen_golden_ref256.py from oracle/_ref/libref256.so) and anchored by
 * tests/golden/anchors.json + py_oracle.py.
 *
 * Build: gcc -O3 -march=native -fopenmp -shared -fPIC -o liboracle.so ntt_oracle.c
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef unsigned __int128 u128;
typedef uint64_t u64;

typedef struct {
  u64 p[4];
  u64 r2[4];  /* R^2 mod p */
  u64 one[4]; /* R mod p */
  u64 np0;    /* -p^-1 mod 2^64 */
} field_t;

static int geq(const u64 a[4], const u64 b[4]) {
  for (int i = 3; i >= 0; --i) {
    if (a[i] > b[i]) return 1;
    if (a[i] < b[i]) return 0;
  }
  return 1;
}

static u64 add4(u64 r[4], const u64 a[4], const u64 b[4]) {
  u128 c = 0;
  for (int i = 0; i < 4; ++i) {
    c += (u128)a[i] + b[i];
    r[i] = (u64)c;
    c >>= 64;
  }
  return (u64)c;
}

static u64 sub4(u64 r[4], const u64 a[4], const u64 b[4]) {
  u64 bw = 0;
  for (int i = 0; i < 4; ++i) {
    u128 t = (u128)a[i] - b[i] - bw;
    r[i] = (u64)t;
    bw = (u64)(t >> 64) & 1;
  }
  return bw;
}

/* (a + b) % P  -- big-num.cu:48 */
static void fadd(u64 r[4], const u64 a[4], const u64 b[4], const field_t* f) {
  u64 t[4], d[4];
  u64 c = add4(t, a, b);
  u64 bw = sub4(d, t, f->p);
  if (c || !bw) memcpy(r, d, 32); else memcpy(r, t, 32);
}

/* (a - b + P) % P  -- big-num.cu:49 */
static void fsub(u64 r[4], const u64 a[4], const u64 b[4], const field_t* f) {
  u64 t[4];
  if (sub4(t, a, b)) add4(t, t, f->p);
  memcpy(r, t, 32);
}

/* a*b*2^-256 mod p, canonical */
static void mont_mul(u64 r[4], const u64 a[4], const u64 b[4], const field_t* f) {
  u64 t[6] = {0, 0, 0, 0, 0, 0};
  for (int i = 0; i < 4; ++i) {
    u128 c = 0;
    for (int j = 0; j < 4; ++j) {
      c += (u128)a[j] * b[i] + t[j];
      t[j] = (u64)c;
      c >>= 64;
    }
    c += t[4];
    t[4] = (u64)c;
    t[5] = (u64)(c >> 64);
    u64 m = t[0] * f->np0;
    c = (u128)m * f->p[0] + t[0];
    c >>= 64;
    for (int j = 1; j < 4; ++j) {
      c += (u128)m * f->p[j] + t[j];
      t[j - 1] = (u64)c;
      c >>= 64;
    }
    c += t[4];
    t[3] = (u64)c;
    t[4] = t[5] + (u64)(c >> 64);
  }
  u64 d[4];
  u64 bw = sub4(d, t, f->p);
  if (t[4] || !bw) memcpy(r, d, 32); else memcpy(r, t, 32);
}

static void field_init(field_t* f, const uint32_t p32[8]) {
  memcpy(f->p, p32, 32);
  /* np0 by Newton iteration */
  u64 inv = 1;
  for (int i = 0; i < 6; ++i) inv *= 2 - f->p[0] * inv;
  f->np0 = (u64)0 - inv;
  /* R mod p by 256 doublings of 1, R^2 by 256 more */
  u64 x[4] = {1, 0, 0, 0};
  for (int i = 0; i < 512; ++i) {
    fadd(x, x, x, f);
    if (i == 255) memcpy(f->one, x, 32);
  }
  memcpy(f->r2, x, 32);
}

static void to_mont(u64 r[4], const u64 a[4], const field_t* f) { mont_mul(r, a, f->r2, f); }
static void from_mont(u64 r[4], const u64 a[4], const field_t* f) {
  u64 one[4] = {1, 0, 0, 0};
  mont_mul(r, a, one, f);
}

/* base^e, base and result in Montgomery form; e is a 256-bit integer */
static void mont_pow(u64 r[4], const u64 base[4], const u64 e[4], const field_t* f) {
  u64 acc[4], b[4];
  memcpy(acc, f->one, 32);
  memcpy(b, base, 32);
  for (int i = 0; i < 256; ++i) {
    if ((e[i >> 6] >> (i & 63)) & 1) mont_mul(acc, acc, b, f);
    mont_mul(b, b, b, f);
  }
  memcpy(r, acc, 32);
}

static void shr256(u64 r[4], const u64 a[4], int s) {
  for (int i = 0; i < 4; ++i) {
    int w = i + (s >> 6);
    u64 lo = w < 4 ? a[w] : 0, hi = (w + 1) < 4 ? a[w + 1] : 0;
    r[i] = (s & 63) ? (lo >> (s & 63)) | (hi << (64 - (s & 63))) : lo;
  }
}

/* omega_n = g^((p-1)/n) in Montgomery form; inverse -> (g^-1)^((p-1)/n)   (big-num.cu:128-134) */
static int root_of_unity(u64 w[4], const field_t* f, const uint32_t g32[8], int log_n, int inverse) {
  u64 g[4], gm[4], e[4], pm1[4], one[4] = {1, 0, 0, 0};
  memcpy(g, g32, 32);
  sub4(pm1, f->p, one);
  for (int i = 0; i < log_n; ++i)
    if ((pm1[i >> 6] >> (i & 63)) & 1) return -1; /* n does not divide p-1 */
  to_mont(gm, g, f);
  if (inverse) {
    u64 pm2[4], two[4] = {2, 0, 0, 0};
    sub4(pm2, f->p, two);
    mont_pow(gm, gm, pm2, f); /* inv(x) = x^(P-2), big-num.cu:27-29 */
  }
  shr256(e, pm1, log_n);
  mont_pow(w, gm, e, f);
  return 0;
}

static uint64_t bitrev(uint64_t i, int bits) {
  uint64_t r = 0;
  for (int b = 0; b < bits; ++b) r |= ((i >> b) & 1) << (bits - 1 - b);
  return r;
}

/* In-place NTT, natural order in and out, canonical residues.
 * threads <= 1: the literal single-threaded loop nest of big-num.cu:37-55 (w stepped by gap).
 * threads  > 1: same butterflies, each stage split across OpenMP threads, twiddles taken from a
 *               table roots[i] = omega_n^i (the table the reference's GPU driver builds,
 *               big-num.cu:298-306) -- identical results, exact arithmetic.
 * returns 0, or -1 when n does not divide p-1. */
int oracle_ntt256(uint32_t* data32, int log_n, const uint32_t p32[8], const uint32_t g32[8], int inverse,
                  int threads) {
  field_t f;
  field_init(&f, p32);
  u64(*data)[4] = (u64(*)[4])data32;
  const uint64_t n = (uint64_t)1 << log_n;
  u64 wn[4];
  if (root_of_unity(wn, &f, g32, log_n, inverse)) return -1;

  for (uint64_t i = 0; i < n; ++i) {
    uint64_t r = bitrev(i, log_n);
    if (i < r) {
      u64 t[4];
      memcpy(t, data[i], 32);
      memcpy(data[i], data[r], 32);
      memcpy(data[r], t, 32);
    }
  }

  if (threads <= 1) {
    for (int s = 0; s < log_n; ++s) {
      const uint64_t stride = (uint64_t)1 << s;
      /* gap = omega^((P-1)/(2*stride)) = omega_n^(n/(2*stride)) */
      u64 gap[4];
      memcpy(gap, wn, 32);
      for (int k = 0; k < log_n - s - 1; ++k) mont_mul(gap, gap, gap, &f);
      for (uint64_t start = 0; start < n; start += stride << 1) {
        u64 w[4];
        memcpy(w, f.one, 32);
        for (uint64_t off = 0; off < stride; ++off) {
          u64 *a = data[start + off], *bptr = data[start + off + stride], b[4], x[4];
          mont_mul(b, w, bptr, &f); /* (wR) * b / R = w*b % P */
          fadd(x, a, b, &f);
          fsub(bptr, a, b, &f);
          memcpy(a, x, 32);
          mont_mul(w, w, gap, &f);
        }
      }
    }
  } else {
    const uint64_t half = n > 1 ? n / 2 : 1;
    u64(*roots)[4] = (u64(*)[4])malloc(half * 32);
    if (!roots) return -2;
    const uint64_t chunk = 4096;
#pragma omp parallel for num_threads(threads) schedule(static)
    for (uint64_t c0 = 0; c0 < half; c0 += chunk) {
      u64 e[4] = {c0, 0, 0, 0}, w[4];
      mont_pow(w, wn, e, &f);
      for (uint64_t i = c0; i < c0 + chunk && i < half; ++i) {
        memcpy(roots[i], w, 32);
        mont_mul(w, w, wn, &f);
      }
    }
    for (int s = 0; s < log_n; ++s) {
      const uint64_t stride = (uint64_t)1 << s;
      const uint64_t tw_step = half >> s;
#pragma omp parallel for num_threads(threads) schedule(static)
      for (uint64_t i = 0; i < half; ++i) {
        const uint64_t off = i & (stride - 1);
        const uint64_t pos = ((i - off) << 1) + off;
        u64 *a = data[pos], *bptr = data[pos + stride], b[4], x[4];
        mont_mul(b, roots[off * tw_step], bptr, &f);
        fadd(x, a, b, &f);
        fsub(bptr, a, b, &f);
        memcpy(a, x, 32);
      }
    }
    free(roots);
  }

  if (inverse) {
    /* co = inv(length); data[i] * co % P  (big-num.cu:472-474) */
    u64 nn[4] = {n, 0, 0, 0}, nm[4], pm2[4], two[4] = {2, 0, 0, 0}, co[4];
    to_mont(nm, nn, &f);
    sub4(pm2, f.p, two);
    mont_pow(co, nm, pm2, &f);
#pragma omp parallel for num_threads(threads > 1 ? threads : 1) schedule(static)
    for (uint64_t i = 0; i < n; ++i) mont_mul(data[i], co, data[i], &f);
  }
  return 0;
}

/* Self-sort-in-place statement: self-sort-in-place.cu:79-128 (NTT_pro1 then NTT_pro2). */
int oracle_ntt256_ssip(uint32_t* data32, int log_n, const uint32_t p32[8], const uint32_t g32[8]) {
  field_t f;
  field_init(&f, p32);
  u64(*data)[4] = (u64(*)[4])data32;
  const uint64_t n = (uint64_t)1 << log_n;
  u64 wn[4];
  if (root_of_unity(wn, &f, g32, log_n, 0)) return -1;
  for (int i = log_n; i >= 1; --i) {
    const uint64_t stride = (uint64_t)1 << (i - 1);
    u64 gap[4];
    memcpy(gap, wn, 32);
    for (int k = 0; k < log_n - i; ++k) mont_mul(gap, gap, gap, &f);
    if (i > log_n / 2) {
      for (uint64_t start = 0; start < n; start += stride << 1) {
        u64 w[4];
        memcpy(w, f.one, 32);
        for (uint64_t off = 0; off < stride; ++off) {
          u64 *a = data[start + off], *b = data[start + off + stride], s[4], d[4];
          fadd(s, a, b, &f);
          fsub(d, a, b, &f);
          memcpy(a, s, 32);
          mont_mul(b, w, d, &f);
          mont_mul(w, w, gap, &f);
        }
      }
    } else {
      const uint64_t ps = (uint64_t)1 << (log_n - i);
      for (uint64_t start = 0; start < n; start += ps << 1)
        for (uint64_t off0 = 0; off0 < ps; off0 += stride << 1) {
          u64 w[4];
          memcpy(w, f.one, 32);
          for (uint64_t off = 0; off < stride; ++off) {
            const uint64_t q = start + off0 + off;
            u64 *a = data[q], *b = data[q + stride], *c = data[q + ps], *d = data[q + ps + stride];
            u64 s1[4], d1[4], s2[4], d2[4];
            fadd(s1, a, b, &f);
            fsub(d1, a, b, &f);
            fadd(s2, c, d, &f);
            fsub(d2, c, d, &f);
            memcpy(a, s1, 32);
            memcpy(b, s2, 32);
            mont_mul(c, w, d1, &f);
            mont_mul(d, w, d2, &f);
            mont_mul(w, w, gap, &f);
          }
        }
    }
  }
  return 0;
}

/* One output X[k] = sum_j x[j] * omega_n^(jk) by Horner -- spot checks at sizes where a full
 * second oracle run is too slow. */
int oracle_dft_sample(uint32_t out32[8], const uint32_t* data32, int log_n, const uint32_t p32[8],
                      const uint32_t g32[8], uint64_t k) {
  field_t f;
  field_init(&f, p32);
  const u64(*data)[4] = (const u64(*)[4])data32;
  const uint64_t n = (uint64_t)1 << log_n;
  u64 wn[4], wk[4], e[4] = {k, 0, 0, 0}, acc[4] = {0, 0, 0, 0};
  if (root_of_unity(wn, &f, g32, log_n, 0)) return -1;
  mont_pow(wk, wn, e, &f);
  for (uint64_t j = n; j-- > 0;) {
    mont_mul(acc, wk, acc, &f);
    fadd(acc, acc, data[j], &f);
  }
  memcpy(out32, acc, 32);
  return 0;
}

/* Pointwise helpers for size-independent property tests (linearity). */
void oracle_vec_add(uint32_t* r32, const uint32_t* a32, const uint32_t* b32, uint64_t n, const uint32_t p32[8]) {
  field_t f;
  field_init(&f, p32);
#pragma omp parallel for schedule(static)
  for (uint64_t i = 0; i < n; ++i)
    fadd((u64*)(r32 + 8 * i), (const u64*)(a32 + 8 * i), (const u64*)(b32 + 8 * i), &f);
}

/* Synthetic inputs: the SplitMix64 stream of oracle/py_oracle.py::splitmix_elements. */
static u64 splitmix_at(u64 seed, u64 idx) {
  u64 z = seed + (idx + 1) * 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

void oracle_splitmix_elements(uint32_t* out32, uint64_t n, const uint32_t p32[8], uint64_t seed) {
  u64 p[4];
  memcpy(p, p32, 32);
  int bits = 256;
  while (bits > 0 && !((p[(bits - 1) >> 6] >> ((bits - 1) & 63)) & 1)) --bits;
  u64 topmask = (bits & 63) ? (((u64)1 << (bits & 63)) - 1) : ~(u64)0;
  int topw = (bits - 1) >> 6;
  u64 draw = 0;
  u64(*out)[4] = (u64(*)[4])out32;
  for (uint64_t i = 0; i < n;) {
    u64 v[4];
    for (int k = 0; k < 4; ++k) v[k] = splitmix_at(seed, draw * 4 + k);
    ++draw;
    for (int k = topw + 1; k < 4; ++k) v[k] = 0;
    v[topw] &= topmask;
    if (!geq(v, p)) {
      memcpy(out[i], v, 32);
      ++i;
    }
  }
}

int oracle_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
