/* TEST INFRASTRUCTURE ONLY -- the CPU oracle for wider containers (W x 64-bit limbs, W = 6: the
 * 384-bit path of include/ntt384.h).  Never linked into or called by the product path.
 *
 * Same restatement as ntt_oracle.c of the reference's CPU NTT (/root/reference/src/big-num.cu:37-55:
 * bit-reverse permute, then radix-2 DIT stages; omega_n = g^((p-1)/n), big-num.cu:128-134; inverse
 * per the commented intent at big-num.cu:470-477), over a field held in W limbs.  Memory image =
 * cgbn_mem_t<64*W> (include/cgbn/cgbn_cuda.h:51-55): little-endian limbs, canonical residues.
 * Parity status: pinned by tests/golden/ref512_gpu_kat.json (the reference's own NTT_GZKP at 512-bit
 * containers over BLS12-377 Fq, run on a B200) and cross-checked against py_oracle.py.
 *
 * Build: gcc -O3 -fopenmp -DW=6 -DNAME=oracle_ntt384 -c ntt_oracle_n.c */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifndef W
#define W 6
#endif
#ifndef NAME
#define NAME oracle_ntt384
#endif
#define BYTES (8 * W)

typedef unsigned __int128 u128;
typedef uint64_t u64;

typedef struct {
  u64 p[W], r2[W], one[W];
  u64 np0;
} fieldn_t;

static u64 addn(u64* r, const u64* a, const u64* b) {
  u128 c = 0;
  for (int i = 0; i < W; ++i) {
    c += (u128)a[i] + b[i];
    r[i] = (u64)c;
    c >>= 64;
  }
  return (u64)c;
}
static u64 subn(u64* r, const u64* a, const u64* b) {
  u64 bw = 0;
  for (int i = 0; i < W; ++i) {
    u128 t = (u128)a[i] - b[i] - bw;
    r[i] = (u64)t;
    bw = (u64)(t >> 64) & 1;
  }
  return bw;
}
/* (a + b) % P  -- big-num.cu:48 */
static void faddn(u64* r, const u64* a, const u64* b, const fieldn_t* f) {
  u64 t[W], d[W];
  u64 c = addn(t, a, b);
  u64 bw = subn(d, t, f->p);
  memcpy(r, (c || !bw) ? d : t, BYTES);
}
/* (a - b + P) % P  -- big-num.cu:49 */
static void fsubn(u64* r, const u64* a, const u64* b, const fieldn_t* f) {
  u64 t[W];
  if (subn(t, a, b)) addn(t, t, f->p);
  memcpy(r, t, BYTES);
}
/* a*b*2^(-64W) mod p, canonical */
static void mont_muln(u64* r, const u64* a, const u64* b, const fieldn_t* f) {
  u64 t[W + 2];
  memset(t, 0, sizeof(t));
  for (int i = 0; i < W; ++i) {
    u128 c = 0;
    for (int j = 0; j < W; ++j) {
      c += (u128)a[j] * b[i] + t[j];
      t[j] = (u64)c;
      c >>= 64;
    }
    c += t[W];
    t[W] = (u64)c;
    t[W + 1] = (u64)(c >> 64);
    u64 m = t[0] * f->np0;
    c = (u128)m * f->p[0] + t[0];
    c >>= 64;
    for (int j = 1; j < W; ++j) {
      c += (u128)m * f->p[j] + t[j];
      t[j - 1] = (u64)c;
      c >>= 64;
    }
    c += t[W];
    t[W - 1] = (u64)c;
    t[W] = t[W + 1] + (u64)(c >> 64);
  }
  u64 d[W];
  u64 bw = subn(d, t, f->p);
  memcpy(r, (t[W] || !bw) ? d : t, BYTES);
}
static void field_initn(fieldn_t* f, const uint32_t* p32) {
  memcpy(f->p, p32, BYTES);
  u64 inv = 1;
  for (int i = 0; i < 6; ++i) inv *= 2 - f->p[0] * inv;
  f->np0 = (u64)0 - inv;
  u64 x[W];
  memset(x, 0, sizeof(x));
  x[0] = 1;
  for (int i = 0; i < 128 * W; ++i) {
    faddn(x, x, x, f);
    if (i == 64 * W - 1) memcpy(f->one, x, BYTES);
  }
  memcpy(f->r2, x, BYTES);
}
static void mont_pown(u64* r, const u64* base, const u64* e, const fieldn_t* f) {
  u64 acc[W], b[W];
  memcpy(acc, f->one, BYTES);
  memcpy(b, base, BYTES);
  for (int i = 0; i < 64 * W; ++i) {
    if ((e[i >> 6] >> (i & 63)) & 1) mont_muln(acc, acc, b, f);
    mont_muln(b, b, b, f);
  }
  memcpy(r, acc, BYTES);
}
static void shrn(u64* r, const u64* a, int s) {
  for (int i = 0; i < W; ++i) {
    int w = i + (s >> 6);
    u64 lo = w < W ? a[w] : 0, hi = (w + 1) < W ? a[w + 1] : 0;
    r[i] = (s & 63) ? (lo >> (s & 63)) | (hi << (64 - (s & 63))) : lo;
  }
}
static uint64_t bitrevn(uint64_t i, int bits) {
  uint64_t r = 0;
  for (int b = 0; b < bits; ++b) r |= ((i >> b) & 1) << (bits - 1 - b);
  return r;
}

/* In-place NTT of n = 2^log_n elements (2W u32 limbs each), natural order in and out, canonical.
 * Stages are split across `threads` OpenMP threads with twiddles from a table roots[i] = omega_n^i
 * (the table the reference's GPU driver builds, big-num.cu:298-306); exact arithmetic, so the result
 * equals the single-threaded loop nest of big-num.cu:37-55.  Returns 0, -1 if n does not divide p-1. */
int NAME(uint32_t* data32, int log_n, const uint32_t* p32, const uint32_t* g32, int inverse, int threads) {
  fieldn_t f;
  field_initn(&f, p32);
  u64(*data)[W] = (u64(*)[W])data32;
  const uint64_t n = (uint64_t)1 << log_n;
  u64 g[W], gm[W], e[W], pm1[W], pm2[W], one[W], two[W], wn[W];
  memset(one, 0, sizeof(one));
  memset(two, 0, sizeof(two));
  one[0] = 1;
  two[0] = 2;
  memcpy(g, g32, BYTES);
  subn(pm1, f.p, one);
  subn(pm2, f.p, two);
  for (int i = 0; i < log_n; ++i)
    if ((pm1[i >> 6] >> (i & 63)) & 1) return -1;
  mont_muln(gm, g, f.r2, &f);
  if (inverse) mont_pown(gm, gm, pm2, &f); /* inv(x) = x^(P-2), big-num.cu:27-29 */
  shrn(e, pm1, log_n);
  mont_pown(wn, gm, e, &f);

  for (uint64_t i = 0; i < n; ++i) {
    uint64_t r = bitrevn(i, log_n);
    if (i < r) {
      u64 t[W];
      memcpy(t, data[i], BYTES);
      memcpy(data[i], data[r], BYTES);
      memcpy(data[r], t, BYTES);
    }
  }
  if (threads < 1) threads = 1;
  const uint64_t half = n > 1 ? n / 2 : 1;
  u64(*roots)[W] = (u64(*)[W])malloc(half * BYTES);
  if (!roots) return -2;
  const uint64_t chunk = 4096;
#pragma omp parallel for num_threads(threads) schedule(static)
  for (uint64_t c0 = 0; c0 < half; c0 += chunk) {
    u64 ee[W], w[W];
    memset(ee, 0, sizeof(ee));
    ee[0] = c0;
    mont_pown(w, wn, ee, &f);
    for (uint64_t i = c0; i < c0 + chunk && i < half; ++i) {
      memcpy(roots[i], w, BYTES);
      mont_muln(w, w, wn, &f);
    }
  }
  for (int s = 0; s < log_n; ++s) {
    const uint64_t stride = (uint64_t)1 << s;
    const uint64_t tw_step = half >> s;
#pragma omp parallel for num_threads(threads) schedule(static)
    for (uint64_t i = 0; i < half; ++i) {
      const uint64_t off = i & (stride - 1);
      const uint64_t pos = ((i - off) << 1) + off;
      u64 *a = data[pos], *bptr = data[pos + stride], b[W], x[W];
      mont_muln(b, roots[off * tw_step], bptr, &f); /* (wR) * b / R = w*b % P */
      faddn(x, a, b, &f);
      fsubn(bptr, a, b, &f);
      memcpy(a, x, BYTES);
    }
  }
  free(roots);
  if (inverse) {
    u64 nn[W], nm[W], co[W];
    memset(nn, 0, sizeof(nn));
    nn[0] = n;
    mont_muln(nm, nn, f.r2, &f);
    mont_pown(co, nm, pm2, &f); /* co = inv(length), big-num.cu:472-474 */
#pragma omp parallel for num_threads(threads) schedule(static)
    for (uint64_t i = 0; i < n; ++i) mont_muln(data[i], co, data[i], &f);
  }
  return 0;
}
