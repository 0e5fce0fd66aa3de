/* TEST INFRASTRUCTURE ONLY -- the CPU oracle of the 64-bit transforms (include/ntt64.h).  Never linked into or
 * called by the product path (ntt_b200/, libntt256.so); only tests/, __graft_entry__.smoke() and bench scripts'
 * CPU legs may load it.
 *
 * Plain-C restatement of the reference's CPU NTT in the container width the reference itself runs it in:
 *   oracle_ntt64():  NTT() at /root/reference/src/big-num.cu:37-55 -- bit-reverse permute (table rule of
 *                    big-num.cu:393-395), then radix-2 DIT stages with gap = omega^((P-1)/(2*stride)) and w
 *                    stepped by gap inside each block -- for any modulus below 2^64 (products on unsigned
 *                    __int128 instead of `long long` with `%`, which is the same arithmetic while it does not
 *                    overflow, i.e. for the reference's P = 469762049).  Inverse per the commented intent at
 *                    big-num.cu:470-477: the same transform with inv(omega), scaled by inv(len).
 * Parity status: PINNED for p = 469762049, g = 3 against tests/golden/ref_i64_kat.json (outputs of the
 * reference's own compiled NTT()) and, in this container, against oracle/_ref/libref_i64.so run live
 * (tests/test_oracle.py); other moduli (Goldilocks) are the same code with another constant and are also checked
 * against oracle/py_oracle.py (Python integers).
 *
 * Build: part of liboracle.so (oracle/Makefile).
 */
#include <stdint.h>
#include <stdlib.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef unsigned __int128 u128;
typedef uint64_t u64;

static u64 mulmod(u64 a, u64 b, u64 p) { return (u64)(((u128)a * b) % p); }
static u64 powmod(u64 b, u64 e, u64 p) {
  u64 acc = 1 % p;
  while (e) {
    if (e & 1) acc = mulmod(acc, b, p);
    b = mulmod(b, b, p);
    e >>= 1;
  }
  return acc;
}

/* data: n = 2^log_n canonical residues, transformed in place.  Returns 0, or 1 on bad arguments. */
int oracle_ntt64(u64* data, int log_n, u64 p, u64 g, int inverse, int threads) {
  if (!data || log_n < 0 || log_n > 40 || p < 3 || !(p & 1)) return 1;
  const u64 n = (u64)1 << log_n;
  if ((p - 1) % n) return 1;
#ifdef _OPENMP
  if (threads < 1) threads = 1;
#else
  (void)threads;
#endif
  /* big-num.cu:393-395 + 39-41: reverse table, swap when i < rev[i] */
  for (u64 i = 0; i < n; ++i) {
    u64 r = 0;
    for (int b = 0; b < log_n; ++b) r |= ((i >> b) & 1) << (log_n - 1 - b);
    if (i < r) {
      u64 t = data[i];
      data[i] = data[r];
      data[r] = t;
    }
  }
  const u64 omega = inverse ? powmod(g, p - 2, p) : g;
  for (u64 stride = 1; stride < n; stride <<= 1) {
    const u64 gap = powmod(omega, (p - 1) / (stride << 1), p); /* big-num.cu:43 */
    const u64 nblocks = n / (stride << 1);
    if (nblocks >= 64 || stride < 1024) {
#pragma omp parallel for num_threads(threads) schedule(static) if (n >= 4096)
      for (u64 blk = 0; blk < nblocks; ++blk) {
        const u64 start = blk * (stride << 1);
        u64 w = 1;
        for (u64 off = 0; off < stride; ++off) { /* big-num.cu:46-51 */
          const u64 a = data[start + off];
          const u64 b = mulmod(w, data[start + off + stride], p);
          const u64 s = a + b; /* may wrap for p > 2^63 */
          data[start + off] = (s < a || s >= p) ? s - p : s;
          data[start + off + stride] = a >= b ? a - b : a + (p - b);
          w = mulmod(gap, w, p);
        }
      }
    } else {
      /* few long blocks: split each block's offsets across threads, each chunk starting from gap^off */
      for (u64 blk = 0; blk < nblocks; ++blk) {
        const u64 start = blk * (stride << 1);
#pragma omp parallel num_threads(threads)
        {
#ifdef _OPENMP
          const int nt = omp_get_num_threads(), me = omp_get_thread_num();
#else
          const int nt = 1, me = 0;
#endif
          const u64 lo = stride * (u64)me / nt, hi = stride * (u64)(me + 1) / nt;
          u64 w = powmod(gap, lo, p);
          for (u64 off = lo; off < hi; ++off) {
            const u64 a = data[start + off];
            const u64 b = mulmod(w, data[start + off + stride], p);
            const u64 s = a + b;
            data[start + off] = (s < a || s >= p) ? s - p : s;
            data[start + off + stride] = a >= b ? a - b : a + (p - b);
            w = mulmod(gap, w, p);
          }
        }
      }
    }
  }
  if (inverse) {
    const u64 co = powmod(n % p, p - 2, p);
#pragma omp parallel for num_threads(threads) schedule(static) if (n >= 4096)
    for (u64 i = 0; i < n; ++i) data[i] = mulmod(data[i], co, p);
  }
  return 0;
}

/* seeded canonical elements (splitmix64 stream reduced mod p; the tiny bias is irrelevant for test inputs) */
void oracle_splitmix64_elements(u64* out, u64 n, u64 p, u64 seed) {
  u64 s = seed;
  for (u64 i = 0; i < n; ++i) {
    s += 0x9E3779B97F4A7C15ull;
    u64 z = s;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    out[i] = z % p;
  }
}
