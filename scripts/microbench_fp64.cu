// Experiment (round 1): can the idle FP64 pipe of B200 (64 DFMA/clk/SM, co-issues with the
// integer pipe, profiles/r01_microbench_pipes.jsonl) serve as a second 256-bit multiplier?
//
// mont_mul_f64: Montgomery product on 11 limbs of 24 bits held in doubles, R = 2^264.
// Product scanning; every column sum has <= 22 products of < 2^48 plus a carry < 2^29, i.e. stays
// below 2^53 and is EXACT in a chain of DFMAs.  Splits use the 2^52 magic constant with
// round-down FMAs.  363 FP64 instructions per product, zero integer instructions.
//
// The program (1) checks mont_mul_f64 against the integer mont_mul of field.cuh on random
// operands (both reduced to canonical values), (2) measures FP64-only, INT-only and interleaved
// INT+FP64 throughput.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "../ntt_b200/csrc/field.cuh"
#include "../ntt_b200/csrc/hostfield.h"

using namespace ntt256;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

struct F64Params {
  double p[11];
  double np0;  // -p^-1 mod 2^24
};

#define TWO52 4503599627370496.0
#define TWO24 16777216.0
#define INV24 (1.0 / 16777216.0)

__device__ __forceinline__ double fma_rd(double a, double b, double c) {
  double r;
  asm("fma.rm.f64 %0, %1, %2, %3;" : "=d"(r) : "d"(a), "d"(b), "d"(c));
  return r;
}
// t = hi*2^24 + lo, 0 <= lo < 2^24, t an exact non-negative integer < 2^53
__device__ __forceinline__ void split24(double t, double& hi, double& lo) {
  hi = fma_rd(t, INV24, TWO52) - TWO52;
  lo = fma(hi, -TWO24, t);
}

// r = a*b*2^-264 mod p (+ at most p): a < p, b < 2^256, all limbs integers in [0, 2^24)
__device__ __forceinline__ void mont_mul_f64(double* r, const double* a, const double* b, const F64Params& f) {
  double m[11];
  double carry = 0.0;
#pragma unroll
  for (int k = 0; k < 11; ++k) {
    double t = carry;
#pragma unroll
    for (int i = 0; i <= k; ++i) t = fma(a[i], b[k - i], t);
#pragma unroll
    for (int i = 0; i < k; ++i) t = fma(m[i], f.p[k - i], t);
    double hi, lo, q, qh;
    split24(t, hi, lo);
    q = lo * f.np0;
    split24(q, qh, m[k]);
    t = fma(m[k], f.p[0], t);
    carry = t * INV24;
  }
#pragma unroll
  for (int k = 11; k < 21; ++k) {
    double t = carry;
#pragma unroll
    for (int i = k - 10; i <= 10; ++i) t = fma(a[i], b[k - i], t);
#pragma unroll
    for (int i = k - 10; i <= 10; ++i) t = fma(m[i], f.p[k - i], t);
    split24(t, carry, r[k - 11]);
  }
  r[10] = carry;
}

// ---------------------------------------------------------------------------------------------
// E_r29: carry-free integer Montgomery product, 9 limbs x 29 bits, R = 2^261, plain mad.wide.u32
// only (full-rate IMAD.WIDE, no carry predicates, so it does not collide with the FP64 pipe).
// a: 9 limbs < 2^29 (twiddle * 2^261 mod p); b and the result: 8 x u32 (value < 2^256 / < 2p).
struct R29Params {
  u32 p[9];
  u32 np0;  // -p^-1 mod 2^29
};
#define M29 0x1fffffffu

__device__ __forceinline__ void mont_mul_r29(u32* r, const u32* a, const u32* x, const R29Params& f) {
  // split x into 29-bit limbs
  u32 b[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) {
    const int bit = 29 * i, w = bit >> 5, s = bit & 31;
    u32 v;
    if (s == 0) v = x[w];
    else if (w + 1 < 8) v = __funnelshift_r(x[w], x[w + 1], s);
    else v = x[w] >> s;
    b[i] = v & M29;
  }
  u64 acc[9];
#pragma unroll
  for (int j = 0; j < 9; ++j) acc[j] = 0;
#pragma unroll
  for (int i = 0; i < 9; ++i) {
#pragma unroll
    for (int j = 0; j < 9; ++j) acc[j] += (u64)a[j] * b[i];
    const u32 m = ((u32)acc[0] * f.np0) & M29;
#pragma unroll
    for (int j = 0; j < 9; ++j) acc[j] += (u64)f.p[j] * m;
    const u64 c = acc[0] >> 29;
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = acc[j + 1];
    acc[0] += c;
    acc[8] = 0;
  }
  // normalise to 29-bit limbs and pack into 8 x u32
  u32 l[9];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    l[j] = (u32)acc[j] & M29;
    acc[j + 1] += acc[j] >> 29;
  }
  l[8] = (u32)acc[8];
#pragma unroll
  for (int w = 0; w < 8; ++w) {
    // word w covers bits [32w, 32w+32)
    const int lo_l = (32 * w) / 29, s = 32 * w - 29 * lo_l;  // first limb touching the word, offset inside it
    u32 v = l[lo_l] >> s;
    v |= l[lo_l + 1] << (29 - s);
    if (29 - s + 29 < 32 && lo_l + 2 < 9) v |= l[lo_l + 2] << (58 - s);
    r[w] = v;
  }
}

// ---- format conversion (integer side): 8 x u32 <-> 11 x 24-bit doubles
__device__ __forceinline__ void to_f64_limbs(double* d, const u32* x) {
#pragma unroll
  for (int k = 0; k < 11; ++k) {
    const int bit = 24 * k, w = bit >> 5, s = bit & 31;
    u32 v = x[w] >> s;
    if (s > 8 && w + 1 < 8) v |= x[w + 1] << (32 - s);
    v &= 0xffffffu;
    d[k] = __hiloint2double(0x43300000, v) - TWO52;  // exact int -> double without I2F
  }
}
__device__ __forceinline__ void from_f64_limbs(u32* x, const double* d) {
  u32 v[11];
#pragma unroll
  for (int k = 0; k < 11; ++k) v[k] = (u32)__double2loint(d[k] + TWO52);
#pragma unroll
  for (int w = 0; w < 8; ++w) x[w] = 0;
#pragma unroll
  for (int k = 0; k < 11; ++k) {
    const int bit = 24 * k, w = bit >> 5, s = bit & 31;
    if (w < 8) x[w] |= v[k] << s;
    if (s > 8 && w + 1 < 8) x[w + 1] |= v[k] >> (32 - s);
  }
}

__global__ void k_check(u32* out_int, u32* out_f64, const u32* a_int, const u32* a_f64, const u32* b,
                        const __grid_constant__ FieldParams fi, const __grid_constant__ F64Params ff, int n) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  u32 ai[8], af[8], bb[8], r1[8], r2[8];
  for (int k = 0; k < 8; ++k) { ai[k] = a_int[t * 8 + k]; af[k] = a_f64[t * 8 + k]; bb[k] = b[t * 8 + k]; }
  mont_mul(r1, ai, bb, fi);
  canonicalize(r1, fi);
  double da[11], db[11], dr[11];
  to_f64_limbs(da, af);
  to_f64_limbs(db, bb);
  mont_mul_f64(dr, da, db, ff);
  from_f64_limbs(r2, dr);
  canonicalize(r2, fi);
  for (int k = 0; k < 8; ++k) { out_int[t * 8 + k] = r1[k]; out_f64[t * 8 + k] = r2[k]; }
}

__global__ void k_check_r29(u32* out, const u32* a29, const u32* b, const __grid_constant__ FieldParams fi,
                            const __grid_constant__ R29Params fr, int n) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  u32 a[9], bb[8], r[8];
  for (int k = 0; k < 9; ++k) a[k] = a29[t * 9 + k];
  for (int k = 0; k < 8; ++k) bb[k] = b[t * 8 + k];
  mont_mul_r29(r, a, bb, fr);
  canonicalize(r, fi);
  for (int k = 0; k < 8; ++k) out[t * 8 + k] = r[k];
}

// NR r29 products + NF f64 products per iteration, independent chains, same thread
template <int NR, int NF>
__global__ void __launch_bounds__(256) k_mix_r29(u32* out, const u32* in, const __grid_constant__ R29Params fr,
                                                 const __grid_constant__ F64Params ff, int iters) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  u32 a[9], a8[8], x[NR > 0 ? NR : 1][8];
  double da[11], dx[NF > 0 ? NF : 1][11];
#pragma unroll
  for (int k = 0; k < 8; ++k) a8[k] = in[(t & 1023) * 8 + k] & (k == 7 ? 0x0fffffffu : 0xffffffffu);
#pragma unroll
  for (int k = 0; k < 9; ++k) a[k] = in[(t & 1023) * 8 + (k & 7)] & (k == 8 ? 0xfffffu : M29);
  to_f64_limbs(da, a8);
#pragma unroll
  for (int n = 0; n < NR; ++n)
#pragma unroll
    for (int k = 0; k < 8; ++k) x[n][k] = in[((t + 7 * n + 1) & 1023) * 8 + k];
#pragma unroll
  for (int n = 0; n < NF; ++n) {
    u32 tmp[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) tmp[k] = in[((t + 13 * n + 5) & 1023) * 8 + k];
    to_f64_limbs(dx[n], tmp);
  }
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int n = 0; n < (NR > NF ? NR : NF); ++n) {
      if (n < NR) {
        u32 r[8];
        mont_mul_r29(r, a, x[n], fr);
#pragma unroll
        for (int k = 0; k < 8; ++k) x[n][k] = r[k];
      }
      if (n < NF) {
        double r[11];
        mont_mul_f64(r, da, dx[n], ff);
#pragma unroll
        for (int k = 0; k < 11; ++k) dx[n][k] = r[k];
      }
    }
  }
  u32 s = 0;
  double sd = 0;
#pragma unroll
  for (int n = 0; n < NR; ++n)
#pragma unroll
    for (int k = 0; k < 8; ++k) s ^= x[n][k];
#pragma unroll
  for (int n = 0; n < NF; ++n)
#pragma unroll
    for (int k = 0; k < 11; ++k) sd += dx[n][k];
  if (s == 0x12345678) out[0] = s;
  if (sd == 1.2345) out[1] = 1;
}

template <int NR, int NF>
static void bench_r29(const char* name, int sms, u32* d_out, u32* d_in, const R29Params& fr, const F64Params& ff) {
  const int threads = 256, iters = 200;
  for (int occ = 1; occ <= 2; ++occ) {
    int blocks = sms * occ;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k_mix_r29<NR, NF><<<blocks, threads>>>(d_out, d_in, fr, ff, iters);
    CK(cudaDeviceSynchronize());
    double best = 1e30;
    for (int rep = 0; rep < 3; ++rep) {
      cudaEventRecord(e0);
      k_mix_r29<NR, NF><<<blocks, threads>>>(d_out, d_in, fr, ff, iters);
      cudaEventRecord(e1);
      CK(cudaEventSynchronize(e1));
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (ms < best) best = ms;
    }
    double tot = (double)blocks * threads * iters * (NR + NF);
    printf("{\"test\": \"%s\", \"r29_per_iter\": %d, \"f64_per_iter\": %d, \"blocks_per_sm\": %d, \"ms\": %.4f, \"Gmodmul\": %.2f}\n",
           name, NR, NF, occ, best, tot / best / 1e6);
  }
}

template <int NI, int NF>
__global__ void __launch_bounds__(256) k_mix(u32* out, const u32* in, const __grid_constant__ FieldParams fi,
                                             const __grid_constant__ F64Params ff, int iters) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  u32 a[8], x[NI > 0 ? NI : 1][8];
  double da[11], dx[NF > 0 ? NF : 1][11];
#pragma unroll
  for (int k = 0; k < 8; ++k) a[k] = in[(t & 1023) * 8 + k] & (k == 7 ? 0x0fffffffu : 0xffffffffu);
  to_f64_limbs(da, a);
#pragma unroll
  for (int n = 0; n < NI; ++n)
#pragma unroll
    for (int k = 0; k < 8; ++k) x[n][k] = in[((t + 7 * n + 1) & 1023) * 8 + k];
#pragma unroll
  for (int n = 0; n < NF; ++n) {
    u32 tmp[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) tmp[k] = in[((t + 13 * n + 5) & 1023) * 8 + k];
    to_f64_limbs(dx[n], tmp);
  }
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int n = 0; n < (NI > NF ? NI : NF); ++n) {
      if (n < NI) {
        u32 r[8];
        mont_mul(r, a, x[n], fi);
#pragma unroll
        for (int k = 0; k < 8; ++k) x[n][k] = r[k];
      }
      if (n < NF) {
        double r[11];
        mont_mul_f64(r, da, dx[n], ff);
#pragma unroll
        for (int k = 0; k < 11; ++k) dx[n][k] = r[k];
      }
    }
  }
  u32 s = 0;
  double sd = 0;
#pragma unroll
  for (int n = 0; n < NI; ++n)
#pragma unroll
    for (int k = 0; k < 8; ++k) s ^= x[n][k];
#pragma unroll
  for (int n = 0; n < NF; ++n)
#pragma unroll
    for (int k = 0; k < 11; ++k) sd += dx[n][k];
  if (s == 0x12345678) out[0] = s;
  if (sd == 1.2345) out[1] = 1;
}

// Same work, but odd warps start with the FP64 product and even warps with the integer one, so
// that at any time some warps of an SM sub-partition feed the FP64 pipe and others the INT pipe.
template <int NI, int NF>
__global__ void __launch_bounds__(256) k_stagger(u32* out, const u32* in, const __grid_constant__ FieldParams fi,
                                                 const __grid_constant__ F64Params ff, int iters) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  u32 a[8], x[NI][8];
  double da[11], dx[NF][11];
#pragma unroll
  for (int k = 0; k < 8; ++k) a[k] = in[(t & 1023) * 8 + k] & (k == 7 ? 0x0fffffffu : 0xffffffffu);
  to_f64_limbs(da, a);
#pragma unroll
  for (int n = 0; n < NI; ++n)
#pragma unroll
    for (int k = 0; k < 8; ++k) x[n][k] = in[((t + 7 * n + 1) & 1023) * 8 + k];
#pragma unroll
  for (int n = 0; n < NF; ++n) {
    u32 tmp[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) tmp[k] = in[((t + 13 * n + 5) & 1023) * 8 + k];
    to_f64_limbs(dx[n], tmp);
  }
  auto do_int = [&]() {
#pragma unroll
    for (int n = 0; n < NI; ++n) {
      u32 r[8];
      mont_mul(r, a, x[n], fi);
#pragma unroll
      for (int k = 0; k < 8; ++k) x[n][k] = r[k];
    }
  };
  auto do_f64 = [&]() {
#pragma unroll
    for (int n = 0; n < NF; ++n) {
      double r[11];
      mont_mul_f64(r, da, dx[n], ff);
#pragma unroll
      for (int k = 0; k < 11; ++k) dx[n][k] = r[k];
    }
  };
  const bool odd = (threadIdx.x >> 5) & 1;
  if (odd) do_f64();
  for (int i = 0; i < iters; ++i) {
    do_int();
    if (!(odd && i == iters - 1)) do_f64();
  }
  u32 s = 0;
  double sd = 0;
#pragma unroll
  for (int n = 0; n < NI; ++n)
#pragma unroll
    for (int k = 0; k < 8; ++k) s ^= x[n][k];
#pragma unroll
  for (int n = 0; n < NF; ++n)
#pragma unroll
    for (int k = 0; k < 11; ++k) sd += dx[n][k];
  if (s == 0x12345678) out[0] = s;
  if (sd == 1.2345) out[1] = 1;
}

template <int NI, int NF>
static void bench_stagger(const char* name, int sms, u32* d_out, u32* d_in, const FieldParams& fi, const F64Params& ff) {
  const int threads = 256, iters = 200;
  for (int occ = 1; occ <= 2; ++occ) {
    int blocks = sms * occ;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k_stagger<NI, NF><<<blocks, threads>>>(d_out, d_in, fi, ff, iters);
    CK(cudaDeviceSynchronize());
    double best = 1e30;
    for (int rep = 0; rep < 3; ++rep) {
      cudaEventRecord(e0);
      k_stagger<NI, NF><<<blocks, threads>>>(d_out, d_in, fi, ff, iters);
      cudaEventRecord(e1);
      CK(cudaEventSynchronize(e1));
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (ms < best) best = ms;
    }
    double tot = (double)blocks * threads * iters * (NI + NF);
    printf("{\"test\": \"%s\", \"int_per_iter\": %d, \"f64_per_iter\": %d, \"blocks_per_sm\": %d, \"ms\": %.4f, \"Gmodmul\": %.2f}\n",
           name, NI, NF, occ, best, tot / best / 1e6);
  }
}

template <int NI, int NF>
static void bench(const char* name, int sms, u32* d_out, u32* d_in, const FieldParams& fi, const F64Params& ff) {
  const int threads = 256, iters = 200;
  for (int occ = 1; occ <= 2; ++occ) {
    int blocks = sms * occ;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k_mix<NI, NF><<<blocks, threads>>>(d_out, d_in, fi, ff, iters);
    CK(cudaDeviceSynchronize());
    double best = 1e30;
    for (int rep = 0; rep < 3; ++rep) {
      cudaEventRecord(e0);
      k_mix<NI, NF><<<blocks, threads>>>(d_out, d_in, fi, ff, iters);
      cudaEventRecord(e1);
      CK(cudaEventSynchronize(e1));
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (ms < best) best = ms;
    }
    double tot = (double)blocks * threads * iters * (NI + NF);
    printf("{\"test\": \"%s\", \"int_per_iter\": %d, \"f64_per_iter\": %d, \"blocks_per_sm\": %d, \"ms\": %.4f, \"Gmodmul\": %.2f}\n",
           name, NI, NF, occ, best, tot / best / 1e6);
  }
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  int sms = prop.multiProcessorCount;
  // BN254 Fr
  const u32 p32[8] = {0xf0000001u, 0x43e1f593u, 0x79b97091u, 0x2833e848u, 0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
  HostField hf; hf.init(u256_from_limbs(p32));
  FieldParams fi; memset(&fi, 0, sizeof(fi));
  for (int k = 0; k < 8; ++k) fi.p[k] = p32[k];
  U256 P2; u256_add(P2, hf.p, hf.p); u256_to_limbs(P2, fi.p2); fi.np0 = (u32)hf.np0;
  F64Params ff;
  for (int k = 0; k < 11; ++k) {
    int bit = 24 * k; u64 v = 0;
    for (int b = 0; b < 24 && bit + b < 256; ++b) v |= (u64)u256_bit(hf.p, bit + b) << b;
    ff.p[k] = (double)v;
  }
  { u64 inv = 1; for (int i = 0; i < 6; ++i) inv *= 2 - hf.p.w[0] * inv; ff.np0 = (double)((0 - inv) & 0xffffffull); }

  R29Params fr;
  for (int k = 0; k < 9; ++k) {
    int bit = 29 * k; u64 v = 0;
    for (int b = 0; b < 29 && bit + b < 256; ++b) v |= (u64)u256_bit(hf.p, bit + b) << b;
    fr.p[k] = (u32)v;
  }
  { u64 inv = 1; for (int i = 0; i < 6; ++i) inv *= 2 - hf.p.w[0] * inv; fr.np0 = (u32)((0 - inv) & M29); }

  // ---- correctness: w*x via INT (twiddle in R=2^256 form) vs FP64 (twiddle in R=2^264 form)
  const int n = 1 << 14;
  std::vector<u32> a_int(n * 8), a_f64(n * 8), b(n * 8), r1(n * 8), r2(n * 8);
  u64 s = 0x9e3779b97f4a7c15ull;
  auto rnd = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; };
  U256 r264 = hf.one;  // 2^256 mod p
  for (int i = 0; i < 8; ++i) r264 = hf.add(r264, r264);  // 2^264 mod p (plain value)
  for (int i = 0; i < n; ++i) {
    U256 w{{rnd(), rnd(), rnd(), rnd() >> 3}};
    while (u256_cmp(w, hf.p) >= 0) w.w[3] >>= 1;
    U256 x{{rnd(), rnd(), rnd(), rnd()}};
    if (i % 7 == 0) x = U256{{~0ull, ~0ull, ~0ull, ~0ull}};
    if (i % 11 == 0) { u256_sub(w, hf.p, u256_small(1)); }
    U256 wm = hf.to_mont(w);                       // w * 2^256
    U256 wf = hf.mul(hf.to_mont(w), hf.to_mont(r264));  // mont(wR * r264R) = w * r264 * R... see below
    // hf.mul(xR, yR) = x*y*R ; we need the plain value w*2^264 mod p = from_mont(mul(to_mont(w), to_mont(r264)))
    wf = hf.from_mont(wf);
    u256_to_limbs(wm, &a_int[i * 8]);
    u256_to_limbs(wf, &a_f64[i * 8]);
    u256_to_limbs(x, &b[i * 8]);
  }
  // r29 twiddles: w * 2^261 mod p as 9 x 29-bit limbs
  std::vector<u32> a_r29(n * 9), r3(n * 8);
  {
    U256 r261 = hf.one;
    for (int i = 0; i < 5; ++i) r261 = hf.add(r261, r261);  // 2^261 mod p
    for (int i = 0; i < n; ++i) {
      U256 wm = u256_from_limbs(&a_int[i * 8]);            // w * 2^256
      U256 w = hf.from_mont(wm);
      U256 w261 = hf.from_mont(hf.mul(hf.to_mont(w), hf.to_mont(r261)));
      for (int k = 0; k < 9; ++k) {
        int bit = 29 * k; u64 v = 0;
        for (int b2 = 0; b2 < 29 && bit + b2 < 256; ++b2) v |= (u64)u256_bit(w261, bit + b2) << b2;
        a_r29[i * 9 + k] = (u32)v;
      }
    }
  }
  u32 *d_ai, *d_af, *d_b, *d_r1, *d_r2;
  CK(cudaMalloc(&d_ai, n * 32)); CK(cudaMalloc(&d_af, n * 32)); CK(cudaMalloc(&d_b, n * 32));
  CK(cudaMalloc(&d_r1, n * 32)); CK(cudaMalloc(&d_r2, n * 32));
  CK(cudaMemcpy(d_ai, a_int.data(), n * 32, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_af, a_f64.data(), n * 32, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_b, b.data(), n * 32, cudaMemcpyHostToDevice));
  k_check<<<n / 128, 128>>>(d_r1, d_r2, d_ai, d_af, d_b, fi, ff, n);
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(r1.data(), d_r1, n * 32, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(r2.data(), d_r2, n * 32, cudaMemcpyDeviceToHost));
  int bad = 0;
  for (int i = 0; i < n * 8; ++i) bad += r1[i] != r2[i];
  printf("{\"test\": \"f64_vs_int_montmul\", \"cases\": %d, \"mismatching_words\": %d}\n", n, bad);

  {
    u32 *d_a29, *d_r3;
    CK(cudaMalloc(&d_a29, n * 36)); CK(cudaMalloc(&d_r3, n * 32));
    CK(cudaMemcpy(d_a29, a_r29.data(), n * 36, cudaMemcpyHostToDevice));
    k_check_r29<<<n / 128, 128>>>(d_r3, d_a29, d_b, fi, fr, n);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(r3.data(), d_r3, n * 32, cudaMemcpyDeviceToHost));
    int bad3 = 0;
    for (int i = 0; i < n * 8; ++i) bad3 += r1[i] != r3[i];
    printf("{\"test\": \"r29_vs_int_montmul\", \"cases\": %d, \"mismatching_words\": %d}\n", n, bad3);
  }
  u32* d_out; u32* d_in;
  CK(cudaMalloc(&d_out, 4096)); CK(cudaMalloc(&d_in, 1024 * 32));
  CK(cudaMemcpy(d_in, b.data(), 1024 * 32, cudaMemcpyHostToDevice));
  bench_r29<2, 0>("r29_only", sms, d_out, d_in, fr, ff);
  bench_r29<1, 1>("r29_1_f64_1", sms, d_out, d_in, fr, ff);
  bench_r29<2, 1>("r29_2_f64_1", sms, d_out, d_in, fr, ff);
  bench_r29<3, 2>("r29_3_f64_2", sms, d_out, d_in, fr, ff);
  bench<2, 0>("int_only", sms, d_out, d_in, fi, ff);
  bench<0, 2>("f64_only", sms, d_out, d_in, fi, ff);
  bench<0, 1>("f64_only_x1", sms, d_out, d_in, fi, ff);
  bench<1, 1>("int1_f64_1", sms, d_out, d_in, fi, ff);
  bench<2, 1>("int2_f64_1", sms, d_out, d_in, fi, ff);
  bench<2, 2>("int2_f64_2", sms, d_out, d_in, fi, ff);
  bench_stagger<1, 1>("stagger_int1_f64_1", sms, d_out, d_in, fi, ff);
  bench_stagger<3, 2>("stagger_int3_f64_2", sms, d_out, d_in, fi, ff);
  bench_stagger<2, 1>("stagger_int2_f64_1", sms, d_out, d_in, fi, ff);
  return 0;
}
