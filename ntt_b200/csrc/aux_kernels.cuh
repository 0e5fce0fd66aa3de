// Kernels either side of the passes (sm_100a): tiny transforms, table builders and the element-wise companions of
// SURVEY.md par. 8f.  Non-template __global__ functions: included by capi.cu only (the pass kernels' translation
// units include ntt_kernels.cuh alone).
#pragma once
#include "ntt_kernels.cuh"

namespace ntt256 {

// Transforms of 2 or 4 points: one thread per transform, exact radix-2 DIF + bit reversal.
struct TinyArgs {
  FieldParams f;
  int log_n;
  u32 w4[8];      // w_4 (or its inverse), Montgomery form
  u32 scale[8];   // 1/n Montgomery form
  int do_scale;
  unsigned long long batch;
};

__global__ void ntt_tiny_kernel(uint4* __restrict__ data, const __grid_constant__ TinyArgs A) {
  const unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= A.batch) return;
  const FieldParams& f = A.f;
  const int n = 1 << A.log_n;
  uint4* p = data + 2 * t * n;
  u32 x[4][8];
  for (int i = 0; i < n; ++i) ld_global_256(x[i], p + 2 * i);
  if (A.log_n == 1) {
    bfly(x[0], x[1], f);
  } else if (A.log_n == 2) {
    bfly(x[0], x[2], f);
    bfly(x[1], x[3], f);
    mul_by(x[3], A.w4, f);
    bfly(x[0], x[1], f);
    bfly(x[2], x[3], f);
    // bit reversal of a 4-point DIF: swap positions 1 and 2
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      u32 tmp = x[1][k];
      x[1][k] = x[2][k];
      x[2][k] = tmp;
    }
  }
  for (int i = 0; i < n; ++i) {
    if (A.do_scale) mul_by(x[i], A.scale, f);
    canonicalize(x[i], f);
    st_global_256(p + 2 * i, x[i]);
  }
}

// out[i] = prod over set bits k of i of pw[k]   (pw[k] = w^(2^k) in Montgomery form), i.e. w^i.
// `first` offsets the exponent: out[i] = w^(first + i).
// shoup != 0: 64-byte entries (w plain, wq) instead of the 32-byte Montgomery form.
__global__ void twiddle_powers_kernel(uint4* __restrict__ out, const uint4* __restrict__ pw, int nbits,
                                      unsigned long long first, unsigned long long count,
                                      const __grid_constant__ FieldParams f, const __grid_constant__ Fe one,
                                      int shoup, const __grid_constant__ Fe pinv) {
  const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const unsigned long long e = first + i;
  u32 acc[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = one.v[k];
  for (int b = 0; b < nbits; ++b) {
    if ((e >> b) & 1ull) {
      u32 w[8];
      ld_table(w, pw, (u32)b);
      mul_by(acc, w, f);
    }
  }
  canonicalize(acc, f);
  if (!shoup) {
    st_global_256(out + 2 * i, acc);
    return;
  }
  u32 w[8], wq[8], unit[8] = {1, 0, 0, 0, 0, 0, 0, 0};
  mont_mul(w, acc, unit, f);  // (w R) * 1 / R = w; acc is canonical, a valid left operand
  canonicalize(w, f);
  shoup_companion(wq, acc, pinv.v);
  st_global_256(out + 4 * i, w);
  st_global_256(out + 4 * i + 2, wq);
}

// ---------------------------------------------------------------- element-wise companions
// (SURVEY.md par. 8f: the prover steps either side of an NTT -- pointwise product, Montgomery
// form conversion, coset shift).  One thread per element, 256-bit loads/stores.
//   mode 0: out = a*b      canonical in/out          (two products: (a*b/R) * R^2/R)
//   mode 1: out = a*b/R    Montgomery-form in/out    (one product)
__global__ void vec_mul_kernel(uint4* __restrict__ out, const uint4* __restrict__ a, const uint4* __restrict__ b,
                               unsigned long long n, int mode, const __grid_constant__ FieldParams f,
                               const __grid_constant__ Fe r2) {
  const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  u32 x[8], y[8];
  ld_global_256(x, a + 2 * i);
  ld_global_256(y, b + 2 * i);
  mul_by(y, x, f);  // x canonical (< p) is a valid left operand
  if (mode == 0) mul_by(y, r2.v, f);
  canonicalize(y, f);
  st_global_256(out + 2 * i, y);
}

// x <- c * x for a Montgomery-form constant c (R^2: to Montgomery form; 1: from Montgomery form)
__global__ void vec_scale_kernel(uint4* __restrict__ data, unsigned long long n, const __grid_constant__ FieldParams f,
                                 const __grid_constant__ Fe c) {
  const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  u32 x[8];
  ld_global_256(x, data + 2 * i);
  mul_by(x, c.v, f);
  canonicalize(x, f);
  st_global_256(data + 2 * i, x);
}

// x[j] <- s^j * x[j] with s^j = hi[j >> lo_bits] * lo[j & mask] (both tables in Montgomery form);
// j is the index inside one transform of 2^log_n points (batched arrays repeat the pattern).
__global__ void coset_scale_kernel(uint4* __restrict__ data, unsigned long long n, int log_n, int lo_bits,
                                   const uint4* __restrict__ hi, const uint4* __restrict__ lo,
                                   const __grid_constant__ FieldParams f) {
  const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const u32 j = (u32)(i & ((1ull << log_n) - 1ull));
  u32 x[8], w[8];
  ld_global_256(x, data + 2 * i);
  ld_table(w, hi, j >> lo_bits);
  mul_by(x, w, f);
  ld_table(w, lo, j & ((1u << lo_bits) - 1u));
  mul_by(x, w, f);
  canonicalize(x, f);
  st_global_256(data + 2 * i, x);
}

// table[j] <- s^j in Montgomery form, j < count: the two factored tables multiplied out once, so that
// the scaling pass of a coset transform is one product per element instead of two.
// shoup != 0: 64-byte (s^j, floor(s^j 2^256 / p)) entries, the format the pass kernels of a shoup context read.
__global__ void coset_table_kernel(uint4* __restrict__ table, unsigned long long count, int lo_bits,
                                   const uint4* __restrict__ hi, const uint4* __restrict__ lo,
                                   const __grid_constant__ FieldParams f, int shoup, const __grid_constant__ Fe pinv) {
  const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  u32 a[8], b[8];
  ld_table(a, hi, (u32)(i >> lo_bits));
  ld_table(b, lo, (u32)i & ((1u << lo_bits) - 1u));
  mul_by(a, b, f);  // (s^hi R)(s^lo R)/R = s^j R
  canonicalize(a, f);
  if (!shoup) {
    st_global_256(table + 2 * i, a);
    return;
  }
  u32 w[8], wq[8], unit[8] = {1, 0, 0, 0, 0, 0, 0, 0};
  mont_mul(w, a, unit, f);
  canonicalize(w, f);
  shoup_companion(wq, a, pinv.v);
  st_global_256(table + 4 * i, w);
  st_global_256(table + 4 * i + 2, wq);
}

// x[j] <- s^j * x[j] with s^j read from the multiplied-out table (one coalesced 256-bit read).
__global__ void coset_scale_direct_kernel(uint4* __restrict__ data, unsigned long long n, int log_n,
                                          const uint4* __restrict__ table, const __grid_constant__ FieldParams f,
                                          int shoup) {
  const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  u32 x[8], w[8], wq[8];
  ld_global_256(x, data + 2 * i);
  const unsigned long long j = i & ((1ull << log_n) - 1ull);
  if (shoup) {
    ld_global_256(w, table + 4 * j);
    ld_global_256(wq, table + 4 * j + 2);
    mul_tw_256<true>(x, w, wq, f);
    canonicalize_wl(x, f);
  } else {
    ld_global_256(w, table + 2 * j);
    mul_by(x, w, f);
    canonicalize(x, f);
  }
  st_global_256(data + 2 * i, x);
}

// out[i] = c * in[i]  (c Montgomery-form constant): scaled copies of small tables.
// shoup != 0: both tables hold 64-byte (w, wq) entries; c stays a Montgomery-form constant, and
// montmul(c R, w) = c*w is already the plain product.
__global__ void table_scale_kernel(uint4* __restrict__ out, const uint4* __restrict__ in, unsigned long long count,
                                   const __grid_constant__ FieldParams f, const __grid_constant__ Fe c, int shoup,
                                   const __grid_constant__ Fe r2, const __grid_constant__ Fe pinv) {
  const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  u32 x[8];
  ld_global_256(x, in + (shoup ? 4 : 2) * i);
  mul_by(x, c.v, f);
  canonicalize(x, f);
  if (shoup)
    st_shoup_entry(out + 4 * i, x, r2.v, pinv.v, f);
  else
    st_global_256(out + 2 * i, x);
}

}  // namespace ntt256
