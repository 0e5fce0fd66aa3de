// Device kernels of the self-sorting in-place 256-bit NTT (sm_100a): the 8-limb field policies (F256<SH>:
// Montgomery or Shoup twiddles) of the one pass body in ntt_pass_generic.cuh, and their kernel instances.
//
//   ntt_pass_kernel   one pass of the plan (plan.h): a CTA owns a tile of 2^T elements picked
//                     by T index bits, each thread keeps 8 elements in registers, runs up to
//                     three DIF stages per register round, exchanges through a limb-major,
//                     XOR-swizzled shared-memory tile between rounds, multiplies by the pass
//                     twiddle and stores the tile back with its index bits permuted.
//                     Replaces rearrange + naive + GZKP of /root/reference/src/big-num.cu:57-258.
//   ntt_tiny_kernel   transforms of 2 or 4 points (one thread per transform).
//   twiddle_powers_kernel   builds w^i tables on the device (the reference builds an n-entry
//                     table serially on the host at every call, big-num.cu:298-309).
//
// Global memory is touched with one 256-bit LDG/STG per element (LDG.E.ENL2.256); a warp reads
// runs of >= 2^g consecutive elements (g = plan coalesce_bits).
#pragma once
#include <cuda_runtime.h>

#include "field.cuh"
#include "field_shoup.cuh"
#include "field_wl.cuh"
#include "kernel_common.cuh"
#include "ntt_pass_generic.cuh"
#include "plan.h"

namespace ntt256 {

// ---------------------------------------------------------------- element I/O
NTT_DEV void ld_global_256(u32* v, const uint4* p) {
  asm volatile("ld.global.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "l"(p));
}
NTT_DEV void st_global_256(uint4* p, const u32* v) {
  asm volatile("st.global.v8.u32 [%8], {%0,%1,%2,%3,%4,%5,%6,%7};" ::"r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]),
               "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "l"(p)
               : "memory");
}
// read-only twiddle tables: non-coherent path, two 128-bit loads
NTT_DEV void ld_table(u32* v, const uint4* tab, u32 idx) {
  const uint4 a = __ldg(tab + 2 * (size_t)idx);
  const uint4 b = __ldg(tab + 2 * (size_t)idx + 1);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
  v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}

// x <- w * x  (w canonical Montgomery-form constant, x lazy)
NTT_DEV void mul_by(u32* x, const u32* w, const FieldParams& f) {
  u32 r[8];
  mont_mul(r, w, x, f);
#pragma unroll
  for (int k = 0; k < 8; ++k) x[k] = r[k];
}

// x <- x * b / R for a lazy b (any value below 2^256) and a canonical x: the product of two variable
// operands, e.g. transform outputs times a caller's Montgomery-form vector (= the plain product).
NTT_DEV void mul_by_left(u32* x, const u32* b, const FieldParams& f) {
  u32 r[8];
  mont_mul(r, x, b, f);
#pragma unroll
  for (int k = 0; k < 8; ++k) x[k] = r[k];
}

// ---------------------------------------------------------------- twiddles in either representation
// SH = false: 32-byte entries, w in Montgomery form, product = mont_mul (128 wide + 8 low multiplies).
// SH = true : 64-byte entries (w, wq), product = shoup_mul (99 wide + 16 low multiplies).
template <bool SH>
NTT_DEV void ld_tw_256(u32* w, u32* q, const uint4* tab, u32 idx) {
  if constexpr (SH) {
    const uint4* e = tab + 4 * (size_t)idx;
    const uint4 a = __ldg(e), b = __ldg(e + 1), c = __ldg(e + 2), d = __ldg(e + 3);
    w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w;
    w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
    q[0] = c.x; q[1] = c.y; q[2] = c.z; q[3] = c.w;
    q[4] = d.x; q[5] = d.y; q[6] = d.z; q[7] = d.w;
  } else {
    ld_table(w, tab, idx);
  }
}
template <bool SH>
NTT_DEV void mul_tw_256(u32* x, const u32* w, const u32* q, const FieldParams& f) {
  u32 r[8];
  if constexpr (SH) shoup_mul(r, x, w, q, f); else mont_mul(r, w, x, f);
#pragma unroll
  for (int k = 0; k < 8; ++k) x[k] = r[k];
}

// r = low 256 bits of a * b (table generation only)
NTT_DEV void mul_lo256(u32* r, const u32* a, const u32* b) {
#pragma unroll
  for (int k = 0; k < 8; ++k) r[k] = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    u32 carry = 0;
#pragma unroll
    for (int j = 0; j + i < 8; ++j) {
      const u64 t = (u64)a[i] * b[j] + r[i + j] + carry;
      r[i + j] = (u32)t;
      carry = (u32)(t >> 32);
    }
  }
}
// wq = floor(w * 2^256 / p) from wm = w * 2^256 mod p (canonical; it is the Montgomery form of w):
// w * 2^256 - wm is an exact multiple of p, so the quotient is (-wm) * p^-1 modulo 2^256.
NTT_DEV void shoup_companion(u32* wq, const u32* wm, const u32* pinv) {
  u32 neg[8], zero[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) zero[k] = 0;
  (void)sub256(neg, zero, wm);
  mul_lo256(wq, neg, pinv);
}
// 64-byte Shoup entry (w, wq) from a plain canonical w: wm = w * R via one Montgomery product with R^2.
NTT_DEV void st_shoup_entry(uint4* dst, const u32* w, const u32* r2, const u32* pinv, const FieldParams& f) {
  u32 wm[8], wq[8];
  mont_mul(wm, r2, w, f);
  canonicalize(wm, f);
  shoup_companion(wq, wm, pinv);
  st_global_256(dst, w);
  st_global_256(dst + 2, wq);
}

// DIF butterfly: (a, b) <- (a + b, a - b)
NTT_DEV void bfly(u32* a, u32* b, const FieldParams& f) {
  u32 s[8], d[8];
  add_lazy(s, a, b, f);
  sub_lazy(d, a, b, f);
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    a[k] = s[k];
    b[k] = d[k];
  }
}

// The same under the wide-lazy policy of Shoup contexts (field_wl.cuh).  to_mul: the difference is multiplied
// by a twiddle next (any 256-bit value is a valid operand of shoup_mul), so it only needs to be non-negative.
NTT_DEV void bfly_wl(u32* a, u32* b, const FieldParams& f, bool to_mul) {
  u32 s[8], d[8];
  add_wl(s, a, b, f);
  if (to_mul) sub_wl_mul(d, a, b, f); else sub_wl(d, a, b, f);
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    a[k] = s[k];
    b[k] = d[k];
  }
}

// [0, 2p) -> [0, p) (Montgomery contexts) / anything the wide-lazy policy produces -> [0, p) (Shoup contexts)
template <bool SH>
NTT_DEV void canon(u32* x, const FieldParams& f) {
  if constexpr (SH) canonicalize_wl(x, f); else canonicalize(x, f);
}


// ---------------------------------------------------------------- the 8-limb field policies of the pass body
using PassArgs = PassArgsT<FieldParams, 8, 16>;

template <bool SH>
struct F256 {
  static constexpr int NL = 8;
  static constexpr int TW = SH ? 16 : 8;
  using Params = FieldParams;
  using Args = PassArgs;
  static NTT_DEV const FieldParams& params(const Args& A) { return A.f; }
  static NTT_DEV void ld(u32* v, const uint4* p) { ld_global_256(v, p); }
  static NTT_DEV void st(uint4* p, const u32* v) { st_global_256(p, v); }
  static NTT_DEV void ld_tw(u32* w, const uint4* tab, u32 idx) { ld_tw_256<SH>(w, w + 8, tab, idx); }
  static NTT_DEV void mul_tw(u32* x, const u32* w, const FieldParams& f) { mul_tw_256<SH>(x, w, w + 8, f); }
  static NTT_DEV void bfly(u32* a, u32* b, const FieldParams& f, bool to_mul) {
    if constexpr (SH) bfly_wl(a, b, f, to_mul); else ntt256::bfly(a, b, f);
  }
  static NTT_DEV void canon(u32* x, const FieldParams& f) { ntt256::canon<SH>(x, f); }
  static NTT_DEV void canon_exact(u32* x, const FieldParams& f) { canonicalize(x, f); }
  static NTT_DEV void st_tw(uint4* tab, u32 idx, const u32* w, const Args& A) {
    if constexpr (SH) st_shoup_entry(tab + 4 * (size_t)idx, w, A.r2, A.pinv, A.f);
    else st_global_256(tab + 2 * (size_t)idx, w);
  }
  // x <- x * v[idx] / R for a caller-supplied Montgomery-form vector (= the plain product)
  static NTT_DEV void mul_var(u32* x, const uint4* vec, u32 idx, const FieldParams& f) {
    u32 w[8];
    ld_table(w, vec, idx);
    ntt256::canon<SH>(x, f);  // mont_mul wants its left operand below p
    mul_by_left(x, w, f);
  }
};

// data_in == data_out: in place (the normal case).  They differ when the first / last pass of
// a host-buffer transform reads / writes pinned host memory directly (zero-copy over PCIe, so the
// transfer overlaps the pass instead of preceding / following it).
template <bool SH, bool FZ = false>
__global__ void __launch_bounds__(512, 1)
ntt_pass_kernel(const uint4* data_in, uint4* data_out, const __grid_constant__ PassArgs A) {
  extern __shared__ u32 smem[];
  ntt_pass_body<3, F256<SH>, FZ>(data_in, data_out, A, smem);
}

// Four elements per thread (radix-4 rounds): twice the warps, half the serial work per thread --
// used for small transforms, which are latency-bound with eight elements per thread.
template <bool SH, bool FZ = false>
__global__ void __launch_bounds__(512, 1)
ntt_pass4_kernel(const uint4* data_in, uint4* data_out, const __grid_constant__ PassArgs A) {
  extern __shared__ u32 smem[];
  ntt_pass_body<2, F256<SH>, FZ>(data_in, data_out, A, smem);
}

// The same body compiled for three resident CTAs of 256 threads (<= 85 registers; 80 used, 24 bytes spilled):
// 24 warps per SM instead of 16.  The path is bound by the latency of dependent wide multiplies, not by
// memory, so the extra warps pay for the extra round of shared-memory exchange (tiles of <= 2^10 elements).
template <bool SH, bool FZ = false>
__global__ void __launch_bounds__(256, 3)
ntt_pass4h_kernel(const uint4* data_in, uint4* data_out, const __grid_constant__ PassArgs A) {
  extern __shared__ u32 smem[];
  ntt_pass_body<2, F256<SH>, FZ>(data_in, data_out, A, smem);
}

// Two elements per thread (radix-2 rounds): the lowest-latency shape, for transforms so small that
// even the four-element kernel leaves most of the chip idle.
template <bool SH, bool FZ = false>
__global__ void __launch_bounds__(1024, 1)
ntt_pass2_kernel(const uint4* data_in, uint4* data_out, const __grid_constant__ PassArgs A) {
  extern __shared__ u32 smem[];
  ntt_pass_body<1, F256<SH>, FZ>(data_in, data_out, A, smem);
}

}  // namespace ntt256
