// Device kernels of the self-sorting in-place 256-bit NTT (sm_100a).
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
#include "kernel_common.cuh"
#include "plan.h"

namespace ntt256 {

struct PassArgs {
  FieldParams f;
  PassDesc pd;
  const uint4* om_hi;  // w_N^(i << lo_bits), i < 2^hi_bits, Montgomery form, canonical
  const uint4* om_lo;  // w_N^i, i < 2^lo_bits
  const uint4* loc;    // w_{2^loc_bits}^i, i < 2^loc_bits (round twiddles; loc_bits = min(12, max_log_n))
  int hi_bits, lo_bits;  // hi_bits + lo_bits = max_log_n
  int loc_bits;
  int inverse;         // 1: exponents are negated (w^-1 tables are the same tables read backwards)
  u32 c8[3][8];        // w_8^1, w_8^2, w_8^3 (or their inverses), Montgomery form
  u32 scale[8];        // 1/n in Montgomery form (used when pd.scale_out)
  unsigned long long batch_stride;  // elements between consecutive transforms of a batch
  // Fused exchange (multi-GPU four-step, last pass of step 1): when peer_bits > 0 the tile is
  // not stored back in place but straight into the receive buffers of the destination ranks
  // over NVLink (peer-mapped pointers): destination = top peer_bits of the local output index,
  // remote index = (my rank << peer_shift) | (low peer_shift bits).  The all-to-all becomes
  // the store phase of the butterfly kernel.
  unsigned long long peer[8];
  int peer_bits, peer_shift, peer_rank;
  // Inverse NTT: round `scaled_round` of this pass reads its twiddles from loc_scaled = (1/n)*loc
  // and multiplies all 8 elements (also the one whose twiddle is 1), which applies the 1/n.
  int scaled_round;
  const uint4* loc_scaled;
  // Streamed pass twiddles.  A pass whose twiddle exponent does not fit one table lookup (the first
  // pass of a transform longer than 2^hi_bits) would need two lookups and an extra product per
  // element; instead its n twiddles can be kept in an array laid out exactly like the pass's
  // output (read with the same coalesced 256-bit accesses as the stores): HBM has ~6x headroom
  // on this path, the integer pipe none.  tw_gen_out != null: the launch only WRITES that array.
  const uint4* tw_stream;
  uint4* tw_gen_out;
  u32 one[8];  // the field's 1 as the twiddle chain starts from it: R mod p (Montgomery tables) or 1 (Shoup tables)
  // Shoup twiddles (kernels instantiated with SH = true; fields with p < 2^254, see field_shoup.cuh): every
  // table entry is 64 bytes, the constant w (plain, canonical) followed by wq = floor(w * 2^256 / p), and the
  // constants passed by value carry their companions too.
  u32 c8q[3][8];
  u32 scaleq[8];
  u32 r2[8];    // R^2 mod p: plain -> Montgomery form (only to derive companions when a table is generated)
  u32 pinv[8];  // p^-1 mod 2^256
  // Rank synchronisation folded into the passes (multi-GPU four-step, one process per GPU).  sync_mode bit 1
  // (arrive, last pass of step 1): when the grid's last CTA has finished its peer stores it publishes
  // sync_epoch in every rank's flag buffer.  Bit 0 (wait, first pass of step 2): every CTA first waits
  // until all ranks' epochs have arrived in the local flag buffer.  No separate barrier kernel, no
  // collective; the step-2 kernel is already resident while the last peer stores land.
  unsigned long long sync_flags[8];   // flag buffer of every rank, as mapped in this process
  unsigned long long sync_epoch;
  unsigned long long sync_timeout_ns;
  unsigned int* sync_counter;         // CTA arrival counter of this context (device memory, zero between uses)
  int sync_mode, sync_nranks, sync_rank;
};

constexpr int kSyncSlotBase = 256 / 8;  // epochs start at byte 256 of a flag buffer (u64 slots, one per source rank)
constexpr int kSyncErrSlot = 128 / 8;   // a waiter that timed out leaves the epoch it was waiting for here

NTT_DEV unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

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

// ---------------------------------------------------------------- twiddles in either representation
// SH = false: 32-byte entries, w in Montgomery form, product = mont_mul (128 wide + 8 low multiplies).
// SH = true : 64-byte entries (w, wq), product = shoup_mul (99 wide + 16 low multiplies).
template <bool SH>
NTT_DEV void ld_tw(u32* w, u32* q, const uint4* tab, u32 idx) {
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
NTT_DEV void mul_tw(u32* x, const u32* w, const u32* q, const FieldParams& f) {
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

// NB DIF stages on the top NB bits of the GB-bit register index, with the constant twiddles of
// the 2^NB-point sub-transform (powers of w_8), followed -- unless this is the last round of the
// pass -- by the round twiddle w_{2^m}^(tau_lo * v) from the table.
template <int GB, int NB, bool SH>
NTT_DEV void round_compute(u32 (*x)[8], const PassArgs& A, const FieldParams& f, bool last, bool scaled, u32 lam0,
                           int sb, int lowp) {
#pragma unroll
  for (int k = 0; k < NB; ++k) {
    const int bit = GB - 1 - k;
#pragma unroll
    for (int u = 0; u < (1 << GB); ++u) {
      if (u & (1 << bit)) continue;
      bfly(x[u], x[u | (1 << bit)], f);
      const int t = (u >> (GB - NB)) & ((1 << (NB - 1 - k)) - 1);  // lower processed bits of u
      const int e8 = t << (3 - (NB - k));                           // twiddle as a power of w_8
      if (e8) mul_tw<SH>(x[u | (1 << bit)], A.c8[e8 - 1], A.c8q[e8 - 1], f);
    }
  }
  if (last || NB == 0) return;
  const int pb = sb + GB - NB;  // lowest bit butterflied in this round
  const int m = sb + GB - lowp; // process bits from this group's top down
  const int sh = A.loc_bits - m;
  const u32 tl_mask = (1u << (pb - lowp)) - 1u;
  const u32 tl0 = (lam0 >> lowp) & tl_mask;
  const int ush = sb - lowp;    // >= 0 here (only the last round can dip below)
  const uint4* loc_tab = scaled ? A.loc_scaled : A.loc;
#pragma unroll
  for (int u = 0; u < (1 << GB); ++u) {
    // v = bit-reversed processed bits of u; tau_lo includes the group's unprocessed bits
    const int uproc = u >> (GB - NB);
    int v = 0;
#pragma unroll
    for (int i = 0; i < NB; ++i) v |= ((uproc >> i) & 1) << (NB - 1 - i);
    const u32 rest = u & ((1 << (GB - NB)) - 1);
    if (scaled || uproc != 0) {
      const u32 tl = (tl0 | (rest << ush)) & tl_mask;
      u32 e = (tl * (u32)v) << sh;
      if (A.inverse) e = (0u - e);
      e &= (1u << A.loc_bits) - 1u;
      u32 w[8], wq[8];
      ld_tw<SH>(w, wq, loc_tab, e);
      mul_tw<SH>(x[u], w, wq, f);
    }
  }
}

// One pass.  GB = log2(elements per thread): 3 is the throughput configuration (radix-8 rounds,
// 64 data registers); 2 halves the serial work per thread and doubles the number of warps, which
// is what small, latency-bound transforms need.
template <int GB, bool SH>
NTT_DEV void ntt_pass_body(const uint4* data_in, uint4* data_out, const PassArgs& A, u32* smem) {
  constexpr int NE = 1 << GB;
  const PassDesc& pd = A.pd;
  const FieldParams& f = A.f;
  const int T = pd.T;
  const int d = pd.d;
  const u32 tid = threadIdx.x;
  const u32 tile_words = 1u << T;

  // ---- block-level index pieces (uniform)
  const unsigned long long blk = blockIdx.x;
  const unsigned long long base_t = deposit64(blk, pd.blk_pos, pd.nblk_bits);  // index inside one transform
  unsigned long long base = base_t;
  u32 jlow_blk = pd.jlow_const;
  for (int i = 0; i < pd.nblk_bits; ++i) jlow_blk += (u32)((blk >> i) & 1ull) * pd.wt_blk[i];
  base += (unsigned long long)blockIdx.y * A.batch_stride;
  const uint4* gbase_in = data_in + 2 * base;
  uint4* gbase = data_out + 2 * base;

  u32 x[NE][8];
  const bool gen = (A.tw_gen_out != nullptr);
  const int lowp = T - d;  // lowest process bit

  if (A.sync_mode & 1) {
    // every rank's step 1 (and its peer stores into this rank's memory) must be complete
    if (tid < (u32)A.sync_nranks) {
      volatile unsigned long long* flags = (volatile unsigned long long*)A.sync_flags[A.sync_rank];
      const unsigned long long t0 = global_timer_ns();
      while (flags[kSyncSlotBase + tid] < A.sync_epoch) {
        if (global_timer_ns() - t0 > A.sync_timeout_ns) {  // a peer died: record it, do not hang the GPU
          flags[kSyncErrSlot] = A.sync_epoch;
          break;
        }
      }
    }
    __threadfence_system();
    __syncthreads();
  }

  for (int r = gen ? pd.nrounds - 1 : 0; r < pd.nrounds; ++r) {
    const int sb = pd.round_sb[r];
    const int nb = pd.round_nb[r];
    const u32 lam0 = deposit32(tid, pd.tmap[r], T - GB);
    const bool first = (r == 0);
    const bool last = (r == pd.nrounds - 1);

    // ---- fetch the elements of this thread
    if (gen) {
#pragma unroll
      for (int u = 0; u < NE; ++u)
#pragma unroll
        for (int k = 0; k < 8; ++k) x[u][k] = A.one[k];
    } else if (first) {
      const u32 g0 = deposit32(lam0, pd.in_pos, T);
      u32 st[GB];
#pragma unroll
      for (int i = 0; i < GB; ++i) st[i] = 1u << pd.in_pos[sb + i];
#pragma unroll
      for (int u = 0; u < NE; ++u) {
        u32 off = g0;
#pragma unroll
        for (int i = 0; i < GB; ++i) off += ((u >> i) & 1) ? st[i] : 0u;
        ld_global_256(x[u], gbase_in + 2 * (size_t)off);
      }
    } else {
      const u32 w0 = swz(lam0);
#pragma unroll
      for (int u = 0; u < NE; ++u) {
        const u32 w = w0 ^ swz((u32)u << sb);
#pragma unroll
        for (int k = 0; k < 8; ++k) x[u][k] = smem[k * tile_words + w];
      }
    }

    // ---- butterflies on the top nb bits of the group (+ round twiddle unless last)
    if (!gen) {
      const bool scaled = (r == A.scaled_round);
      if (nb == 3) {
        if constexpr (GB >= 3) round_compute<GB, 3, SH>(x, A, f, last, scaled, lam0, sb, lowp);
      } else if (nb == 2) {
        if constexpr (GB >= 2) round_compute<GB, 2, SH>(x, A, f, last, scaled, lam0, sb, lowp);
      } else if (nb == 1) {
        round_compute<GB, 1, SH>(x, A, f, last, scaled, lam0, sb, lowp);
      }
    }

    if (!last) {
      // ---- exchange
      const u32 w0 = swz(lam0);
#pragma unroll
      for (int u = 0; u < NE; ++u) {
        const u32 w = w0 ^ swz((u32)u << sb);
#pragma unroll
        for (int k = 0; k < 8; ++k) smem[k * tile_words + w] = x[u][k];
      }
      __syncthreads();
    } else {
      // ---- pass twiddle, scale, canonical form, store
      const u32 g0 = deposit32(lam0, pd.out_pos, T);
      u32 st[GB];
#pragma unroll
      for (int i = 0; i < GB; ++i) st[i] = 1u << pd.out_pos[sb + i];
      if (pd.has_tw && A.tw_stream != nullptr && !gen) {
#pragma unroll
        for (int u = 0; u < NE; ++u) {
          u32 off = (u32)base_t + g0;
#pragma unroll
          for (int i = 0; i < GB; ++i) off += ((u >> i) & 1) ? st[i] : 0u;
          u32 w[8], wq[8];
          ld_tw<SH>(w, wq, A.tw_stream, off);
          mul_tw<SH>(x[u], w, wq, f);
        }
      } else if (pd.has_tw) {
        u32 jl0 = jlow_blk;
        for (int i = 0; i < lowp; ++i) jl0 += ((lam0 >> i) & 1u) * pd.wt_loc[i];
        u32 gw[GB];
#pragma unroll
        for (int i = 0; i < GB; ++i) gw[i] = (sb + i < lowp) ? pd.wt_loc[sb + i] : 0u;
        const u32 nmask = (A.hi_bits + A.lo_bits >= 32) ? 0xffffffffu : ((1u << (A.hi_bits + A.lo_bits)) - 1u);
        const u32 lomask = (1u << A.lo_bits) - 1u;
        const bool two = pd.tw_lshift < A.lo_bits;
#pragma unroll
        for (int u = 0; u < NE; ++u) {
          const u32 lam = lam0 | ((u32)u << sb);
          const u32 khi = d ? (__brev(lam >> lowp) >> (32 - d)) : 0u;
          u32 jl = jl0;
#pragma unroll
          for (int i = 0; i < GB; ++i) jl += ((u >> i) & 1) ? gw[i] : 0u;
          u32 e = (jl * khi) << pd.tw_lshift;
          if (A.inverse) e = (0u - e);
          e &= nmask;
          u32 w[8], wq[8];
          ld_tw<SH>(w, wq, A.om_hi, e >> A.lo_bits);
          mul_tw<SH>(x[u], w, wq, f);
          if (two) {
            ld_tw<SH>(w, wq, A.om_lo, e & lomask);
            mul_tw<SH>(x[u], w, wq, f);
          }
        }
      }
      if (pd.scale_out) {
#pragma unroll
        for (int u = 0; u < NE; ++u) mul_tw<SH>(x[u], A.scale, A.scaleq, f);
      }
      if (pd.is_final) {
#pragma unroll
        for (int u = 0; u < NE; ++u) canonicalize(x[u], f);
      }
      // a single-round pass loads and stores global memory in the same round: every thread of
      // the CTA must have loaded before any permuted store lands on another thread's input
      if (first) __syncthreads();
      if (gen) {
#pragma unroll
        for (int u = 0; u < NE; ++u) {
          u32 off = (u32)base_t + g0;
#pragma unroll
          for (int i = 0; i < GB; ++i) off += ((u >> i) & 1) ? st[i] : 0u;
          canonicalize(x[u], f);
          if constexpr (SH)
            st_shoup_entry(A.tw_gen_out + 4 * (size_t)off, x[u], A.r2, A.pinv, f);
          else
            st_global_256(A.tw_gen_out + 2 * (size_t)off, x[u]);
        }
      } else if (A.peer_bits == 0) {
#pragma unroll
        for (int u = 0; u < NE; ++u) {
          u32 off = g0;
#pragma unroll
          for (int i = 0; i < GB; ++i) off += ((u >> i) & 1) ? st[i] : 0u;
          st_global_256(gbase + 2 * (size_t)off, x[u]);
        }
      } else {
        const u32 lowmask = (1u << A.peer_shift) - 1u;
        const u32 mine = (u32)A.peer_rank << A.peer_shift;
#pragma unroll
        for (int u = 0; u < NE; ++u) {
          // full local output index: block part + in-tile part
          u32 off = (u32)base + g0;
#pragma unroll
          for (int i = 0; i < GB; ++i) off += ((u >> i) & 1) ? st[i] : 0u;
          uint4* dst = reinterpret_cast<uint4*>(A.peer[off >> A.peer_shift]) + 2 * (size_t)(mine | (off & lowmask));
          st_global_256(dst, x[u]);
        }
      }
    }
  }
  if (A.sync_mode & 2) {
    // this CTA's stores (peer memory included) are visible system-wide before it counts as arrived;
    // the last CTA of the grid tells every rank
    __threadfence_system();
    __syncthreads();
    if (tid == 0) {
      const unsigned int total = gridDim.x * gridDim.y;
      if (atomicAdd(A.sync_counter, 1u) == total - 1u) {
        *A.sync_counter = 0u;
        __threadfence_system();
        for (int r = 0; r < A.sync_nranks; ++r)
          ((volatile unsigned long long*)A.sync_flags[r])[kSyncSlotBase + A.sync_rank] = A.sync_epoch;
      }
    }
  }
}

// data_in == data_out: in place (the normal case).  They differ when the first / last pass of
// a host-buffer transform reads / writes pinned host memory directly (zero-copy over PCIe, so the
// transfer overlaps the pass instead of preceding / following it).
template <bool SH>
__global__ void __launch_bounds__(512, 1)
ntt_pass_kernel(const uint4* data_in, uint4* data_out, const __grid_constant__ PassArgs A) {
  extern __shared__ u32 smem[];
  ntt_pass_body<3, SH>(data_in, data_out, A, smem);
}

// Four elements per thread (radix-4 rounds): twice the warps, half the serial work per thread --
// used for small transforms, which are latency-bound with eight elements per thread.
template <bool SH>
__global__ void __launch_bounds__(512, 1)
ntt_pass4_kernel(const uint4* data_in, uint4* data_out, const __grid_constant__ PassArgs A) {
  extern __shared__ u32 smem[];
  ntt_pass_body<2, SH>(data_in, data_out, A, smem);
}

// Two elements per thread (radix-2 rounds): the lowest-latency shape, for transforms so small that
// even the four-element kernel leaves most of the chip idle.
template <bool SH>
__global__ void __launch_bounds__(1024, 1)
ntt_pass2_kernel(const uint4* data_in, uint4* data_out, const __grid_constant__ PassArgs A) {
  extern __shared__ u32 smem[];
  ntt_pass_body<1, SH>(data_in, data_out, A, smem);
}

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
__global__ void coset_table_kernel(uint4* __restrict__ table, unsigned long long count, int lo_bits,
                                   const uint4* __restrict__ hi, const uint4* __restrict__ lo,
                                   const __grid_constant__ FieldParams f) {
  const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  u32 a[8], b[8];
  ld_table(a, hi, (u32)(i >> lo_bits));
  ld_table(b, lo, (u32)i & ((1u << lo_bits) - 1u));
  mul_by(a, b, f);  // (s^hi R)(s^lo R)/R = s^j R
  canonicalize(a, f);
  st_global_256(table + 2 * i, a);
}

// x[j] <- s^j * x[j] with s^j read from the multiplied-out table (one coalesced 256-bit read).
__global__ void coset_scale_direct_kernel(uint4* __restrict__ data, unsigned long long n, int log_n,
                                          const uint4* __restrict__ table, const __grid_constant__ FieldParams f) {
  const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  u32 x[8], w[8];
  ld_global_256(x, data + 2 * i);
  ld_global_256(w, table + 2 * (i & ((1ull << log_n) - 1ull)));
  mul_by(x, w, f);
  canonicalize(x, f);
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
