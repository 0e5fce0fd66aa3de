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
#include "field_wl.cuh"
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
  // Companions fused into the end passes (SURVEY.md par. 8f row 2; the reference's bellperson-style kernel
  // premultiplies inside the pass as well, /root/reference/src/GZKP-NTT.cu:423-429).  Tables in the
  // context's twiddle format, indexed by the element's natural index inside one transform:
  //   in_scale : first pass, every loaded element is multiplied by in_scale[j]   (coset shift s^j)
  //   out_scale: last pass, every output by out_scale[k] before it is made canonical (s^-k of a coset inverse)
  //   out_mul  : last pass, every output by a caller-supplied Montgomery-form vector (pointwise product)
  const uint4* in_scale;
  const uint4* out_scale;
  const uint4* out_mul;
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

// NB DIF stages on the top NB bits of the GB-bit register index, with the constant twiddles of
// the 2^NB-point sub-transform (powers of w_8), followed -- unless this is the last round of the
// pass -- by the round twiddle w_{2^m}^(tau_lo * v) from the table.
template <int GB, int NB, bool SH>
NTT_DEV void round_compute(u32 (*x)[8], const PassArgs& A, const FieldParams& f, bool last, bool scaled, u32 lam0,
                           int sb, int lowp, bool tail_mul) {
#pragma unroll
  for (int k = 0; k < NB; ++k) {
    const int bit = GB - 1 - k;
#pragma unroll
    for (int u = 0; u < (1 << GB); ++u) {
      if (u & (1 << bit)) continue;
      const int t = (u >> (GB - NB)) & ((1 << (NB - 1 - k)) - 1);  // lower processed bits of u
      const int e8 = t << (3 - (NB - k));                           // twiddle as a power of w_8
      if constexpr (SH) {
        // the differences of the round's last stage meet the round / pass twiddle next (tail_mul), those with
        // a constant twiddle meet it right here; only the others must come back into the lazy range themselves
        bfly_wl(x[u], x[u | (1 << bit)], f, e8 != 0 || (k == NB - 1 && tail_mul));
      } else {
        bfly(x[u], x[u | (1 << bit)], f);
      }
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
// FZ: the launch carries element-wise companions (in_scale / out_scale / out_mul); plain transforms run the
// FZ = false instances, whose code does not know about them.
template <int GB, bool SH, bool FZ>
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
        if (FZ && A.in_scale != nullptr) {
          u32 w[8], wq[8];
          ld_tw<SH>(w, wq, A.in_scale, (u32)base_t + off);
          mul_tw<SH>(x[u], w, wq, f);
        }
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
      // every element with a processed bit set is multiplied after the round: by the round twiddle, or by the
      // pass twiddle / the 1/n of an inverse when this is the last round
      const bool tail_mul = !last || pd.has_tw || pd.scale_out;
      if (nb == 3) {
        if constexpr (GB >= 3) round_compute<GB, 3, SH>(x, A, f, last, scaled, lam0, sb, lowp, tail_mul);
      } else if (nb == 2) {
        if constexpr (GB >= 2) round_compute<GB, 2, SH>(x, A, f, last, scaled, lam0, sb, lowp, tail_mul);
      } else if (nb == 1) {
        round_compute<GB, 1, SH>(x, A, f, last, scaled, lam0, sb, lowp, tail_mul);
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
      if (FZ && (A.out_scale != nullptr || A.out_mul != nullptr)) {
#pragma unroll
        for (int u = 0; u < NE; ++u) {
          u32 off = (u32)base_t + g0;
#pragma unroll
          for (int i = 0; i < GB; ++i) off += ((u >> i) & 1) ? st[i] : 0u;
          u32 w[8], wq[8];
          if (A.out_scale != nullptr) {
            ld_tw<SH>(w, wq, A.out_scale, off);
            mul_tw<SH>(x[u], w, wq, f);
          }
          if (A.out_mul != nullptr) {
            // the vector is per call, not per transform: batches index it with the full offset
            ld_table(w, A.out_mul, (u32)(base - base_t) + off);
            canon<SH>(x[u], f);  // mont_mul wants its left operand below p
            mul_by_left(x[u], w, f);
          }
        }
      }
      if (pd.is_final) {
#pragma unroll
        for (int u = 0; u < NE; ++u) canon<SH>(x[u], f);
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
template <bool SH, bool FZ = false>
__global__ void __launch_bounds__(512, 1)
ntt_pass_kernel(const uint4* data_in, uint4* data_out, const __grid_constant__ PassArgs A) {
  extern __shared__ u32 smem[];
  ntt_pass_body<3, SH, FZ>(data_in, data_out, A, smem);
}

// Four elements per thread (radix-4 rounds): twice the warps, half the serial work per thread --
// used for small transforms, which are latency-bound with eight elements per thread.
template <bool SH, bool FZ = false>
__global__ void __launch_bounds__(512, 1)
ntt_pass4_kernel(const uint4* data_in, uint4* data_out, const __grid_constant__ PassArgs A) {
  extern __shared__ u32 smem[];
  ntt_pass_body<2, SH, FZ>(data_in, data_out, A, smem);
}

// The same body compiled for three resident CTAs of 256 threads (<= 85 registers; 80 used, 24 bytes spilled):
// 24 warps per SM instead of 16.  The path is bound by the latency of dependent wide multiplies, not by
// memory, so the extra warps pay for the extra round of shared-memory exchange (tiles of <= 2^10 elements).
template <bool SH, bool FZ = false>
__global__ void __launch_bounds__(256, 3)
ntt_pass4h_kernel(const uint4* data_in, uint4* data_out, const __grid_constant__ PassArgs A) {
  extern __shared__ u32 smem[];
  ntt_pass_body<2, SH, FZ>(data_in, data_out, A, smem);
}

// Two elements per thread (radix-2 rounds): the lowest-latency shape, for transforms so small that
// even the four-element kernel leaves most of the chip idle.
template <bool SH, bool FZ = false>
__global__ void __launch_bounds__(1024, 1)
ntt_pass2_kernel(const uint4* data_in, uint4* data_out, const __grid_constant__ PassArgs A) {
  extern __shared__ u32 smem[];
  ntt_pass_body<1, SH, FZ>(data_in, data_out, A, smem);
}

}  // namespace ntt256
