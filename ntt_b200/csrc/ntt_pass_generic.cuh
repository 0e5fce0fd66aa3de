// The pass body shared by every limb count (SURVEY.md par. 8f row 4: the reference's kernels are templates over the
// container width, /root/reference/src/big-num.cu:57,67,172,260).  A field policy F supplies the element layout and
// the arithmetic; the pass / round / exchange / twiddle / permuted-store structure below exists once:
//
//   F::NL        u32 limbs per element (8: 256-bit containers, 12: 384-bit ...)
//   F::TW        u32 words per twiddle-table entry (NL for Montgomery tables, 2 NL for Shoup pairs)
//   F::Args      the PassArgsT instance the kernels take by value;   F::params(A)   the field constants (A.f, or
//                compile-time constants for an instance specialised on one modulus)
//   F::ld / st   one element from / to global memory
//   F::ld_tw     table entry idx -> TW words;   F::mul_tw   x <- w * x
//   F::bfly      (a, b) <- (a + b, a - b) under the policy's lazy ranges (to_mul: the difference meets a product next)
//   F::canon     lazy -> [0, p);   F::canon_exact   [0, 2p) -> [0, p) (table generation)
//   F::st_tw     write a generated table entry;   F::mul_var   x <- x * v[idx] for a caller-supplied vector
//
// Instances: F256<SH> (ntt_kernels.cuh: Montgomery or Shoup twiddles over 8 limbs) and FN<NL> (ntt_kernels_n.cuh).
#pragma once
#include <cuda_runtime.h>

#include "kernel_common.cuh"
#include "plan.h"

namespace ntt256 {

// NL limbs per element; by-value twiddle constants are stored in TWS words (w, then its Shoup companion when the
// policy uses one).
template <class PARAMS, int NL, int TWS>
struct PassArgsT {
  PARAMS f;
  PassDesc pd;
  const uint4* om_hi;  // w_N^(i << lo_bits), i < 2^hi_bits, Montgomery form, canonical
  const uint4* om_lo;  // w_N^i, i < 2^lo_bits
  const uint4* loc;    // w_{2^loc_bits}^i, i < 2^loc_bits (round twiddles; loc_bits = min(12, max_log_n))
  int hi_bits, lo_bits;  // hi_bits + lo_bits = max_log_n
  int loc_bits;
  int inverse;         // 1: exponents are negated (w^-1 tables are the same tables read backwards)
  u32 c8[3][TWS];      // w_8^1, w_8^2, w_8^3 (or their inverses) as table entries
  u32 scale[TWS];      // 1/n as a table entry (used when pd.scale_out)
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
  u32 one[NL];  // the field's 1 as the twiddle chain starts from it: R mod p (Montgomery tables) or 1 (Shoup tables)
  // Shoup twiddles (8-limb kernels instantiated with SH = true; fields with p < 2^254, see field_shoup.cuh): every
  // table entry is 64 bytes, the constant w (plain, canonical) followed by wq = floor(w * 2^256 / p), and the
  // constants passed by value (c8, scale) carry their companions in words NL .. 2 NL - 1.
  u32 r2[NL];    // R^2 mod p: plain -> Montgomery form (only to derive companions when a table is generated)
  u32 pinv[NL];  // p^-1 mod 2^(32 NL)
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
  // Index tables of this pass, built on the host with the plan (pass_tables.h): per thread and round the local element
  // index and its swizzled shared-memory offset, the global offsets of the thread's first element when loading /
  // storing, its share of j_low; per CTA the tile base and its share of j_low.
  const u32* thr;
  const uint2* blk;
};

constexpr int kSyncSlotBase = 256 / 8;  // epochs start at byte 256 of a flag buffer (u64 slots, one per source rank)
constexpr int kSyncErrSlot = 128 / 8;   // a waiter that timed out leaves the epoch it was waiting for here

NTT_DEV unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// NB DIF stages on the top NB bits of the GB-bit register index, with the constant twiddles of
// the 2^NB-point sub-transform (powers of w_8), followed -- unless this is the last round of the
// pass -- by the round twiddle w_{2^m}^(tau_lo * v) from the table.
template <int GB, int NB, class F>
NTT_DEV void round_compute(u32 (*x)[F::NL], const typename F::Args& A, const typename F::Params& f, bool last, bool scaled, u32 lam0,
                           int sb, int lowp, bool tail_mul) {
#pragma unroll
  for (int k = 0; k < NB; ++k) {
    const int bit = GB - 1 - k;
#pragma unroll
    for (int u = 0; u < (1 << GB); ++u) {
      if (u & (1 << bit)) continue;
      const int t = (u >> (GB - NB)) & ((1 << (NB - 1 - k)) - 1);  // lower processed bits of u
      const int e8 = t << (3 - (NB - k));                           // twiddle as a power of w_8
      // the differences of the round's last stage meet the round / pass twiddle next (tail_mul), those with a
      // constant twiddle meet it right here; policies with wide lazy ranges only bring the others back themselves
      F::bfly(x[u], x[u | (1 << bit)], f, e8 != 0 || (k == NB - 1 && tail_mul));
      if (e8) F::mul_tw(x[u | (1 << bit)], A.c8[e8 - 1], f);
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
  const u32 loc_mask = (1u << A.loc_bits) - 1u;
  // the table index of element u is (tau_lo * v_u) << sh, negated for the inverse: the shift and the sign are taken
  // once per round (on tau_lo), which leaves one multiply by the constant v_u and one mask per element
  u32 tbase[1 << (GB - NB)];
#pragma unroll
  for (int rst = 0; rst < (1 << (GB - NB)); ++rst) {
    const u32 t0 = ((tl0 | ((u32)rst << ush)) & tl_mask) << sh;
    tbase[rst] = A.inverse ? (0u - t0) : t0;
  }
#pragma unroll
  for (int u = 0; u < (1 << GB); ++u) {
    // v = bit-reversed processed bits of u; tau_lo includes the group's unprocessed bits
    const int uproc = u >> (GB - NB);
    int v = 0;
#pragma unroll
    for (int i = 0; i < NB; ++i) v |= ((uproc >> i) & 1) << (NB - 1 - i);
    const u32 rest = u & ((1 << (GB - NB)) - 1);
    if (scaled || uproc != 0) {
      const u32 e = (tbase[rest] * (u32)v) & loc_mask;
      u32 w[F::TW];
      F::ld_tw(w, loc_tab, e);
      F::mul_tw(x[u], w, f);
    }
  }
}

// One pass.  GB = log2(elements per thread): 3 is the throughput configuration (radix-8 rounds,
// 64 data registers); 2 halves the serial work per thread and doubles the number of warps, which
// is what small, latency-bound transforms need.
// FZ: the launch carries element-wise companions (in_scale / out_scale / out_mul); plain transforms run the
// FZ = false instances, whose code does not know about them.
// Programmatic dependent launch (launches made with cudaLaunchAttributeProgrammaticStreamSerialization): the next
// kernel of the stream may become resident while this one drains its last wave, and does its index arithmetic
// there; it touches no data before grid_dependency_wait() returns, i.e. before this grid has completed and its
// stores are visible.  Both are no-ops for ordinary launches.
NTT_DEV void grid_dependency_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
NTT_DEV void grid_dependency_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <int GB, class F, bool FZ>
NTT_DEV void ntt_pass_body(const uint4* data_in, uint4* data_out, const typename F::Args& A, u32* smem) {
  constexpr int NE = 1 << GB;
  constexpr int NL = F::NL;
  constexpr int Q = NL / 4;  // uint4 per element
  const PassDesc& pd = A.pd;
  const typename F::Params& f = F::params(A);
  const int T = pd.T;
  const int d = pd.d;
  const u32 tid = threadIdx.x;
  const u32 tile_words = 1u << T;

  // ---- block-level index pieces (uniform)
  const uint2 bt = __ldg(A.blk + blockIdx.x);
  const unsigned long long base_t = bt.x;  // index inside one transform
  unsigned long long base = base_t;
  const u32 jlow_blk = bt.y;
  const u32 nthr = blockDim.x;
  base += (unsigned long long)blockIdx.y * A.batch_stride;
  const uint4* gbase_in = data_in + Q * base;
  uint4* gbase = data_out + Q * base;

  u32 x[NE][NL];
  const bool gen = (A.tw_gen_out != nullptr);
  const int lowp = T - d;  // lowest process bit

  grid_dependency_wait();  // the previous kernel of the stream is complete: data, tables and flags may be touched
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
      // acquire: only the observing threads fence (system scope); the CTA barrier extends the order to the rest
      // (fences are cumulative).  A fence by every thread of every CTA costs ~100 us per pass at 8 GPUs.
      __threadfence_system();
    }
    __syncthreads();
  }

  for (int r = gen ? pd.nrounds - 1 : 0; r < pd.nrounds; ++r) {
    const int sb = pd.round_sb[r];
    const int nb = pd.round_nb[r];
    const u32 lw = __ldg(A.thr + (3 + r) * nthr + tid);
    const u32 lam0 = lw & 0xffffu;
    const u32 w0 = lw >> 16;  // swz(lam0)
    // swz is linear over XOR: the swizzled offset of element u is w0 ^ (XOR of sx[i] over the set bits i of u)
    u32 sx[GB];
#pragma unroll
    for (int i = 0; i < GB; ++i) sx[i] = swz(1u << (sb + i));
    const bool first = (r == 0);
    const bool last = (r == pd.nrounds - 1);

    // ---- fetch the elements of this thread
    if (gen) {
#pragma unroll
      for (int u = 0; u < NE; ++u)
#pragma unroll
        for (int k = 0; k < NL; ++k) x[u][k] = A.one[k];
    } else if (first) {
      const u32 g0 = __ldg(A.thr + tid);
      u32 st[GB];
#pragma unroll
      for (int i = 0; i < GB; ++i) st[i] = 1u << pd.in_pos[sb + i];
#pragma unroll
      for (int u = 0; u < NE; ++u) {
        u32 off = g0;
#pragma unroll
        for (int i = 0; i < GB; ++i) off += ((u >> i) & 1) ? st[i] : 0u;
        F::ld(x[u], gbase_in + Q * (size_t)off);
        if (FZ && A.in_scale != nullptr) {
          u32 w[F::TW];
          F::ld_tw(w, A.in_scale, (u32)base_t + off);
          F::mul_tw(x[u], w, f);
        }
      }
    } else {
#pragma unroll
      for (int u = 0; u < NE; ++u) {
        u32 w = w0;
#pragma unroll
        for (int i = 0; i < GB; ++i) w ^= ((u >> i) & 1) ? sx[i] : 0u;
#pragma unroll
        for (int k = 0; k < NL; ++k) x[u][k] = smem[k * tile_words + w];
      }
    }

    // ---- butterflies on the top nb bits of the group (+ round twiddle unless last)
    if (!gen) {
      const bool scaled = (r == A.scaled_round);
      // every element with a processed bit set is multiplied after the round: by the round twiddle, or by the
      // pass twiddle / the 1/n of an inverse when this is the last round
      const bool tail_mul = !last || pd.has_tw || pd.scale_out;
      if (nb == 3) {
        if constexpr (GB >= 3) round_compute<GB, 3, F>(x, A, f, last, scaled, lam0, sb, lowp, tail_mul);
      } else if (nb == 2) {
        if constexpr (GB >= 2) round_compute<GB, 2, F>(x, A, f, last, scaled, lam0, sb, lowp, tail_mul);
      } else if (nb == 1) {
        round_compute<GB, 1, F>(x, A, f, last, scaled, lam0, sb, lowp, tail_mul);
      }
    }

    if (!last) {
      // ---- exchange
#pragma unroll
      for (int u = 0; u < NE; ++u) {
        u32 w = w0;
#pragma unroll
        for (int i = 0; i < GB; ++i) w ^= ((u >> i) & 1) ? sx[i] : 0u;
#pragma unroll
        for (int k = 0; k < NL; ++k) smem[k * tile_words + w] = x[u][k];
      }
      __syncthreads();
    } else {
      // ---- pass twiddle, scale, canonical form, store
      const u32 g0 = __ldg(A.thr + nthr + tid);
      u32 st[GB];
#pragma unroll
      for (int i = 0; i < GB; ++i) st[i] = 1u << pd.out_pos[sb + i];
      if (pd.has_tw && A.tw_stream != nullptr && !gen) {
#pragma unroll
        for (int u = 0; u < NE; ++u) {
          u32 off = (u32)base_t + g0;
#pragma unroll
          for (int i = 0; i < GB; ++i) off += ((u >> i) & 1) ? st[i] : 0u;
          u32 w[F::TW];
          F::ld_tw(w, A.tw_stream, off);
          F::mul_tw(x[u], w, f);
        }
      } else if (pd.has_tw) {
        const u32 jl0 = jlow_blk + __ldg(A.thr + 2 * nthr + tid);
        u32 gw[GB];
#pragma unroll
        for (int i = 0; i < GB; ++i) gw[i] = (sb + i < lowp) ? pd.wt_loc[sb + i] : 0u;
        const u32 nmask = (A.hi_bits + A.lo_bits >= 32) ? 0xffffffffu : ((1u << (A.hi_bits + A.lo_bits)) - 1u);
        const u32 lomask = (1u << A.lo_bits) - 1u;
        const bool two = pd.tw_lshift < A.lo_bits;
        // k_hi = the bit-reversed processed bits of the element's local index: the thread's part once, the
        // element's part as an OR of per-bit constants (disjoint bits)
        const u32 khi0 = d ? (__brev(lam0 >> lowp) >> (32 - d)) : 0u;
        u32 kb[GB];
#pragma unroll
        for (int i = 0; i < GB; ++i) kb[i] = d ? (__brev((1u << (sb + i)) >> lowp) >> (32 - d)) : 0u;
#pragma unroll
        for (int u = 0; u < NE; ++u) {
          u32 khi = khi0;
#pragma unroll
          for (int i = 0; i < GB; ++i) khi |= ((u >> i) & 1) ? kb[i] : 0u;
          u32 jl = jl0;
#pragma unroll
          for (int i = 0; i < GB; ++i) jl += ((u >> i) & 1) ? gw[i] : 0u;
          u32 e = (jl * khi) << pd.tw_lshift;
          if (A.inverse) e = (0u - e);
          e &= nmask;
          u32 w[F::TW];
          F::ld_tw(w, A.om_hi, e >> A.lo_bits);
          F::mul_tw(x[u], w, f);
          if (two) {
            F::ld_tw(w, A.om_lo, e & lomask);
            F::mul_tw(x[u], w, f);
          }
        }
      }
      if (pd.scale_out) {
#pragma unroll
        for (int u = 0; u < NE; ++u) F::mul_tw(x[u], A.scale, f);
      }
      if (FZ && (A.out_scale != nullptr || A.out_mul != nullptr)) {
#pragma unroll
        for (int u = 0; u < NE; ++u) {
          u32 off = (u32)base_t + g0;
#pragma unroll
          for (int i = 0; i < GB; ++i) off += ((u >> i) & 1) ? st[i] : 0u;
          if (A.out_scale != nullptr) {
            u32 w[F::TW];
            F::ld_tw(w, A.out_scale, off);
            F::mul_tw(x[u], w, f);
          }
          if (A.out_mul != nullptr) {
            // the vector is per call, not per transform: batches index it with the full offset
            F::mul_var(x[u], A.out_mul, (u32)(base - base_t) + off, f);
          }
        }
      }
      if (pd.is_final) {
#pragma unroll
        for (int u = 0; u < NE; ++u) F::canon(x[u], f);
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
          F::canon_exact(x[u], f);
          F::st_tw(A.tw_gen_out, off, x[u], A);
        }
      } else if (A.peer_bits == 0) {
#pragma unroll
        for (int u = 0; u < NE; ++u) {
          u32 off = g0;
#pragma unroll
          for (int i = 0; i < GB; ++i) off += ((u >> i) & 1) ? st[i] : 0u;
          F::st(gbase + Q * (size_t)off, x[u]);
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
          uint4* dst = reinterpret_cast<uint4*>(A.peer[off >> A.peer_shift]) + Q * (size_t)(mine | (off & lowmask));
          F::st(dst, x[u]);
        }
      }
    }
  }
  // Only now may the next kernel of the stream become resident: a trigger at the start would let the waiting CTAs
  // of several future launches pile up on the SMs of a small grid and squeeze out the running one (2^16: 35 -> 60 us).
  grid_dependency_launch_dependents();
  if (A.sync_mode & 2) {
    // this CTA's stores (peer memory included) are visible system-wide before it counts as arrived;
    // the last CTA of the grid tells every rank
    // release: the CTA barrier orders every thread's stores before thread 0, whose single system-scope fence then
    // orders them before the arrival count (cumulativity; the pattern of a cooperative grid barrier)
    __syncthreads();
    if (tid == 0) {
      __threadfence_system();
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

}  // namespace ntt256
