// Host side of the index tables the pass kernels look up instead of depositing bits (every container width).
//
// A pass needs, per thread and round, the local element index its thread id stands for (the bits of tid deposited
// at the positions PassDesc::tmap[r] lists), its swizzled shared-memory offset, the global offset of its first
// element when loading / storing (a second deposit through in_pos / out_pos) and its share of the twiddle index
// j_low; per CTA the tile's base index and its share of j_low.  They depend on the plan only, so they are built
// once per plan on the host and read with one LDG each -- a loop over T index bits per thread and round is 5 % of
// the instructions of the 256-bit kernels and a quarter of those of the 64-bit ones.
//   thr[0 .. t)          element offset of the thread's first element when loading  (round 0's lam0 through in_pos)
//   thr[t .. 2t)         ... when storing (last round's lam0 through out_pos)
//   thr[2t .. 3t)        the thread's share of j_low (passive local bits below the processed ones, last round)
//   thr[3t + r t ...)    round r: lam0 | swz(lam0) << 16
//   blk[b]               .x = index of the tile's first element inside one transform, .y = the block's share of j_low
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <vector>

#include "plan.h"

namespace ntt256 {

struct PassTables {  // per pass of one plan: device pointers (one allocation each)
  std::vector<uint32_t*> thr;
  std::vector<uint2*> blk;
};

inline void free_pass_tables(PassTables* t) {
  for (uint32_t* p : t->thr) cudaFree(p);
  for (uint2* p : t->blk) cudaFree(p);
  t->thr.clear();
  t->blk.clear();
}

inline uint32_t pass_tables_deposit(unsigned long long v, const int8_t* pos, int n) {
  unsigned long long r = 0;
  for (int i = 0; i < n; ++i) r |= ((v >> i) & 1ull) << pos[i];
  return (uint32_t)r;
}
inline uint32_t pass_tables_swz(uint32_t l) { return l ^ ((l >> 5) & 31u) ^ ((l >> 10) & 31u); }  // kernel_common.cuh swz

// The tables of one pass, on the host (pure arithmetic: tests/cpp/pass_tables_test.cpp checks their invariants without a GPU).
inline void compute_pass_tables(const PassDesc& pd, int GB, std::vector<uint32_t>* thr_out, std::vector<uint2>* blk_out) {
  const int T = pd.T, lowp = pd.T - pd.d;
  const uint32_t t = 1u << (T - GB);
  std::vector<uint32_t>& thr = *thr_out;
  thr.assign((size_t)(3 + pd.nrounds) * t, 0u);
  for (uint32_t tid = 0; tid < t; ++tid) {
    const uint32_t lam_first = pass_tables_deposit(tid, pd.tmap[0], T - GB);
    const uint32_t lam_last = pass_tables_deposit(tid, pd.tmap[pd.nrounds - 1], T - GB);
    thr[tid] = pass_tables_deposit(lam_first, pd.in_pos, T);
    thr[t + tid] = pass_tables_deposit(lam_last, pd.out_pos, T);
    uint32_t jl = 0;
    for (int i = 0; i < lowp; ++i) jl += ((lam_last >> i) & 1u) * pd.wt_loc[i];
    thr[2 * t + tid] = jl;
    for (int r = 0; r < pd.nrounds; ++r) {
      const uint32_t lam = pass_tables_deposit(tid, pd.tmap[r], T - GB);
      thr[(size_t)(3 + r) * t + tid] = lam | (pass_tables_swz(lam) << 16);
    }
  }
  const size_t nblk = (size_t)1 << pd.nblk_bits;
  std::vector<uint2>& blk = *blk_out;
  blk.resize(nblk);
  for (size_t b = 0; b < nblk; ++b) {
    uint32_t jl = pd.jlow_const;
    for (int i = 0; i < pd.nblk_bits; ++i) jl += (uint32_t)((b >> i) & 1) * pd.wt_blk[i];
    uint2 e;
    e.x = pass_tables_deposit(b, pd.blk_pos, pd.nblk_bits);
    e.y = jl;
    blk[b] = e;
  }
}

// cudaSuccess, or the first CUDA error (tables built so far are freed).
inline cudaError_t build_pass_tables(const Plan& plan, PassTables* out) {
  const int GB = plan.group_bits;
  cudaError_t e = cudaSuccess;
  for (const PassDesc& pd : plan.passes) {
    std::vector<uint32_t> thr;
    std::vector<uint2> blk;
    compute_pass_tables(pd, GB, &thr, &blk);
    uint32_t* d_thr = nullptr;
    uint2* d_blk = nullptr;
    if ((e = cudaMalloc((void**)&d_thr, thr.size() * 4)) != cudaSuccess) break;
    out->thr.push_back(d_thr);
    if ((e = cudaMalloc((void**)&d_blk, blk.size() * 8)) != cudaSuccess) break;
    out->blk.push_back(d_blk);
    if ((e = cudaMemcpy(d_thr, thr.data(), thr.size() * 4, cudaMemcpyHostToDevice)) != cudaSuccess) break;
    if ((e = cudaMemcpy(d_blk, blk.data(), blk.size() * 8, cudaMemcpyHostToDevice)) != cudaSuccess) break;
  }
  // the DMA of a pageable copy may trail its return; callers' streams need not be blocking
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) free_pass_tables(out);
  return e;
}

}  // namespace ntt256
