// ntt256_compat.hpp -- the reference's C++ driver signatures on top of the C ABI (ntt256.h).
//
// A program written against /root/reference/src/big-num.cu can keep calling
//     NTT_GPU_Naive<TPI, BITS>(data_d, length, reverse2_d, reverse_num, prime, omega);   (big-num.cu:439)
//     NTT_GZKP<TPI, BITS>(data_d, length, reverse2_d, reverse_num, prime, omega, 5, 8);  (big-num.cu:458)
// with the same argument meaning: `data` is a DEVICE array of cgbn_mem_t<256> (8 LE u32 limbs,
// canonical residues) transformed in place, natural order in and out; `prime`, `omega` are host
// values and `omega` is the GENERATOR (w_n = omega^((p-1)/len) is derived inside).  The
// bit-reversal pair list (`reverse`, `reverse_len`) and the tile knobs (B, G) are accepted and
// ignored: this implementation has no permutation pass and plans its own tiling.
// Like the reference drivers the calls are synchronous, print the kernel time
// ("naive: %fms" / "GZKP: %fms", big-num.cu:166, 349; define NTT256_COMPAT_QUIET to silence) and
// treat errors as fatal (the reference asserts / ignores CUDA status).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <utility>
#include <vector>

#include "ntt256.h"

#ifndef CGBN_H  // the reference's own type when CGBN is not on the include path
template <uint32_t bits>
struct cgbn_mem_t {
  uint32_t _limbs[(bits + 31) / 32];  // include/cgbn/cgbn_cuda.h:51-55
};
#endif

namespace ntt256_compat {

inline ntt256_ctx* context_for(const uint32_t p[8], const uint32_t g[8], int log_n) {
  struct Key {
    std::vector<uint32_t> v;
    bool operator<(const Key& o) const { return v < o.v; }
  };
  static std::map<Key, std::pair<ntt256_ctx*, int>> cache;
  Key k;
  k.v.assign(p, p + 8);
  k.v.insert(k.v.end(), g, g + 8);
  auto it = cache.find(k);
  if (it != cache.end() && it->second.second >= log_n) return it->second.first;
  if (it != cache.end()) ntt256_ctx_destroy(it->second.first);
  // largest size this field supports, capped at the reference's own limit (assert(logl <= 26), big-num.cu:386)
  int two_adicity = 0;
  for (int w = 0; w < 8 && two_adicity == 32 * w; ++w) {
    uint32_t x = (w == 0) ? p[0] - 1 : p[w];
    if (x == 0) { two_adicity += 32; continue; }
    while (!(x & 1)) { x >>= 1; ++two_adicity; }
  }
  int max_log_n = two_adicity < 28 ? two_adicity : 28;
  if (max_log_n < log_n) max_log_n = log_n;
  ntt256_ctx* ctx = nullptr;
  int dev = 0;
  cudaGetDevice(&dev);
  int rc = ntt256_ctx_create(&ctx, dev, p, g, max_log_n);
  if (rc) {
    std::fprintf(stderr, "ntt256: %s\n", ntt256_strerror(rc));
    std::abort();
  }
  cache[k] = {ctx, max_log_n};
  return ctx;
}

template <uint32_t bits>
inline void run(const char* label, cgbn_mem_t<bits>* data, unsigned len, cgbn_mem_t<bits> prime, cgbn_mem_t<bits> omega) {
  static_assert(bits == 256, "libntt256 transforms 256-bit containers");
  int log_n = 0;
  while ((1u << log_n) < len) ++log_n;
  if ((1u << log_n) != len) {
    std::fprintf(stderr, "ntt256: length must be a power of two\n");
    std::abort();
  }
  ntt256_ctx* ctx = context_for(prime._limbs, omega._limbs, log_n);
  cudaEvent_t start, end;
  cudaEventCreate(&start);
  cudaEventCreate(&end);
  cudaEventRecord(start);
  int rc = ntt256_forward(ctx, data, log_n, nullptr);
  cudaEventRecord(end);
  cudaEventSynchronize(end);
  if (rc) {
    std::fprintf(stderr, "ntt256: %s (%s)\n", ntt256_strerror(rc), ntt256_last_cuda_error(ctx));
    std::abort();
  }
  float t = 0;
  cudaEventElapsedTime(&t, start, end);
#ifndef NTT256_COMPAT_QUIET
  std::printf("%s: %fms\n", label, t);
#endif
  cudaEventDestroy(start);
  cudaEventDestroy(end);
}

}  // namespace ntt256_compat

// == NTT_GPU_Naive<tpi,bits>, big-num.cu:106-107
template <const unsigned tpi, const unsigned bits>
void NTT_GPU_Naive(cgbn_mem_t<bits> data[], unsigned len, uint2* /*reverse*/, unsigned /*reverse_len*/,
                   cgbn_mem_t<bits> prime, cgbn_mem_t<bits> omega) {
  ntt256_compat::run<bits>("naive", data, len, prime, omega);
}

// == NTT_GZKP<tpi,bits>, big-num.cu:260-261
template <const unsigned tpi, const unsigned bits>
void NTT_GZKP(cgbn_mem_t<bits> data[], unsigned len, uint2* /*reverse*/, unsigned /*reverse_len*/,
              cgbn_mem_t<bits> prime, cgbn_mem_t<bits> omega, unsigned /*B*/, unsigned /*G*/) {
  ntt256_compat::run<bits>("GZKP", data, len, prime, omega);
}
