// Index helpers shared by the pass kernels of every limb count (ntt_kernels.cuh, ntt_kernels_n.cuh).
#pragma once
#include <cstdint>

#include "field.cuh"

namespace ntt256 {

// shared-memory tile: limb-major words, element index swizzled so that any 5 lane bits that
// are distinct mod 5 hit 32 distinct banks.
NTT_DEV u32 swz(u32 l) { return l ^ ((l >> 5) & 31u) ^ ((l >> 10) & 31u); }

// bit deposit: bit i of v goes to position pos[i]
NTT_DEV u32 deposit32(u32 v, const int8_t* pos, int n) {
  u32 r = 0;
  for (int i = 0; i < n; ++i) r |= ((v >> i) & 1u) << pos[i];
  return r;
}
NTT_DEV unsigned long long deposit64(unsigned long long v, const int8_t* pos, int n) {
  unsigned long long r = 0;
  for (int i = 0; i < n; ++i) r |= ((v >> i) & 1ull) << pos[i];
  return r;
}

}  // namespace ntt256
