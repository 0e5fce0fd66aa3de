// TEST INFRASTRUCTURE ONLY.  The reference's own 256-bit GPU drivers behind a C ABI, so that the
// CUDA path of this repo can be pinned against them on ANY 256-bit field (the reference's main()
// only ever runs p = 469762049, but its drivers and kernels take the modulus and the generator as
// 256-bit arguments):
//   NTT_GPU_Naive<8,256>   /root/reference/src/big-num.cu:106-170  (rearrange + one launch per stage)
//   NTT_GZKP<8,256>        /root/reference/src/big-num.cu:260-353  (B = 5, G = 8 as in its main(), :458)
// The reference source is compiled from where it lies (nothing is copied); <gmp.h> is the
// declaration shim in oracle/gmp_shim (the image has libgmp.so.10 but not the header).
// Output: oracle/_ref/libref256.so (git-ignored, travels with gpurun).  Needs a GPU to run:
// scripts/gen_golden_ref256.py turns its outputs into tests/golden/ref256_gpu_kat.json on the box.
#include <cstdint>
#include <cstring>
#include <vector>
#define main ref_bignum_main_unused
#include "/root/reference/src/big-num.cu"
#undef main
#undef P
#undef root

extern "C" {

// data: n = 2^log_n elements of 8 little-endian u32 limbs (host memory), transformed in place by the
// reference driver `which` (0 = NTT_GPU_Naive, 1 = NTT_GZKP(...,5,8)); p, g: modulus and generator.
// Returns 0, or a CUDA error code.
int ref256_ntt(uint32_t* data, unsigned log_n, const uint32_t p[8], const uint32_t g[8], int which) {
  const uint len = 1u << log_n;
  // the bit-reversal pair list exactly as the reference main() builds it (big-num.cu:393-395, 414-423)
  std::vector<uint> reverse(len, 0);
  for (uint i = 0; i < len; i++) reverse[i] = (reverse[i >> 1] >> 1) | ((i & 1u) << (log_n - 1));
  std::vector<uint2> pairs;
  pairs.reserve(len / 2 + 1);
  for (uint i = 0; i < len; i++)
    if (reverse[i] < i) pairs.push_back(make_uint2(reverse[i], i));
  uint2* pairs_d = nullptr;
  cgbn_mem_t<256>* data_d = nullptr;
  if (cudaMalloc(&pairs_d, (pairs.size() + 1) * sizeof(uint2)) != cudaSuccess) return (int)cudaGetLastError();
  if (cudaMalloc(&data_d, (size_t)len * sizeof(*data_d)) != cudaSuccess) return (int)cudaGetLastError();
  cudaMemcpy(pairs_d, pairs.data(), pairs.size() * sizeof(uint2), cudaMemcpyHostToDevice);
  cudaMemcpy(data_d, data, (size_t)len * sizeof(*data_d), cudaMemcpyHostToDevice);
  cgbn_mem_t<256> prime, omega;
  std::memcpy(prime._limbs, p, 32);
  std::memcpy(omega._limbs, g, 32);
  if (which == 0)
    NTT_GPU_Naive<8, 256>(data_d, len, pairs_d, (uint)pairs.size(), prime, omega);
  else
    NTT_GZKP<8, 256>(data_d, len, pairs_d, (uint)pairs.size(), prime, omega, 5, 8);
  cudaError_t e = cudaDeviceSynchronize();
  cudaMemcpy(data, data_d, (size_t)len * sizeof(*data_d), cudaMemcpyDeviceToHost);
  cudaFree(data_d);
  cudaFree(pairs_d);
  if (e == cudaSuccess) e = cudaGetLastError();
  return (int)e;
}
}
