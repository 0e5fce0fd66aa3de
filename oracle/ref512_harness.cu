// TEST INFRASTRUCTURE ONLY.  The reference's NTT_GZKP driver (/root/reference/src/big-num.cu:260-353)
// instantiated at 512-bit containers (tpi = 16, bits = 512) -- a separate translation unit from
// ref256_harness.cu because the kernel's `extern __shared__ cgbn_mem_t<bits> s[]` (big-num.cu:177)
// cannot be instantiated for two widths in one.  Compiled from where the reference lies, nothing copied.
// Output: oracle/_ref/libref512.so (git-ignored, travels with gpurun); needs a GPU to run.
#include <cstdint>
#include <cstring>
#include <vector>
// `int main() {` becomes a declaration followed by an (never instantiated) function template, so the
// program's own NTT_GZKP<8,256> call is not instantiated next to the <16,512> one below.
#define main ref_main_decl(int = 0); template <int Unused_> int ref_bignum_main_unused
#include "/root/reference/src/big-num.cu"
#undef main
#undef P
#undef root

extern "C" {

// The same NTT_GZKP driver instantiated at 512-bit containers (tpi = 16): the only wider shape the
// reference's kernels accept (they keep one limb per thread, big-num.cu:108, 268).  Used to pin this
// repo's 384-bit path: a 377-bit field element is the same integer in a 384- or a 512-bit container.
// B = 5, G = 4 keeps block.x * tpi at the 1024-thread limit (big-num.cu:266-267).
int ref512_ntt_gzkp(uint32_t* data /* n x 16 limbs */, unsigned log_n, const uint32_t p[16], const uint32_t g[16]) {
  const uint len = 1u << log_n;
  std::vector<uint> reverse(len, 0);
  for (uint i = 0; i < len; i++) reverse[i] = (reverse[i >> 1] >> 1) | ((i & 1u) << (log_n - 1));
  std::vector<uint2> pairs;
  for (uint i = 0; i < len; i++)
    if (reverse[i] < i) pairs.push_back(make_uint2(reverse[i], i));
  uint2* pairs_d = nullptr;
  cgbn_mem_t<512>* data_d = nullptr;
  if (cudaMalloc(&pairs_d, (pairs.size() + 1) * sizeof(uint2)) != cudaSuccess) return (int)cudaGetLastError();
  if (cudaMalloc(&data_d, (size_t)len * sizeof(*data_d)) != cudaSuccess) return (int)cudaGetLastError();
  cudaMemcpy(pairs_d, pairs.data(), pairs.size() * sizeof(uint2), cudaMemcpyHostToDevice);
  cudaMemcpy(data_d, data, (size_t)len * sizeof(*data_d), cudaMemcpyHostToDevice);
  cgbn_mem_t<512> prime, omega;
  std::memcpy(prime._limbs, p, 64);
  std::memcpy(omega._limbs, g, 64);
  NTT_GZKP<16, 512>(data_d, len, pairs_d, (uint)pairs.size(), prime, omega, 5, 4);
  cudaError_t e = cudaDeviceSynchronize();
  cudaMemcpy(data, data_d, (size_t)len * sizeof(*data_d), cudaMemcpyDeviceToHost);
  cudaFree(data_d);
  cudaFree(pairs_d);
  if (e == cudaSuccess) e = cudaGetLastError();
  return (int)e;
}
}
