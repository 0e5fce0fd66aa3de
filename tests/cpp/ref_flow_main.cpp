// TEST PROGRAM (tests/ only: it links the CPU oracle).  Re-creates the flow of the reference's
// own self-checking main() (/root/reference/src/big-num.cu:369-488) against libntt256.so through
// include/ntt256_compat.hpp:  for kk = 5..24: x[i] = i as cgbn_mem_t<256>, CPU NTT, H2D,
// NTT_GPU_Naive, D2H, compare, H2D, NTT_GZKP(..., 5, 8), D2H, compare -- but over the FULL 256 bits
// (the reference compares the low 64 bits only, :446-451) and for three fields (the reference's
// p = 469762049 / g = 3, BN254 Fr / 5, BLS12-381 Fr / 7).  Mismatches are printed as
// "expected got index" like the reference does, and -- unlike the reference -- the exit code is 1.
#include <cstring>
#include <vector>

#include "ntt256_compat.hpp"

#define TPI 8
#define BITS 256

extern "C" int oracle_ntt256(uint32_t* data, int log_n, const uint32_t p[8], const uint32_t g[8], int inverse, int threads);

template <uint32_t bits>
static cgbn_mem_t<bits> int64_to_cgbn(long long x) {  // big-num.cu:355-362
  cgbn_mem_t<bits> r;
  r._limbs[0] = (uint32_t)(x & 0xffffffff);
  r._limbs[1] = (uint32_t)(x >> 32);
  for (unsigned i = 2; i < bits / 32; i++) r._limbs[i] = 0;
  return r;
}

int main(int argc, char** argv) {
  const int kk_max = argc > 1 ? std::atoi(argv[1]) : 24;
  struct FieldCase {
    const char* name;
    uint32_t p[8];
    long long root;
    int max_kk;
  } fields[] = {
      {"ref29", {469762049u, 0, 0, 0, 0, 0, 0, 0}, 3, 24},
      {"bn254_fr", {0xf0000001u, 0x43e1f593u, 0x79b97091u, 0x2833e848u, 0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u}, 5, 24},
      {"bls12_381_fr", {0x00000001u, 0xffffffffu, 0xfffe5bfeu, 0x53bda402u, 0x09a1d805u, 0x3339d808u, 0x299d7d48u, 0x73eda753u}, 7, 24},
  };
  long long mismatches = 0;
  for (auto& fc : fields) {
    std::printf("== field %s\n", fc.name);
    for (int kk = 5; kk <= kk_max && kk <= fc.max_kk; kk++) {
      const unsigned length = 1u << kk;
      std::vector<cgbn_mem_t<BITS>> data(length), data_copy(length), tmp(length);
      for (unsigned i = 0; i < length; i++) data[i] = data_copy[i] = int64_to_cgbn<BITS>(i);  // data[i] = i, :399-402
      cgbn_mem_t<BITS> prime, omega = int64_to_cgbn<BITS>(fc.root);
      std::memcpy(prime._limbs, fc.p, 32);
      // cpu implementation (:405-412): the oracle restates NTT() over the 256-bit field
      if (oracle_ntt256(data[0]._limbs, kk, prime._limbs, omega._limbs, 0, 8)) {
        std::printf("oracle failed\n");
        return 2;
      }
      cgbn_mem_t<BITS>* data_d;
      cudaMalloc(&data_d, length * sizeof(*data_d));
      for (int variant = 0; variant < 2; ++variant) {
        cudaMemcpy(data_d, data_copy.data(), length * sizeof(*data_d), cudaMemcpyHostToDevice);
        if (variant == 0)
          NTT_GPU_Naive<TPI, BITS>(data_d, length, nullptr, 0, prime, omega);  // :439
        else
          NTT_GZKP<TPI, BITS>(data_d, length, nullptr, 0, prime, omega, 5, 8);  // :458
        cudaMemcpy(tmp.data(), data_d, sizeof(*data_d) * length, cudaMemcpyDeviceToHost);
        for (unsigned i = 0; i < length; i++)
          if (std::memcmp(&data[i], &tmp[i], 32)) {
            if (mismatches < 10)
              std::printf("%08x.. %08x.. %u\n", data[i]._limbs[0], tmp[i]._limbs[0], i);  // expected got index
            ++mismatches;
          }
      }
      cudaFree(data_d);
    }
  }
  std::printf(mismatches ? "error: %lld mismatches\n" : "correct (%lld mismatches)\n", mismatches);
  return mismatches ? 1 : 0;
}
