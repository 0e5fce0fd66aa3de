// The BN254 Fr instance of the throughput kernel (field constants as immediates): policy, kernel and launcher in their
// own translation unit, so that ptxas works on it in parallel with the generic instances.
#include "ntt_kernels.cuh"
#include "pass_launch.h"

namespace ntt256 {

// The headline field with its constants known at compile time: the modulus, 2p, 3p and their complements become
// immediates of the add / multiply instructions instead of constant-bank loads and register copies (-8 instructions
// per butterfly of ~250; every instruction is an issue slot here).  Same arithmetic as F256<true>; the host launches
// this instance when the context's modulus is BN254 Fr (capi.cu) and checks the constants against the context's.
__device__ constexpr FieldParams kBn254FrParams = {
    {0xf0000001u, 0x43e1f593u, 0x79b97091u, 0x2833e848u, 0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u},
    {0xe0000002u, 0x87c3eb27u, 0xf372e122u, 0x5067d090u, 0x0302b0bau, 0x70a08b6du, 0xc2634053u, 0x60c89ce5u},
    {0x0fffffffu, 0xbc1e0a6cu, 0x86468f6eu, 0xd7cc17b7u, 0x7e7ea7a2u, 0x47afba49u, 0x1ece5fd6u, 0xcf9bb18du},
    {0x1ffffffeu, 0x783c14d8u, 0x0c8d1eddu, 0xaf982f6fu, 0xfcfd4f45u, 0x8f5f7492u, 0x3d9cbfacu, 0x9f37631au},
    {0xd0000003u, 0xcba5e0bbu, 0x6d2c51b3u, 0x789bb8d9u, 0x84840917u, 0x28f0d123u, 0xa394e07du, 0x912ceb58u},
    0xefffffffu,
    {0u, 0u, 0u}};
struct F256Bn254 : F256<true> {
  static NTT_DEV const FieldParams& params(const Args&) { return kBn254FrParams; }
  static NTT_DEV void st_tw(uint4* tab, u32 idx, const u32* w, const Args& A) {
    st_shoup_entry(tab + 4 * (size_t)idx, w, A.r2, A.pinv, kBn254FrParams);
  }
  static NTT_DEV void mul_var(u32* x, const uint4* vec, u32 idx, const FieldParams& f) { F256<true>::mul_var(x, vec, idx, f); }
};
__global__ void __launch_bounds__(512, 1)
ntt_pass_bn254_kernel(const uint4* data_in, uint4* data_out, const __grid_constant__ PassArgs A) {
  extern __shared__ u32 smem[];
  ntt_pass_body<3, F256Bn254, false>(data_in, data_out, A, smem);
}

void launch_pass_bn254(dim3 grid, unsigned threads, size_t smem, cudaStream_t stream, const uint4* in, uint4* out, const PassArgs& A,
                       bool pdl) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(threads, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  cudaLaunchKernelEx(&cfg, ntt_pass_bn254_kernel, in, out, A);
}
bool pass_bn254_set_smem(int bytes) {
  return cudaFuncSetAttribute(ntt_pass_bn254_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes) == cudaSuccess;
}
bool pass_bn254_params_match(const FieldParams& f) {
  const FieldParams k = kBn254FrParams;
  for (int i = 0; i < 8; ++i)
    if (k.p[i] != f.p[i] || k.p2[i] != f.p2[i] || k.pbar[i] != f.pbar[i] || k.p2bar[i] != f.p2bar[i] || k.kp[i] != f.kp[i]) return false;
  return k.np0 == f.np0;
}
}  // namespace ntt256
