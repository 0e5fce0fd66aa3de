// Launchers of the pass kernels.  The twelve instances (Montgomery / Shoup twiddles x plain / fused companions x
// 8, 4 or 2 elements per thread) are compiled in four translation units (pass_inst_*.cu) so that ptxas works on
// them in parallel; capi.cu only sees these declarations.
#pragma once
#include <cuda_runtime.h>

namespace ntt256 {
struct FieldParams;
template <class PARAMS, int NL, int TWS>
struct PassArgsT;
using PassArgs = PassArgsT<FieldParams, 8, 16>;
// pdl: launch with cudaLaunchAttributeProgrammaticStreamSerialization (see grid_dependency_wait in ntt_pass_generic.cuh)
template <bool SH, bool FZ>
void launch_pass_inst(int group_bits, bool dense, dim3 grid, unsigned threads, size_t smem, cudaStream_t stream, const uint4* in,
                      uint4* out, const PassArgs& A, bool pdl);
// The eight-elements-per-thread Shoup kernel specialised on BN254 Fr (constants as immediates; plain launches only);
// pass_bn254_params_match: the context's constants are exactly the compiled-in ones.
void launch_pass_bn254(dim3 grid, unsigned threads, size_t smem, cudaStream_t stream, const uint4* in, uint4* out, const PassArgs& A,
                       bool pdl);
bool pass_bn254_set_smem(int bytes);
bool pass_bn254_params_match(const FieldParams& f);
// dense: four elements per thread on the three-CTAs-per-SM instance (ntt_pass4h_kernel) when the CTA fits it
// raises the dynamic shared memory limit of the three kernels of one instance; false on a CUDA error
template <bool SH, bool FZ>
bool pass_inst_set_smem(int bytes_pass8, int bytes_small);
}  // namespace ntt256
