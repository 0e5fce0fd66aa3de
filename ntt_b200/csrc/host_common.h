// Host-side pieces shared by the C-ABI layers of every container width (capi.cu, capi384.cu, capi64.cu): the
// optional parts of a plan launch sequence, the error macro, and the rank-synchronisation setter.
#pragma once
#include <cuda_runtime.h>

#include <string>

#include "../../include/ntt256.h"

namespace ntt256 {

// the last pass of a four-step step 1 stores into the ranks' receive buffers (peer-mapped pointers)
struct PeerSpec {
  const void* const* ptrs = nullptr;
  int bits = 0, shift = 0, rank = 0;
};
// optional out-of-place ends of a plan (device-visible pointers): the first pass reads src_first, the last writes dst_last
struct IoSpec {
  const void* src_first = nullptr;
  void* dst_last = nullptr;
};
// element-wise companions fused into the first / last pass (PassArgsT::in_scale ...)
struct FuseSpec {
  const uint4* in_scale = nullptr;
  const uint4* out_scale = nullptr;
  const uint4* out_mul = nullptr;
};

// Arms (flag_bufs: every rank's flag buffer as mapped in this process) or disarms (null) the rank synchronisation
// folded into the four-step pass kernels of a context with the fields sync_flags[8], sync_nranks, sync_rank,
// sync_epoch, sync_timeout_ns, sync_counter, device, last_err.
template <class Ctx>
int set_peer_sync_impl(Ctx* ctx, void* const* flag_bufs, int nranks, int rank, double timeout_s) {
  if (!ctx) return NTT256_ERR_ARG;
  if (cudaSetDevice(ctx->device) != cudaSuccess) {
    ctx->last_err = "cudaSetDevice";
    return NTT256_ERR_CUDA;
  }
  if (!flag_bufs || nranks <= 1) {  // disarm
    ctx->sync_nranks = 0;
    ctx->sync_epoch = 0;
    return NTT256_OK;
  }
  if (nranks > 8 || (nranks & (nranks - 1)) || rank < 0 || rank >= nranks) return NTT256_ERR_ARG;
  for (int r = 0; r < nranks; ++r) {
    if (!flag_bufs[r]) return NTT256_ERR_ARG;
    ctx->sync_flags[r] = (unsigned long long)flag_bufs[r];
  }
  if (!ctx->sync_counter) {
    if (cudaMalloc((void**)&ctx->sync_counter, sizeof(unsigned int)) != cudaSuccess ||
        cudaMemset(ctx->sync_counter, 0, sizeof(unsigned int)) != cudaSuccess) {
      ctx->last_err = "cudaMalloc / cudaMemset (rank synchronisation counter)";
      return NTT256_ERR_CUDA;
    }
  }
  ctx->sync_nranks = nranks;
  ctx->sync_rank = rank;
  ctx->sync_epoch = 0;
  if (timeout_s > 0) ctx->sync_timeout_ns = (unsigned long long)(timeout_s * 1e9);
  return NTT256_OK;
}

// Host batches through a three-slot staging ring on three streams, so that the PCIe uplink, the kernels and the
// downlink work on different chunks at the same time (pinned host memory makes the copies asynchronous; pageable
// memory still gives the right answer).  `one` = bytes of one transform, per_chunk transforms per slot;
// run_chunk(device pointer, transforms in the chunk, stream) launches the transforms of one chunk.  The context
// supplies s_h2d / s_d2h / s_comp, ev_in / ev_done / ev_out[3], d_stage (at least 3 * one * per_chunk bytes).
template <class Ctx, class RunChunk>
int host_ring(Ctx* ctx, const void* h_in, void* h_out, size_t one, int batch, int per_chunk, RunChunk run_chunk) {
#define HR_CUDA(call)                                                       \
  do {                                                                      \
    cudaError_t e_ = (call);                                                \
    if (e_ != cudaSuccess) {                                                \
      ctx->last_err = std::string(#call) + ": " + cudaGetErrorString(e_);   \
      return NTT256_ERR_CUDA;                                               \
    }                                                                       \
  } while (0)
  if (!ctx->s_h2d) {
    HR_CUDA(cudaStreamCreateWithFlags(&ctx->s_h2d, cudaStreamNonBlocking));
    HR_CUDA(cudaStreamCreateWithFlags(&ctx->s_d2h, cudaStreamNonBlocking));
    HR_CUDA(cudaStreamCreateWithFlags(&ctx->s_comp, cudaStreamNonBlocking));
    for (int i = 0; i < 3; ++i) {
      HR_CUDA(cudaEventCreateWithFlags(&ctx->ev_in[i], cudaEventDisableTiming));
      HR_CUDA(cudaEventCreateWithFlags(&ctx->ev_done[i], cudaEventDisableTiming));
      HR_CUDA(cudaEventCreateWithFlags(&ctx->ev_out[i], cudaEventDisableTiming));
    }
  }
  const size_t slot_bytes = one * per_chunk;
  const int nchunks = (batch + per_chunk - 1) / per_chunk;
  for (int c = 0; c < nchunks; ++c) {
    const int slot = c % 3;
    const int b0 = c * per_chunk, nb = batch - b0 < per_chunk ? batch - b0 : per_chunk;
    char* d = (char*)ctx->d_stage + slot * slot_bytes;
    if (c >= 3) HR_CUDA(cudaStreamWaitEvent(ctx->s_h2d, ctx->ev_out[slot], 0));  // slot drained
    HR_CUDA(cudaMemcpyAsync(d, (const char*)h_in + (size_t)b0 * one, (size_t)nb * one, cudaMemcpyHostToDevice, ctx->s_h2d));
    HR_CUDA(cudaEventRecord(ctx->ev_in[slot], ctx->s_h2d));
    HR_CUDA(cudaStreamWaitEvent(ctx->s_comp, ctx->ev_in[slot], 0));
    const int rc = run_chunk((void*)d, nb, ctx->s_comp);
    if (rc) return rc;
    HR_CUDA(cudaEventRecord(ctx->ev_done[slot], ctx->s_comp));
    HR_CUDA(cudaStreamWaitEvent(ctx->s_d2h, ctx->ev_done[slot], 0));
    HR_CUDA(cudaMemcpyAsync((char*)h_out + (size_t)b0 * one, d, (size_t)nb * one, cudaMemcpyDeviceToHost, ctx->s_d2h));
    HR_CUDA(cudaEventRecord(ctx->ev_out[slot], ctx->s_d2h));
  }
  HR_CUDA(cudaStreamSynchronize(ctx->s_d2h));
  HR_CUDA(cudaStreamSynchronize(ctx->s_comp));
#undef HR_CUDA
  return NTT256_OK;
}

}  // namespace ntt256
