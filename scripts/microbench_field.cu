// Field-layer microbenchmark (round 2): the Montgomery product of field.cuh against the Shoup/Barrett
// product of field_shoup.cuh, alone and inside a DIF butterfly, plus a bit-exactness check of both
// against known answers (scripts/shoup_testvec.h, generated with Python integers).
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o microbench_field microbench_field.cu
// Run:   ./microbench_field [iters [reps]]
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "../ntt_b200/csrc/field.cuh"
#include "../ntt_b200/csrc/field_shoup.cuh"
#include "../ntt_b200/csrc/field_wl.cuh"
#include "shoup_testvec.h"

using namespace ntt256;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

// vec[i] = {x, w, wq, wR, x*w mod p}
__global__ void k_check(const u32* __restrict__ vec, int n, const __grid_constant__ FieldParams f, int* bad) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  u32 x[8], w[8], wq[8], wm[8], want[8], r1[8], r2[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    x[k] = vec[(i * 5 + 0) * 8 + k];
    w[k] = vec[(i * 5 + 1) * 8 + k];
    wq[k] = vec[(i * 5 + 2) * 8 + k];
    wm[k] = vec[(i * 5 + 3) * 8 + k];
    want[k] = vec[(i * 5 + 4) * 8 + k];
  }
  shoup_mul(r1, x, w, wq, f);
  canonicalize_wl(r1, f);
  mont_mul(r2, wm, x, f);
  canonicalize(r2, f);
  int b1 = 0, b2 = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    b1 |= (r1[k] != want[k]);
    b2 |= (r2[k] != want[k]);
  }
  if (b1) atomicAdd(&bad[0], 1);
  if (b2) atomicAdd(&bad[1], 1);
}

template <int N, int MODE>  // MODE 0: mont_mul, 1: shoup_mul
__global__ void __launch_bounds__(256) k_mul(u32* out, const u32* in, const __grid_constant__ FieldParams f, int iters) {
  u32 a[8], aq[8], b[N][8];
  int t = blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    a[k] = in[((t & 255) * 5 + (MODE ? 1 : 3)) * 8 + k];
    aq[k] = in[((t & 255) * 5 + 2) * 8 + k];
  }
#pragma unroll
  for (int n = 0; n < N; ++n)
#pragma unroll
    for (int k = 0; k < 8; ++k) b[n][k] = in[(((t + n * 7) & 255) * 5 + 0) * 8 + k];
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int n = 0; n < N; ++n) {
      u32 r[8];
      if (MODE) shoup_mul(r, b[n], a, aq, f); else mont_mul(r, a, b[n], f);
#pragma unroll
      for (int k = 0; k < 8; ++k) b[n][k] = r[k];
    }
  }
  u32 s = 0;
#pragma unroll
  for (int n = 0; n < N; ++n)
#pragma unroll
    for (int k = 0; k < 8; ++k) s ^= b[n][k];
  if (s == 0x12345678) out[0] = s;
}

template <int N, int MODE>
__global__ void __launch_bounds__(256) k_bfly(u32* out, const u32* in, const __grid_constant__ FieldParams f, int iters) {
  u32 w[8], wq[8], x[N][8], y[N][8];
  int t = blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    w[k] = in[((t & 255) * 5 + (MODE ? 1 : 3)) * 8 + k];
    wq[k] = in[((t & 255) * 5 + 2) * 8 + k];
  }
#pragma unroll
  for (int n = 0; n < N; ++n)
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      x[n][k] = in[(((t + n * 7) & 255) * 5 + 4) * 8 + k];
      y[n][k] = in[(((t + n * 13 + 5) & 255) * 5 + 4) * 8 + k];
    }
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int n = 0; n < N; ++n) {
      u32 s[8], d[8], r[8];
      if (MODE == 2) {  // wide-lazy policy (field_wl.cuh)
        add_wl(s, x[n], y[n], f);
        sub_wl_mul(d, x[n], y[n], f);
      } else {
        add_lazy(s, x[n], y[n], f);
        sub_lazy(d, x[n], y[n], f);
      }
      if (MODE) shoup_mul(r, d, w, wq, f); else mont_mul(r, w, d, f);
#pragma unroll
      for (int k = 0; k < 8; ++k) { x[n][k] = s[k]; y[n][k] = r[k]; }
    }
  }
  u32 s = 0;
#pragma unroll
  for (int n = 0; n < N; ++n)
#pragma unroll
    for (int k = 0; k < 8; ++k) s ^= x[n][k] ^ y[n][k];
  if (s == 0x12345678) out[0] = s;
}

static int g_reps = 5;
template <typename F>
static double best_ms(F launch) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  if (g_reps > 1) launch();
  CK(cudaDeviceSynchronize());
  double best = 1e30;
  for (int r = 0; r < g_reps; ++r) {
    cudaEventRecord(a);
    launch();
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    if (ms < best) best = ms;
  }
  CK(cudaGetLastError());
  return best;
}

int main(int argc, char** argv) {
  int iters = argc > 1 ? atoi(argv[1]) : 400;
  if (argc > 2) g_reps = atoi(argv[2]);
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  int sms = prop.multiProcessorCount;
  int clk_khz = 0;
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  printf("{\"device\": \"%s\", \"sms\": %d, \"max_clock_mhz\": %.0f, \"iters\": %d}\n", prop.name, sms, clk_khz / 1000.0, iters);

  u32* d_out; u32* d_in; int* d_bad;
  CK(cudaMalloc(&d_out, 1 << 20));
  CK(cudaMalloc(&d_in, sizeof(kVec)));
  CK(cudaMalloc(&d_bad, 8));
  CK(cudaMemset(d_bad, 0, 8));
  CK(cudaMemcpy(d_in, kVec, sizeof(kVec), cudaMemcpyHostToDevice));

  FieldParams f;
  const u32 p[8] = {0xf0000001u, 0x43e1f593u, 0x79b97091u, 0x2833e848u, 0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
  u64 c = 0, bw = 0;
  for (int k = 0; k < 8; ++k) {
    f.p[k] = p[k];
    u64 t = ((u64)p[k] << 1) | c; f.p2[k] = (u32)t; c = t >> 32;
    u64 d = (u64)0 - (u64)p[k] - bw;  // 2^256 - p, limb by limb
    f.pbar[k] = (u32)d; bw = (p[k] != 0 || bw) ? 1 : 0;
  }
  {
    u64 b2 = 0, c2 = 0;
    for (int k = 0; k < 8; ++k) {
      u64 d = (u64)0 - (u64)f.p2[k] - b2;  // 2^256 - 2p
      f.p2bar[k] = (u32)d; b2 = (f.p2[k] != 0 || b2) ? 1 : 0;
      u64 t = (u64)f.p2[k] + (u64)p[k] + c2;  // 3p
      f.kp[k] = (u32)t; c2 = t >> 32;
    }
  }
  f.np0 = 0xefffffffu;

  k_check<<<(kNumVec + 127) / 128, 128>>>(d_in, kNumVec, f, d_bad);
  int bad[2];
  CK(cudaMemcpy(bad, d_bad, 8, cudaMemcpyDeviceToHost));
  printf("{\"test\": \"known_answers\", \"vectors\": %d, \"shoup_mismatch\": %d, \"mont_mismatch\": %d}\n", kNumVec, bad[0], bad[1]);

  const int threads = 256;
  for (int occ = 1; occ <= 2; ++occ) {
    int blocks = sms * occ;
    double tot = (double)blocks * threads * iters * 4;
    double ms;
    double cyc = 1e-3 * clk_khz * 1e3 / ((double)occ * threads / 32 / 4 * iters * 4);  // SMSP cycles per warp-op per ms
    ms = best_ms([&] { k_mul<4, 0><<<blocks, threads>>>(d_out, d_in, f, iters); });
    printf("{\"test\": \"mont_mul_x4\", \"blocks_per_sm\": %d, \"ms\": %.4f, \"G_per_s\": %.2f, \"smsp_clk_per_warp_op\": %.1f}\n", occ, ms, tot / ms / 1e6, ms * cyc);
    ms = best_ms([&] { k_mul<4, 1><<<blocks, threads>>>(d_out, d_in, f, iters); });
    printf("{\"test\": \"shoup_mul_x4\", \"blocks_per_sm\": %d, \"ms\": %.4f, \"G_per_s\": %.2f, \"smsp_clk_per_warp_op\": %.1f}\n", occ, ms, tot / ms / 1e6, ms * cyc);
    ms = best_ms([&] { k_bfly<4, 0><<<blocks, threads>>>(d_out, d_in, f, iters); });
    printf("{\"test\": \"butterfly_mont_x4\", \"blocks_per_sm\": %d, \"ms\": %.4f, \"G_per_s\": %.2f, \"smsp_clk_per_warp_op\": %.1f}\n", occ, ms, tot / ms / 1e6, ms * cyc);
    ms = best_ms([&] { k_bfly<4, 1><<<blocks, threads>>>(d_out, d_in, f, iters); });
    printf("{\"test\": \"butterfly_shoup_x4\", \"blocks_per_sm\": %d, \"ms\": %.4f, \"G_per_s\": %.2f, \"smsp_clk_per_warp_op\": %.1f}\n", occ, ms, tot / ms / 1e6, ms * cyc);
    ms = best_ms([&] { k_bfly<4, 2><<<blocks, threads>>>(d_out, d_in, f, iters); });
    printf("{\"test\": \"butterfly_shoup_widelazy_x4\", \"blocks_per_sm\": %d, \"ms\": %.4f, \"G_per_s\": %.2f, \"smsp_clk_per_warp_op\": %.1f}\n", occ, ms, tot / ms / 1e6, ms * cyc);
  }
  return 0;
}
