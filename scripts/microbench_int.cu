// Integer-pipe microbenchmark for B200 (sm_100a): what is the real ceiling for a
// 256-bit Montgomery product?  Measures, per SM per clock and chip-wide per second:
//   (1) independent mad.lo.u32            (IMAD)
//   (2) independent mad.wide.u32          (IMAD.WIDE.U32, 64-bit accumulate)
//   (3) mad.lo.cc/madc.hi.cc chains       (IMAD.WIDE.U32.X, carry in/out)
//   (4) the library's mont_mul()          (128 wide + 8 low multiplies)
//   (5) one DIF butterfly (add, sub, mont_mul)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o microbench_int microbench_int.cu
// The figure from (4) is the INT roofline denominator quoted in DESIGN.md / bench.py.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "../ntt_b200/csrc/field.cuh"

using namespace ntt256;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

template <int CHAINS>
__global__ void k_imad(u32* out, u32 seed, int iters) {
  u32 acc[CHAINS];
  u32 a = seed + threadIdx.x, b = seed * 3 + blockIdx.x;
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) acc[c] = c + threadIdx.x;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(acc[c]) : "r"(a), "r"(b));
  }
  u32 s = 0;
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) s ^= acc[c];
  if (s == 0x12345678) out[0] = s;
}

template <int CHAINS>
__global__ void k_imad_wide(u32* out, u32 seed, int iters) {
  u64 acc[CHAINS];
  u32 a = seed + threadIdx.x, b = seed * 3 + blockIdx.x;
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) acc[c] = c + threadIdx.x;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[c]) : "r"(a), "r"(b));
  }
  u64 s = 0;
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) s ^= acc[c];
  if (s == 0x12345678) out[0] = (u32)s;
}

// CHAINS independent 8-limb accumulators, each updated by one carry chain of 4 wide MADs.
template <int CHAINS>
__global__ void k_imad_chain(u32* out, u32 seed, int iters) {
  u32 acc[CHAINS][8];
  u32 a[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) a[k] = seed * (k + 1) + threadIdx.x;
  u32 b = seed * 3 + blockIdx.x;
#pragma unroll
  for (int c = 0; c < CHAINS; ++c)
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[c][k] = c + k + threadIdx.x;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) (void)mad_row(acc[c], a, b);
  }
  u32 s = 0;
#pragma unroll
  for (int c = 0; c < CHAINS; ++c)
#pragma unroll
    for (int k = 0; k < 8; ++k) s ^= acc[c][k];
  if (s == 0x12345678) out[0] = s;
}

template <int N>
__global__ void __launch_bounds__(256) k_mont(u32* out, const u32* in, const __grid_constant__ FieldParams f, int iters) {
  u32 a[8], b[N][8];
  int t = blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll
  for (int k = 0; k < 8; ++k) a[k] = in[(t & 1023) * 8 + k] & (k == 7 ? 0x0fffffffu : 0xffffffffu);
#pragma unroll
  for (int n = 0; n < N; ++n)
#pragma unroll
    for (int k = 0; k < 8; ++k) b[n][k] = in[((t + n * 7) & 1023) * 8 + k];
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int n = 0; n < N; ++n) {
      u32 r[8];
      mont_mul(r, a, b[n], f);
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

template <int N>
__global__ void __launch_bounds__(256) k_bfly(u32* out, const u32* in, const __grid_constant__ FieldParams f, int iters) {
  u32 w[8], x[N][8], y[N][8];
  int t = blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll
  for (int k = 0; k < 8; ++k) w[k] = in[(t & 1023) * 8 + k] & (k == 7 ? 0x0fffffffu : 0xffffffffu);
#pragma unroll
  for (int n = 0; n < N; ++n)
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      x[n][k] = in[((t + n * 7) & 1023) * 8 + k] & (k == 7 ? 0x1fffffffu : 0xffffffffu);
      y[n][k] = in[((t + n * 13 + 5) & 1023) * 8 + k] & (k == 7 ? 0x1fffffffu : 0xffffffffu);
    }
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int n = 0; n < N; ++n) {
      u32 s[8], d[8], r[8];
      add_lazy(s, x[n], y[n], f);
      sub_lazy(d, x[n], y[n], f);
      mont_mul(r, w, d, f);
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

struct Timer {
  cudaEvent_t a, b;
  Timer() { cudaEventCreate(&a); cudaEventCreate(&b); }
  void start() { cudaEventRecord(a); }
  float stop() { cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); return ms; }
};

template <typename F>
static double best_ms(F launch, int reps = 5) {
  Timer t;
  launch();
  CK(cudaDeviceSynchronize());
  double best = 1e30;
  for (int r = 0; r < reps; ++r) {
    t.start();
    launch();
    double ms = t.stop();
    if (ms < best) best = ms;
  }
  CK(cudaGetLastError());
  return best;
}

int main() {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  int sms = prop.multiProcessorCount;
  int clk_khz = 0;
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  printf("{\"device\": \"%s\", \"sms\": %d, \"max_clock_mhz\": %.0f}\n", prop.name, sms, clk_khz / 1000.0);

  u32* d_out; u32* d_in;
  CK(cudaMalloc(&d_out, 1 << 20));
  CK(cudaMalloc(&d_in, 1024 * 8 * 4));
  std::vector<u32> h(1024 * 8);
  u64 s = 0x9e3779b97f4a7c15ull;
  for (auto& v : h) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; v = (u32)(s >> 16); }
  CK(cudaMemcpy(d_in, h.data(), h.size() * 4, cudaMemcpyHostToDevice));

  // BN254 Fr
  FieldParams f;
  const u32 p[8] = {0xf0000001u, 0x43e1f593u, 0x79b97091u, 0x2833e848u, 0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
  u64 c = 0;
  for (int k = 0; k < 8; ++k) { f.p[k] = p[k]; u64 t = ((u64)p[k] << 1) | c; f.p2[k] = (u32)t; c = t >> 32; }
  f.np0 = 0xefffffffu;

  const int threads = 256;
  const int iters = 2000;
  // resident: blocks per SM chosen so that 8 warps/SMSP are present for the pure pipes
  {
    int blocks = sms * 8;
    double tot = (double)blocks * threads * iters * 8;
    double ms = best_ms([&] { k_imad<8><<<blocks, threads>>>(d_out, 12345, iters); });
    printf("{\"test\": \"imad_lo_independent\", \"ms\": %.4f, \"Gops\": %.1f, \"per_sm_per_clk_at_max\": %.2f}\n", ms, tot / ms / 1e6, tot / ms / 1e3 / sms / (clk_khz));
    ms = best_ms([&] { k_imad_wide<8><<<blocks, threads>>>(d_out, 12345, iters); });
    printf("{\"test\": \"imad_wide_independent\", \"ms\": %.4f, \"Gops\": %.1f, \"per_sm_per_clk_at_max\": %.2f}\n", ms, tot / ms / 1e6, tot / ms / 1e3 / sms / (clk_khz));
  }
  {
    int blocks = sms * 4;
    double tot = (double)blocks * threads * iters * 4 /*chains*/ * 4 /*wide mads*/;
    double ms = best_ms([&] { k_imad_chain<4><<<blocks, threads>>>(d_out, 12345, iters); });
    printf("{\"test\": \"imad_wide_carry_chain\", \"ms\": %.4f, \"Gops\": %.1f, \"per_sm_per_clk_at_max\": %.2f}\n", ms, tot / ms / 1e6, tot / ms / 1e3 / sms / (clk_khz));
  }
  for (int occ = 1; occ <= 3; ++occ) {
    int blocks = sms * occ;
    const int it2 = 400;
    {
      double tot = (double)blocks * threads * it2 * 2;
      double ms = best_ms([&] { k_mont<2><<<blocks, threads>>>(d_out, d_in, f, it2); });
      printf("{\"test\": \"mont_mul_x2\", \"blocks_per_sm\": %d, \"ms\": %.4f, \"Gmodmul\": %.2f, \"wide_mac_per_sm_per_clk_at_max\": %.2f}\n", occ, ms, tot / ms / 1e6, tot * 136 / ms / 1e3 / sms / clk_khz);
    }
    {
      double tot = (double)blocks * threads * it2 * 4;
      double ms = best_ms([&] { k_mont<4><<<blocks, threads>>>(d_out, d_in, f, it2); });
      printf("{\"test\": \"mont_mul_x4\", \"blocks_per_sm\": %d, \"ms\": %.4f, \"Gmodmul\": %.2f, \"wide_mac_per_sm_per_clk_at_max\": %.2f}\n", occ, ms, tot / ms / 1e6, tot * 136 / ms / 1e3 / sms / clk_khz);
    }
    {
      double tot = (double)blocks * threads * it2 * 4;
      double ms = best_ms([&] { k_bfly<4><<<blocks, threads>>>(d_out, d_in, f, it2); });
      printf("{\"test\": \"butterfly_x4\", \"blocks_per_sm\": %d, \"ms\": %.4f, \"Gbutterfly\": %.2f}\n", occ, ms, tot / ms / 1e6);
    }
  }
  return 0;
}
