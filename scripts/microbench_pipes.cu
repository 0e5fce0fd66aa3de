// Pipe-level microbenchmark for B200 (sm_100a): which integer instructions run at full
// rate, which at half rate, and what co-issues.  Decides the multiplier design in
// ntt_b200/csrc/field.cuh.  Every test: 8 independent chains per thread, 256 threads,
// 8 blocks/SM, result in thread-instructions per SM per clock at the measured SM clock.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
typedef unsigned int u32;
typedef unsigned long long u64;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

#define KERNEL_BEGIN(name)                                                      \
  __global__ void __launch_bounds__(256) name(u32* out, u32 seed, int iters, long long* cyc) { \
    u32 r[16]; u64 q[8]; double d[8];                                           \
    u32 a = seed + threadIdx.x, b = seed * 3 + blockIdx.x + 1;                  \
    _Pragma("unroll") for (int c = 0; c < 16; ++c) r[c] = c * 77 + threadIdx.x; \
    _Pragma("unroll") for (int c = 0; c < 8; ++c) { q[c] = c * 1315423911ull + threadIdx.x; d[c] = 1.0 + c * 1e-3 + threadIdx.x * 1e-6; } \
    double da = 1.0000001, db = 1e-9;                                           \
    long long t0 = clock64();                                                   \
    for (int i = 0; i < iters; ++i) {
#define KERNEL_END                                                              \
    }                                                                           \
    long long t1 = clock64();                                                   \
    u32 s = 0; u64 s2 = 0; double s3 = 0;                                       \
    _Pragma("unroll") for (int c = 0; c < 16; ++c) s ^= r[c];                   \
    _Pragma("unroll") for (int c = 0; c < 8; ++c) { s2 ^= q[c]; s3 += d[c]; }   \
    if (s == 0x12345678) out[0] = s + a + b;                                   \
    if (s2 == 0x1234567812345678ull) out[1] = (u32)s2;                          \
    if (s3 == 1.2345) out[2] = (u32)da + (u32)db;                               \
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;                            \
  }

#define REP8(X) X(0) X(1) X(2) X(3) X(4) X(5) X(6) X(7)

// --- single instruction streams
#define I_IMAD(c)  asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(r[c]) : "r"(a), "r"(b));
#define I_IMADHI(c) asm volatile("mad.hi.u32 %0, %1, %2, %0;" : "+r"(r[c]) : "r"(a), "r"(b));
#define I_WIDE(c)  asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(q[c]) : "r"(a), "r"(b));
#define I_WIDES(c) asm volatile("mad.wide.s32 %0, %1, %2, %0;" : "+l"(q[c]) : "r"(a), "r"(b));
#define I_IADD3(c) asm volatile("add.u32 %0, %0, %1;" : "+r"(r[c]) : "r"(a));
#define I_LOP3(c)  asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r[c]) : "r"(a), "r"(b));
#define I_SHF(c)   asm volatile("shf.r.wrap.b32 %0, %0, %1, 7;" : "+r"(r[c]) : "r"(a));
#define I_DFMA(c)  asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[c]) : "d"(da), "d"(db));
// 64-bit add as a 2-instruction carry chain (IADD3 + IADD3.X)
#define I_ADD64(c) asm volatile("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, %1, %3;" : "+r"(r[2*(c)]), "+r"(r[2*(c)+1]) : "r"(a), "r"(b));
// wide mad pair chained through the carry flag: first has carry-out, second carry-in
#define I_WIDE_CC2(c) asm volatile("mad.lo.cc.u32 %0, %4, %5, %0;\n\tmadc.hi.cc.u32 %1, %4, %5, %1;\n\tmadc.lo.cc.u32 %2, %4, %5, %2;\n\tmadc.hi.u32 %3, %4, %5, %3;" \
     : "+r"(r[4*((c)&3)]), "+r"(r[4*((c)&3)+1]), "+r"(r[4*((c)&3)+2]), "+r"(r[4*((c)&3)+3]) : "r"(a), "r"(b));
// wide mad with carry-out only, carry consumed by an IADD3.X
#define I_WIDE_COUT(c) asm volatile("mad.lo.cc.u32 %0, %3, %4, %0;\n\tmadc.hi.cc.u32 %1, %3, %4, %1;\n\taddc.u32 %2, %2, 0;" \
     : "+r"(r[4*((c)&3)]), "+r"(r[4*((c)&3)+1]), "+r"(r[4*((c)&3)+2]) : "r"(a), "r"(b));

KERNEL_BEGIN(k_imad) REP8(I_IMAD) KERNEL_END
KERNEL_BEGIN(k_imadhi) REP8(I_IMADHI) KERNEL_END
KERNEL_BEGIN(k_wide) REP8(I_WIDE) KERNEL_END
KERNEL_BEGIN(k_wides) REP8(I_WIDES) KERNEL_END
KERNEL_BEGIN(k_iadd3) REP8(I_IADD3) KERNEL_END
KERNEL_BEGIN(k_lop3) REP8(I_LOP3) KERNEL_END
KERNEL_BEGIN(k_shf) REP8(I_SHF) KERNEL_END
KERNEL_BEGIN(k_dfma) REP8(I_DFMA) KERNEL_END
KERNEL_BEGIN(k_add64) REP8(I_ADD64) KERNEL_END
KERNEL_BEGIN(k_wide_cc2) REP8(I_WIDE_CC2) KERNEL_END
KERNEL_BEGIN(k_wide_cout) REP8(I_WIDE_COUT) KERNEL_END
// --- mixes (count = thread-instructions of both kinds)
#define I_WIDE_IADD(c) I_WIDE(c) I_IADD3(c)
#define I_WIDE_IADD2(c) I_WIDE(c) I_IADD3(c) I_LOP3(c + 8)
#define I_WIDE_ADD64(c) I_WIDE(c) I_ADD64(c)
#define I_WIDE_DFMA(c) I_WIDE(c) I_DFMA(c)
#define I_WIDE_DFMA_IADD(c) I_WIDE(c) I_DFMA(c) I_IADD3(c)
KERNEL_BEGIN(k_wide_iadd) REP8(I_WIDE_IADD) KERNEL_END
KERNEL_BEGIN(k_wide_iadd2) REP8(I_WIDE_IADD2) KERNEL_END
KERNEL_BEGIN(k_wide_add64) REP8(I_WIDE_ADD64) KERNEL_END
KERNEL_BEGIN(k_wide_dfma) REP8(I_WIDE_DFMA) KERNEL_END
KERNEL_BEGIN(k_wide_dfma_iadd) REP8(I_WIDE_DFMA_IADD) KERNEL_END

typedef void (*kern_t)(u32*, u32, int, long long*);

static void run(const char* name, kern_t k, int per_iter, int sms, u32* d_out, long long* d_cyc) {
  const int iters = 4000, threads = 256, bps = 8;
  int blocks = sms * bps;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<<<blocks, threads>>>(d_out, 12345, iters, d_cyc);
  CK(cudaDeviceSynchronize());
  double best = 1e30;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0);
    k<<<blocks, threads>>>(d_out, 12345, iters, d_cyc);
    cudaEventRecord(e1);
    CK(cudaEventSynchronize(e1));
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  long long* h = (long long*)malloc(blocks * sizeof(long long));
  CK(cudaMemcpy(h, d_cyc, blocks * sizeof(long long), cudaMemcpyDeviceToHost));
  double cyc = 0; for (int i = 0; i < blocks; ++i) cyc += h[i]; cyc /= blocks;
  free(h);
  // all 8 blocks of an SM run concurrently: per-SM work = bps*threads*iters*per_iter during ~cyc clocks
  double per_sm_clk = (double)bps * threads * iters * per_iter / cyc;
  double tot = (double)blocks * threads * iters * per_iter;
  printf("{\"test\": \"%s\", \"instr_per_iter\": %d, \"ms\": %.4f, \"Ginstr_s\": %.0f, \"thread_instr_per_sm_per_clk\": %.1f, \"sm_mhz_eff\": %.0f}\n",
         name, per_iter, best, tot / best / 1e6, per_sm_clk, cyc / (best * 1e3));
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  int sms = prop.multiProcessorCount;
  u32* d_out; long long* d_cyc;
  CK(cudaMalloc(&d_out, 4096)); CK(cudaMalloc(&d_cyc, sms * 8 * sizeof(long long)));
  run("imad_lo", k_imad, 8, sms, d_out, d_cyc);
  run("imad_hi", k_imadhi, 8, sms, d_out, d_cyc);
  run("imad_wide_u32", k_wide, 8, sms, d_out, d_cyc);
  run("imad_wide_s32", k_wides, 8, sms, d_out, d_cyc);
  run("iadd3", k_iadd3, 8, sms, d_out, d_cyc);
  run("lop3", k_lop3, 8, sms, d_out, d_cyc);
  run("shf", k_shf, 8, sms, d_out, d_cyc);
  run("dfma", k_dfma, 8, sms, d_out, d_cyc);
  run("add64_iadd3_x_pair", k_add64, 16, sms, d_out, d_cyc);
  run("imad_wide_cc_pair(2 wide)", k_wide_cc2, 16, sms, d_out, d_cyc);
  run("imad_wide_cout+iadd3x(1 wide+1 add)", k_wide_cout, 16, sms, d_out, d_cyc);
  run("mix_wide+iadd3", k_wide_iadd, 16, sms, d_out, d_cyc);
  run("mix_wide+iadd3+lop3", k_wide_iadd2, 24, sms, d_out, d_cyc);
  run("mix_wide+add64", k_wide_add64, 24, sms, d_out, d_cyc);
  run("mix_wide+dfma", k_wide_dfma, 16, sms, d_out, d_cyc);
  run("mix_wide+dfma+iadd3", k_wide_dfma_iadd, 24, sms, d_out, d_cyc);
  return 0;
}
