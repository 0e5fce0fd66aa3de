// Round-2 integer-pipe microbenchmark for B200 (sm_100a): cost of every instruction form the
// 256-bit field layer uses, in SMSP issue clocks per warp instruction.
//
// Why a second program: scripts/microbench_pipes.cu (round 1) fed loop-invariant multiplicands to
// `mad.wide.u32`, and ptxas hoisted the product out of the loop -- its "IMAD.WIDE" rows timed 64-bit
// ADDS (cuobjdump: IADD3 + IADD3.X/IMAD.X, no IMAD.WIDE at all).  Here every multiplicand is a
// register written inside the loop, and scripts/check_microbench_sass.py asserts on the SASS of each
// kernel (opcode histogram of the loop body) before the numbers are trusted.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o microbench_issue microbench_issue.cu
// Run:   ./microbench_issue [iters [reps [kernel]]]   (JSON lines; wall time by CUDA events, best of reps)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>
typedef unsigned int u32;
typedef unsigned long long u64;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

// 16 u32 registers r[], 8 doubles d[]; b is a per-thread multiplier that changes every iteration
// (b += r[0]) so that no product is loop-invariant.
#define KB(name)                                                                  \
  __global__ void __launch_bounds__(256) name(u32* out, u32 seed, int iters) {    \
    u32 r[16]; double d[8]; u64 q[8];                                             \
    u32 b = seed * 3 + blockIdx.x + 1;                                            \
    _Pragma("unroll") for (int c = 0; c < 8; ++c) q[c] = (c * 1315423911ull + threadIdx.x) * seed; \
    _Pragma("unroll") for (int c = 0; c < 16; ++c) r[c] = (c * 77 + threadIdx.x) * seed; \
    _Pragma("unroll") for (int c = 0; c < 8; ++c) d[c] = 1.0 + c * 1e-3 + threadIdx.x * 1e-6; \
    double da = 1.0000001, db = 1e-9;                                             \
    for (int i = 0; i < iters; ++i) {                                             \
      b += r[0];
#define KE                                                                        \
    }                                                                             \
    u32 s = 0; double s3 = 0;                                                     \
    _Pragma("unroll") for (int c = 0; c < 16; ++c) s ^= r[c];                     \
    _Pragma("unroll") for (int c = 0; c < 8; ++c) s ^= (u32)q[c] ^ (u32)(q[c] >> 32); \
    _Pragma("unroll") for (int c = 0; c < 8; ++c) s3 += d[c];                     \
    if (s == 0x12345678) out[0] = s + b;                                          \
    if (s3 == 1.2345) out[2] = (u32)da + (u32)db;                                 \
  }
#define R8(X) X(0) X(1) X(2) X(3) X(4) X(5) X(6) X(7)
#define R4(X) X(0) X(1) X(2) X(3)

// ---- single forms.  N(c) = another live register, so every source is loop-variant.
#define N(c) r[((c) + 1) & 7]
#define I_IMAD(c)   asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(r[c]) : "r"(N(c)), "r"(b));
#define I_IMADHI(c) asm volatile("mad.hi.u32 %0, %1, %2, %0;" : "+r"(r[c]) : "r"(N(c)), "r"(b));
// plain IMAD.WIDE.U32: 64-bit accumulator q[c] += lo32(q[c+1]) * b
#define I_WIDE(c)   asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(q[c]) : "r"((u32)q[((c) + 1) & 7]), "r"(b));
#define I_IADD3(c)  asm volatile("add.u32 %0, %0, %1;" : "+r"(r[c]) : "r"(N(c)));
#define I_IADD3_3(c) asm volatile("{.reg .u32 t; add.u32 t, %1, %2; add.u32 %0, %0, t;}" : "+r"(r[c]) : "r"(N(c)), "r"(b));
#define I_LOP3(c)   asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r[c]) : "r"(N(c)), "r"(b));
#define I_SHF(c)    asm volatile("shf.r.wrap.b32 %0, %0, %1, 7;" : "+r"(r[c]) : "r"(N(c)));
#define I_DFMA(c)   asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[c]) : "d"(da), "d"(db));
#define I_SEL(c)    asm volatile("{.reg .pred q; setp.lt.u32 q, %1, %2; selp.u32 %0, %0, %1, q;}" : "+r"(r[c]) : "r"(N(c)), "r"(b));
// ---- carry forms
// 64-bit add: IADD3 (carry out) + IADD3.X
#define I_ADD64(c)  asm volatile("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, %1, %3;" : "+r"(r[2*(c)]), "+r"(r[2*(c)+1]) : "r"(r[(2*(c)+2) & 15]), "r"(b));
// 256-bit add chain: IADD3 + 7 IADD3.X  (r[0..7] += r[8..15], then the other way round)
#define I_ADD256(lo, hi) asm volatile("add.cc.u32 %0, %0, %8;\n\taddc.cc.u32 %1, %1, %9;\n\taddc.cc.u32 %2, %2, %10;\n\taddc.cc.u32 %3, %3, %11;\n\t" \
     "addc.cc.u32 %4, %4, %12;\n\taddc.cc.u32 %5, %5, %13;\n\taddc.cc.u32 %6, %6, %14;\n\taddc.u32 %7, %7, %15;" \
     : "+r"(r[lo]), "+r"(r[lo+1]), "+r"(r[lo+2]), "+r"(r[lo+3]), "+r"(r[lo+4]), "+r"(r[lo+5]), "+r"(r[lo+6]), "+r"(r[lo+7]) \
     : "r"(r[hi]), "r"(r[hi+1]), "r"(r[hi+2]), "r"(r[hi+3]), "r"(r[hi+4]), "r"(r[hi+5]), "r"(r[hi+6]), "r"(r[hi+7]));
// predicated 256-bit add chain (the add-back of sub_lazy)
#define I_PADD256(lo, hi) asm volatile("{.reg .pred q; setp.ne.u32 q, %16, 0;\n\t@q add.cc.u32 %0, %0, %8;\n\t@q addc.cc.u32 %1, %1, %9;\n\t@q addc.cc.u32 %2, %2, %10;\n\t@q addc.cc.u32 %3, %3, %11;\n\t" \
     "@q addc.cc.u32 %4, %4, %12;\n\t@q addc.cc.u32 %5, %5, %13;\n\t@q addc.cc.u32 %6, %6, %14;\n\t@q addc.u32 %7, %7, %15;}" \
     : "+r"(r[lo]), "+r"(r[lo+1]), "+r"(r[lo+2]), "+r"(r[lo+3]), "+r"(r[lo+4]), "+r"(r[lo+5]), "+r"(r[lo+6]), "+r"(r[lo+7]) \
     : "r"(r[hi]), "r"(r[hi+1]), "r"(r[hi+2]), "r"(r[hi+3]), "r"(r[hi+4]), "r"(r[hi+5]), "r"(r[hi+6]), "r"(r[hi+7]), "r"(b & 1));
// two wide MADs linked by the carry flag: IMAD.WIDE.U32 (carry out) + IMAD.WIDE.U32.X (carry in)
#define I_WIDE_CC2(c) asm volatile("mad.lo.cc.u32 %0, %4, %5, %0;\n\tmadc.hi.cc.u32 %1, %4, %5, %1;\n\tmadc.lo.cc.u32 %2, %6, %5, %2;\n\tmadc.hi.u32 %3, %6, %5, %3;" \
     : "+r"(r[4*(c)]), "+r"(r[4*(c)+1]), "+r"(r[4*(c)+2]), "+r"(r[4*(c)+3]) : "r"(r[(4*(c)+4) & 15]), "r"(b), "r"(r[(4*(c)+6) & 15]));
// wide MAD with carry-out only, the carry consumed by one IADD3.X
#define I_WIDE_COUT(c) asm volatile("mad.lo.cc.u32 %0, %3, %4, %0;\n\tmadc.hi.cc.u32 %1, %3, %4, %1;\n\taddc.u32 %2, %2, 0;" \
     : "+r"(r[4*(c)]), "+r"(r[4*(c)+1]), "+r"(r[4*(c)+2]) : "r"(r[(4*(c)+4) & 15]), "r"(b));
// IADD3 (carry out) feeding one wide MAD with carry-in only
#define I_WIDE_CIN(c) asm volatile("add.cc.u32 %0, %0, %4;\n\tmadc.lo.cc.u32 %1, %3, %4, %1;\n\tmadc.hi.u32 %2, %3, %4, %2;" \
     : "+r"(r[4*(c)]), "+r"(r[4*(c)+1]), "+r"(r[4*(c)+2]) : "r"(r[(4*(c)+4) & 15]), "r"(b));
// the library's row: acc[0..7] += (a0,a2,a4,a6)*b as one 4-long chain + addc (exactly mad_row of field.cuh)
#define I_ROW4(lo, hi) asm volatile("mad.lo.cc.u32 %0, %9, %13, %0;\n\tmadc.hi.cc.u32 %1, %9, %13, %1;\n\tmadc.lo.cc.u32 %2, %10, %13, %2;\n\tmadc.hi.cc.u32 %3, %10, %13, %3;\n\t" \
     "madc.lo.cc.u32 %4, %11, %13, %4;\n\tmadc.hi.cc.u32 %5, %11, %13, %5;\n\tmadc.lo.cc.u32 %6, %12, %13, %6;\n\tmadc.hi.cc.u32 %7, %12, %13, %7;\n\taddc.u32 %8, %8, 0;" \
     : "+r"(r[lo]), "+r"(r[lo+1]), "+r"(r[lo+2]), "+r"(r[lo+3]), "+r"(r[lo+4]), "+r"(r[lo+5]), "+r"(r[lo+6]), "+r"(r[lo+7]), "+r"(b) \
     : "r"(r[hi]), "r"(r[hi+2]), "r"(r[hi+4]), "r"(r[hi+6]), "r"(r[hi+1]));
// the same four products with NO carry links (what a carry-free layout would issue): 4 plain IMAD.WIDE
#define I_ROW4_PLAIN(lo, hi) I_WIDEP(lo, hi) I_WIDEP(lo+2, hi+2) I_WIDEP(lo+4, hi+4) I_WIDEP(lo+6, hi+6)
#define I_WIDEP(x, y) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(q[(x) / 2]) : "r"((u32)q[(y) / 2]), "r"((u32)(q[(y) / 2] >> 32)));
// 8-long chain over 16 limbs
#define I_ROW8() asm volatile("mad.lo.cc.u32 %0, %9, %16, %0;\n\tmadc.hi.cc.u32 %1, %9, %16, %1;\n\tmadc.lo.cc.u32 %2, %11, %16, %2;\n\tmadc.hi.cc.u32 %3, %11, %16, %3;\n\t" \
     "madc.lo.cc.u32 %4, %13, %16, %4;\n\tmadc.hi.cc.u32 %5, %13, %16, %5;\n\tmadc.lo.cc.u32 %6, %15, %16, %6;\n\tmadc.hi.cc.u32 %7, %15, %16, %7;\n\t" \
     "madc.lo.cc.u32 %8, %1, %16, %8;\n\tmadc.hi.cc.u32 %9, %1, %16, %9;\n\tmadc.lo.cc.u32 %10, %3, %16, %10;\n\tmadc.hi.cc.u32 %11, %3, %16, %11;\n\t" \
     "madc.lo.cc.u32 %12, %5, %16, %12;\n\tmadc.hi.cc.u32 %13, %5, %16, %13;\n\tmadc.lo.cc.u32 %14, %7, %16, %14;\n\tmadc.hi.u32 %15, %7, %16, %15;" \
     : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), \
       "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]) : "r"(b));

// fused wide MAD with accumulate, no carry in or out: (r[2c], r[2c+1]) += r[2c+2] * b  -> IMAD.WIDE.U32 R, a, b, R
#define I_WIDE_ACC(c) asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.u32 %1, %2, %3, %1;" \
     : "+r"(r[2*(c)]), "+r"(r[2*(c)+1]) : "r"(r[(2*(c)+2) & 15]), "r"(b));
KB(k_wide_acc) R8(I_WIDE_ACC) KE
KB(k_imad) R8(I_IMAD) KE
KB(k_imadhi) R8(I_IMADHI) KE
KB(k_wide) R8(I_WIDE) KE
KB(k_iadd) R8(I_IADD3) KE
KB(k_iadd3) R8(I_IADD3_3) KE
KB(k_lop3) R8(I_LOP3) KE
KB(k_shf) R8(I_SHF) KE
KB(k_dfma) R8(I_DFMA) KE
KB(k_sel) R8(I_SEL) KE
KB(k_add64) R8(I_ADD64) KE
KB(k_add256) I_ADD256(0, 8) I_ADD256(8, 0) KE
KB(k_padd256) I_PADD256(0, 8) I_PADD256(8, 0) KE
KB(k_wide_cc2) R4(I_WIDE_CC2) KE
KB(k_wide_cout) R4(I_WIDE_COUT) KE
KB(k_wide_cin) R4(I_WIDE_CIN) KE
KB(k_row4) I_ROW4(0, 8) I_ROW4(8, 0) KE
KB(k_row4_plain) I_ROW4_PLAIN(0, 8) I_ROW4_PLAIN(8, 0) KE
KB(k_row8) I_ROW8() KE
// ---- mixes: does ALU-pipe work hide under multiplier-pipe work?
#define M_WIDE_IADD(c) I_WIDE(c) I_IADD3_3(c)
#define M_WIDE_LOP(c) I_WIDE(c) I_LOP3(c)
#define M_IMAD_LOP(c) I_IMAD(c) I_LOP3(((c) + 4) & 7)
#define M_WIDE_DFMA(c) I_WIDE(c) I_DFMA(c)
KB(k_wide_iadd3) R8(M_WIDE_IADD) KE
KB(k_wide_lop3) R8(M_WIDE_LOP) KE
KB(k_imad_lop3) R8(M_IMAD_LOP) KE
KB(k_wide_dfma) R8(M_WIDE_DFMA) KE
KB(k_row4_add256) I_ROW4(0, 8) I_ADD256(8, 0) KE
KB(k_row4_lop3) I_ROW4(0, 8) I_LOP3(8) I_LOP3(9) I_LOP3(10) I_LOP3(11) I_LOP3(12) I_LOP3(13) I_LOP3(14) I_LOP3(15) KE
KB(k_row4_dfma) I_ROW4(0, 8) I_ROW4(8, 0) R8(I_DFMA) KE

typedef void (*kern_t)(u32*, u32, int);
static int g_iters = 4000;
static int g_reps = 5;

static void run(const char* name, kern_t k, int per_iter, const char* what, int sms, double clk_mhz, u32* d_out, int bps = 8) {
  const int threads = 256;
  int blocks = sms * bps;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  if (g_reps > 1) k<<<blocks, threads>>>(d_out, 12345, g_iters);  // warm-up (skipped under ncu: reps = 1)
  CK(cudaDeviceSynchronize());
  double best = 1e30;
  for (int rep = 0; rep < g_reps; ++rep) {
    cudaEventRecord(e0);
    k<<<blocks, threads>>>(d_out, 12345, g_iters);
    cudaEventRecord(e1);
    CK(cudaEventSynchronize(e1));
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  // SMSP issue clocks per warp: each SMSP runs bps*threads/32/4 warps, each executing iters*per_iter counted instructions
  double warps_per_smsp = (double)bps * threads / 32 / 4;
  double clk_per_iter = best * 1e-3 * clk_mhz * 1e6 / (warps_per_smsp * g_iters);
  double tot = (double)blocks * threads * g_iters * per_iter;
  printf("{\"test\": \"%s\", \"counted_instr_per_iter\": %d, \"ms\": %.4f, \"T_thread_instr_s\": %.3f, \"smsp_clk_per_warp_iter\": %.2f, "
         "\"smsp_clk_per_counted_instr\": %.3f, \"warps_per_smsp\": %.0f, \"what\": \"%s\"}\n",
         name, per_iter, best, tot / best / 1e9, clk_per_iter, clk_per_iter / per_iter, warps_per_smsp, what);
  fflush(stdout);
}

int main(int argc, char** argv) {
  if (argc > 1) g_iters = atoi(argv[1]);
  if (argc > 2) g_reps = atoi(argv[2]);
  const char* only = argc > 3 ? argv[3] : nullptr;
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  int sms = prop.multiProcessorCount;
  int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  double mhz = clk_khz / 1000.0;
  printf("{\"device\": \"%s\", \"sms\": %d, \"clock_mhz_assumed\": %.0f, \"iters\": %d, \"note\": \"smsp_clk figures assume the max SM clock; bench.py-style nvidia-smi sampling is in the gpurun script\"}\n", prop.name, sms, mhz, g_iters);
  u32* d_out; CK(cudaMalloc(&d_out, 4096));
#define RUN(k, n, what) if (!only || !strcmp(only, #k)) run(#k, k, n, what, sms, mhz, d_out)
  RUN(k_imad, 8, "8 IMAD");
  RUN(k_imadhi, 8, "8 IMAD.HI.U32");
  RUN(k_wide, 8, "8 plain IMAD.WIDE.U32 (multiplicand written in the loop)");
  RUN(k_wide_acc, 8, "8 IMAD.WIDE.U32 with 64-bit accumulate, no carry (mad.lo.cc + madc.hi fused)");
  RUN(k_iadd, 8, "8 two-input adds (ptxas picks IADD3 / IMAD.IADD / VIADD)");
  RUN(k_iadd3, 8, "8 three-input adds (IADD3)");
  RUN(k_lop3, 8, "8 LOP3");
  RUN(k_shf, 8, "8 SHF");
  RUN(k_dfma, 8, "8 DFMA");
  RUN(k_sel, 8, "8 (ISETP + SEL)");
  RUN(k_add64, 16, "8 x (IADD3 carry-out + IADD3.X)");
  RUN(k_add256, 16, "2 x 256-bit add chain (IADD3 + 7 IADD3.X)");
  RUN(k_padd256, 16, "2 x predicated 256-bit add chain");
  RUN(k_wide_cc2, 8, "4 x (IMAD.WIDE carry-out + IMAD.WIDE.X carry-in)");
  RUN(k_wide_cout, 8, "4 x (IMAD.WIDE carry-out + IADD3.X)");
  RUN(k_wide_cin, 8, "4 x (IADD3 carry-out + IMAD.WIDE.X carry-in only)");
  RUN(k_row4, 8, "2 x field.cuh mad_row: IMAD.WIDE carry-out + 3 IMAD.WIDE.X + IADD3.X (counted: the 8 wide MADs)");
  RUN(k_row4_plain, 8, "the same 8 products as 8 unlinked plain IMAD.WIDE");
  RUN(k_row8, 8, "one 8-long chain: IMAD.WIDE carry-out + 7 IMAD.WIDE.X");
  RUN(k_wide_iadd3, 16, "8 plain IMAD.WIDE + 8 IADD3");
  RUN(k_wide_lop3, 16, "8 plain IMAD.WIDE + 8 LOP3");
  RUN(k_imad_lop3, 16, "8 IMAD + 8 LOP3");
  RUN(k_wide_dfma, 16, "8 plain IMAD.WIDE + 8 DFMA");
  RUN(k_row4_add256, 12, "mad_row (4 wide) + 256-bit add chain (8 adds)");
  RUN(k_row4_lop3, 12, "mad_row (4 wide) + 8 LOP3");
  RUN(k_row4_dfma, 16, "2 mad_row (8 wide) + 8 DFMA");
  return 0;
}
