// C ABI of the 64-bit transforms (include/ntt64.h): the reference's `bits` template parameter
// (/root/reference/src/big-num.cu:57,67,172,260) at 64 -- the width its own CPU NTT runs in (big-num.cu:37-55).
// Same planner as the wide libraries (plan.cpp), kernels ntt_kernels_64.cuh, field layer field64.cuh.
// Host side: field constants and table seeds (64-bit modular arithmetic on unsigned __int128), plan and kernel
// shape by call size, device tables built on the device, 1/n folded into a round-twiddle table, coset shifts
// fused into the end passes, pipelined host-buffer batches.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstring>
#include <map>
#include <new>
#include <string>
#include <tuple>
#include <vector>

#include "../../include/ntt64.h"
#include "host_common.h"
#include "ntt_kernels_64.cuh"
#include "pass_tables.h"
#include "plan.h"

using namespace ntt64;
using ntt256::Plan;
using ntt256::PlanRequest;
using ntt256::make_plan;

namespace {

typedef unsigned __int128 u128;
constexpr int kTileBits64 = 12;   // 2^12 x 8 B = 32 KiB of shared memory, 256 threads, four CTAs per SM
constexpr int kCoalesce64 = 4;    // runs of >= 16 elements = 128 bytes
constexpr int kHiBits64 = 20;     // 2^20 x 8 B = 8 MiB: stays in L2
constexpr int kLocBits64 = 12;
constexpr int kMaxTile64 = 14;    // 128 KiB of shared memory, 1024 threads
constexpr size_t kCosetDirectMax64 = (size_t)256 << 20;  // one multiplied-out shift^j table (2^25 points)

struct HostField64 {
  u64 p = 0;
  bool gl = false;
  u64 np = 0;   // -p^-1 mod 2^64
  u64 r = 0;    // 2^64 mod p
  u64 r2 = 0;   // 2^128 mod p
  void init(u64 P) {
    p = P;
    gl = (P == kGoldilocksP);
    u64 inv = 1;  // Newton: p^-1 mod 2^64
    for (int i = 0; i < 6; ++i) inv *= 2 - P * inv;
    np = 0 - inv;
    r = (u64)(((u128)1 << 64) % P);
    r2 = (u64)(((u128)r * r) % P);
  }
  u64 mul(u64 a, u64 b) const { return (u64)(((u128)a * b) % p); }
  u64 pow(u64 b, u64 e) const {
    u64 acc = 1 % p;
    while (e) {
      if (e & 1) acc = mul(acc, b);
      b = mul(b, b);
      e >>= 1;
    }
    return acc;
  }
  u64 inv(u64 a) const { return pow(a, p - 2); }
  // what a twiddle table holds for the plain value w: w itself (Goldilocks) or its Montgomery form w * 2^64
  u64 tab(u64 w) const { return gl ? w : mul(w, r); }
};

struct PlanKey64 {
  int log_n, tile_bits, group_bits, coalesce;
  int kind = 0, log_r = 0, log_p = 0, rank = 0;  // kind 0: whole transform; 1..4: four-step pieces (plan.h FourStepKey)
  bool operator<(const PlanKey64& o) const {
    return std::tie(log_n, tile_bits, group_bits, coalesce, kind, log_r, log_p, rank) <
           std::tie(o.log_n, o.tile_bits, o.group_bits, o.coalesce, o.kind, o.log_r, o.log_p, o.rank);
  }
};

}  // namespace

struct ntt64_ctx {
  int device = 0;
  int max_log_n = 0;
  HostField64 hf;
  Params64 fp;
  u64 wN = 0, wN_inv = 0;  // plain
  u64 *om_hi = nullptr, *om_lo = nullptr, *loc = nullptr;
  int hi_bits = 0, lo_bits = 0, loc_bits = 0;
  u64 c16[2][7];  // table format
  int k16[2] = {0, 0};  // Goldilocks: w_16^(+-1) = 2^(12 k16) when k16 is one of the compiled-in values (13, 3), else 0
  u64 w4[2];
  int tile_bits = kTileBits64, coalesce_bits = kCoalesce64;
  bool auto_shape = true;
  bool use_shift_twiddles = true;  // testing knob: false keeps the constant twiddles on the multiplier
  size_t coset_direct_max = kCosetDirectMax64;  // largest multiplied-out shift^j table; longer transforms use two factored ones
  typedef ntt256::PassTables PlanTables;  // per pass: the index tables of PassArgs64 (pass_tables.h)
  std::map<PlanKey64, Plan> plans;
  std::map<PlanKey64, PlanTables> plan_tables;
  std::map<int, u64*> loc_scaled;  // log_n -> (1/n) * loc[]
  struct Coset {
    u64 shift;     // plain
    int log_n;
    int with_ninv; // the table carries 1/n as well (inverse coset transforms)
    u64* direct;   // shift^j (times 1/n), j < 2^log_n, table format; or null (too large)
    u64 *hi, *lo;  // factored tables when direct is null
    int lo_bits;
  };
  std::vector<Coset> cosets;
  void* d_stage = nullptr;
  size_t stage_bytes = 0;
  cudaStream_t s_h2d = nullptr, s_d2h = nullptr, s_comp = nullptr;
  cudaEvent_t ev_in[3] = {nullptr, nullptr, nullptr}, ev_done[3] = {nullptr, nullptr, nullptr},
              ev_out[3] = {nullptr, nullptr, nullptr};
  unsigned long long launches = 0;
  std::string last_err;
};

namespace {

#define CTX_CUDA(call)                                                     \
  do {                                                                     \
    cudaError_t e_ = (call);                                               \
    if (e_ != cudaSuccess) {                                               \
      ctx->last_err = std::string(#call) + ": " + cudaGetErrorString(e_);  \
      return NTT256_ERR_CUDA;                                              \
    }                                                                      \
  } while (0)

u64 pow2k(const HostField64& hf, u64 b, int k) {  // b^(2^k)
  for (int i = 0; i < k; ++i) b = hf.mul(b, b);
  return b;
}

template <bool GL>
void launch_powers(u64* out, const u64* d_pw, int bits, unsigned long long count, const ntt64_ctx* ctx, u64 one_tab,
                   cudaStream_t stream) {
  const int threads = 256;
  twiddle_powers_kernel64<GL><<<(unsigned)((count + threads - 1) / threads), threads, 0, stream>>>(
      out, d_pw, bits, count, ctx->fp, one_tab);
}

// out[i] = start * base^i, i < 2^bits, in table format (start and base plain)
int build_power_table(ntt64_ctx* ctx, u64** out, u64 base, int bits, u64 start = 1, cudaStream_t stream = 0) {
  const unsigned long long count = 1ull << bits;
  *out = nullptr;
  CTX_CUDA(cudaMalloc((void**)out, count * 8));
  std::vector<u64> pw(64, 0);
  u64 b = base;
  for (int k = 0; k < std::max(bits, 1); ++k) {
    pw[k] = ctx->hf.tab(b);
    b = ctx->hf.mul(b, b);
  }
  u64* d_pw = nullptr;
  auto fail = [&](const char* what) {
    if (d_pw) cudaFree(d_pw);
    cudaFree(*out);
    *out = nullptr;
    ctx->last_err = what;
    return NTT256_ERR_CUDA;
  };
  if (cudaMalloc((void**)&d_pw, pw.size() * 8) != cudaSuccess) return fail("cudaMalloc (power seeds)");
  if (cudaMemcpyAsync(d_pw, pw.data(), pw.size() * 8, cudaMemcpyHostToDevice, stream) != cudaSuccess)
    return fail("cudaMemcpy (power seeds)");
  if (ctx->hf.gl) launch_powers<true>(*out, d_pw, bits, count, ctx, ctx->hf.tab(start), stream);
  else launch_powers<false>(*out, d_pw, bits, count, ctx, ctx->hf.tab(start), stream);
  ++ctx->launches;
  if (cudaGetLastError() != cudaSuccess || cudaStreamSynchronize(stream) != cudaSuccess) return fail("twiddle_powers_kernel64");
  cudaFree(d_pw);
  return NTT256_OK;
}

// Kernel shape by call size: sixteen elements per thread in 2^12-element tiles for throughput; smaller tiles,
// then four elements per thread, for calls too small to fill 148 SMs with such CTAs.
void choose_shape(const ntt64_ctx* ctx, int log_n, int batch, int* tile_bits, int* group_bits, int* coalesce) {
  int T = std::min(ctx->tile_bits, log_n), gb = 4, g = ctx->coalesce_bits;
  if (ctx->auto_shape) {
    const long long E = (long long)batch << log_n;
    if (E <= (1ll << 15)) {
      gb = 2;
      T = std::min(8, log_n);
      g = 0;
    } else if (E <= (1ll << 19)) {
      T = std::min(10, log_n);
      g = std::min(g, 2);
    } else if (ctx->tile_bits == kTileBits64 && ctx->coalesce_bits == kCoalesce64) {
      // the plans the search finds under the default constraints are not equally good for every length
      // (profiles/r02_ntt64_shape_sweep.jsonl): 2^19 and 2^22 lose a pass with runs of 8 elements instead of 16,
      // batches of 2^16 (and 2^13) want 2^13-element tiles (two passes of eight stages instead of three)
      if (log_n == 19 || log_n == 22) g = 3;
      if ((log_n == 16 && batch >= 64) || log_n == 13) {
        T = 13;
        g = 2;
      }
    }
  }
  if (T < gb) gb = 2;
  g = std::min(g, T);
  *tile_bits = T;
  *group_bits = gb;
  *coalesce = g;
}

bool plan_for(int log_n, int T, int gb, int g, int max_log_n, Plan* plan, const ntt256::FourStepKey* fk = nullptr) {
  PlanRequest req;
  if (fk) {
    req = ntt256::fourstep_request(*fk, T, g, max_log_n);
  } else {
    req.log_n = log_n;
    req.tile_bits = T;
    req.coalesce_bits = g;
    req.max_log_n = max_log_n;
  }
  req.group_bits = gb;
  req.max_tile_bits = kMaxTile64;
  bool ok = make_plan(req, plan);
  // small tiles with wide coalescing leave too few free tile bits: relax the run length, then allow more
  // passes, then grow the tile
  for (int gg = req.coalesce_bits - 1; !ok && gg >= 0; --gg) {
    req.coalesce_bits = gg;
    ok = make_plan(req, plan, 12);
  }
  for (int t = req.tile_bits + 1; !ok && t <= kMaxTile64; ++t) {
    req.tile_bits = t;
    ok = make_plan(req, plan, 12);
  }
  return ok;
}

void free_plan_tables(ntt64_ctx* ctx) {
  for (auto& kv : ctx->plan_tables) ntt256::free_pass_tables(&kv.second);
  ctx->plan_tables.clear();
}

int get_plan(ntt64_ctx* ctx, int log_n, int batch, const Plan** out, const ntt64_ctx::PlanTables** tabs,
             const ntt256::FourStepKey* fk = nullptr) {
  PlanKey64 key;
  key.log_n = log_n;
  // a four-step piece works on the rank's 2^(log_n - log_p) local elements
  choose_shape(ctx, fk ? log_n - fk->log_p : log_n, batch, &key.tile_bits, &key.group_bits, &key.coalesce);
  if (fk) {
    key.kind = fk->kind;
    key.log_r = fk->log_r;
    key.log_p = fk->log_p;
    key.rank = fk->rank;
  }
  auto it = ctx->plans.find(key);
  if (it == ctx->plans.end()) {
    Plan plan;
    if (!plan_for(log_n, key.tile_bits, key.group_bits, key.coalesce, ctx->max_log_n, &plan, fk)) return NTT256_ERR_PLAN;
    ntt64_ctx::PlanTables pt;
    if (ntt256::build_pass_tables(plan, &pt) != cudaSuccess) {
      ctx->last_err = "index tables of a plan (cudaMalloc / cudaMemcpy)";
      return NTT256_ERR_CUDA;
    }
    ctx->plan_tables.emplace(key, pt);
    it = ctx->plans.emplace(key, plan).first;
  }
  *out = &it->second;
  *tabs = &ctx->plan_tables.find(key)->second;
  return NTT256_OK;
}

template <class K>
void launch_one64(K kernel, dim3 grid, unsigned threads, size_t smem, cudaStream_t stream, const u64* in, u64* out,
                  const PassArgs64& A) {
  // programmatic stream serialization: the pass becomes resident while its predecessor drains (the kernels wait
  // with griddepcontrol.wait before they touch data)
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(threads, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, kernel, in, out, A);
}
template <bool GL, int K16>
void launch_pass_k(int gb, dim3 grid, unsigned threads, size_t smem, cudaStream_t stream, const u64* in, u64* out,
                   const PassArgs64& A) {
  if (gb == 4 && threads <= 256) launch_one64(ntt64_pass16d_kernel<GL, K16>, grid, threads, smem, stream, in, out, A);
  else if (gb == 4) launch_one64(ntt64_pass16_kernel<GL, K16>, grid, threads, smem, stream, in, out, A);
  else launch_one64(ntt64_pass4_kernel<GL, K16>, grid, threads, smem, stream, in, out, A);
}
// k16: 13 / 3 select the Goldilocks instances whose constant twiddles are shifts (w_16 = 2^(12 k16)); 0 = products
void launch_pass(bool gl, int k16, int gb, dim3 grid, unsigned threads, size_t smem, cudaStream_t stream, const u64* in,
                 u64* out, const PassArgs64& A) {
  if (!gl) launch_pass_k<false, 0>(gb, grid, threads, smem, stream, in, out, A);
  else if (k16 == 13) launch_pass_k<true, 13>(gb, grid, threads, smem, stream, in, out, A);
  else if (k16 == 3) launch_pass_k<true, 3>(gb, grid, threads, smem, stream, in, out, A);
  else launch_pass_k<true, 0>(gb, grid, threads, smem, stream, in, out, A);
}

template <bool GL, int K16>
bool set_smem_limits_k() {
  const int big = 8 << kMaxTile64, small = 8 << 12;
  return cudaFuncSetAttribute(ntt64_pass16_kernel<GL, K16>, cudaFuncAttributeMaxDynamicSharedMemorySize, big) == cudaSuccess &&
         cudaFuncSetAttribute(ntt64_pass16d_kernel<GL, K16>, cudaFuncAttributeMaxDynamicSharedMemorySize, small) == cudaSuccess &&
         cudaFuncSetAttribute(ntt64_pass4_kernel<GL, K16>, cudaFuncAttributeMaxDynamicSharedMemorySize, small) == cudaSuccess;
}
bool set_smem_limits(bool gl) {
  if (!gl) return set_smem_limits_k<false, 0>();
  return set_smem_limits_k<true, 0>() && set_smem_limits_k<true, 13>() && set_smem_limits_k<true, 3>();
}

struct Fuse64 {
  const u64* in_scale = nullptr;
  const u64* out_scale = nullptr;
  bool out_has_ninv = false;  // out_scale already carries 1/n
};

// The passes of one plan over `batch` transforms starting at `p` (batch <= 65535); scale_log_n >= 0: multiply by
// 2^-scale_log_n on the way (folded into a round-twiddle table when the last pass has two rounds).
int run_plan64(ntt64_ctx* ctx, const Plan& plan, const ntt64_ctx::PlanTables& tabs, u64* p, unsigned long long batch_stride,
               int batch, bool inverse, int scale_log_n, cudaStream_t stream, const Fuse64* fuse, const ntt256::PeerSpec* peer) {
  PassArgs64 A;
  std::memset(&A, 0, sizeof(A));
  A.f = ctx->fp;
  A.om_hi = ctx->om_hi;
  A.om_lo = ctx->om_lo;
  A.loc = ctx->loc;
  A.hi_bits = ctx->hi_bits;
  A.lo_bits = ctx->lo_bits;
  A.loc_bits = ctx->loc_bits;
  A.inverse = inverse ? 1 : 0;
  std::memcpy(A.c16, ctx->c16[inverse ? 1 : 0], sizeof(A.c16));
  A.batch_stride = batch_stride;
  A.scaled_round = -1;
  const bool scale = scale_log_n >= 0;
  const u64 ninv_tab = scale ? ctx->hf.tab(ctx->hf.inv((1ull << scale_log_n) % ctx->hf.p)) : 0;
  A.scale = ninv_tab;
  const u64* loc_scaled = nullptr;
  if (scale && plan.passes.back().nrounds >= 2) {
    // 1/n rides on the round twiddles of the last pass's first round (all sixteen elements of that round are
    // multiplied, the one whose twiddle is 1 by 1/n itself)
    auto it = ctx->loc_scaled.find(scale_log_n);
    if (it == ctx->loc_scaled.end()) {
      u64* tab = nullptr;
      const unsigned long long count = 1ull << ctx->loc_bits;
      CTX_CUDA(cudaMalloc((void**)&tab, count * 8));
      if (ctx->hf.gl) table_scale_kernel64<true><<<(unsigned)((count + 255) / 256), 256, 0, stream>>>(tab, ctx->loc, count, ctx->fp, ninv_tab);
      else table_scale_kernel64<false><<<(unsigned)((count + 255) / 256), 256, 0, stream>>>(tab, ctx->loc, count, ctx->fp, ninv_tab);
      ++ctx->launches;
      CTX_CUDA(cudaGetLastError());
      CTX_CUDA(cudaStreamSynchronize(stream));  // one-off: later calls may use other streams
      it = ctx->loc_scaled.emplace(scale_log_n, tab).first;
    }
    loc_scaled = it->second;
  }
  const int GB = plan.group_bits;
  for (size_t i = 0; i < plan.passes.size(); ++i) {
    A.pd = plan.passes[i];
    A.scaled_round = -1;
    A.loc_scaled = nullptr;
    const bool last_pass = (i + 1 == plan.passes.size());
    if (scale && last_pass) {
      if (loc_scaled) {
        A.scaled_round = 0;
        A.loc_scaled = loc_scaled;
      } else {
        A.pd.scale_out = 1;
      }
    }
    A.thr = tabs.thr[i];
    A.blk = tabs.blk[i];
    A.in_scale = (fuse && i == 0) ? fuse->in_scale : nullptr;
    A.out_scale = (fuse && last_pass) ? fuse->out_scale : nullptr;
    A.peer_bits = 0;
    if (peer && last_pass) {
      for (int r = 0; r < (1 << peer->bits); ++r) A.peer[r] = (unsigned long long)peer->ptrs[r];
      A.peer_bits = peer->bits;
      A.peer_shift = peer->shift;
      A.peer_rank = peer->rank;
    }
    const int T = A.pd.T;
    const unsigned long long nblk = 1ull << A.pd.nblk_bits;
    if (nblk > 0x7fffffffull || batch > 65535) return NTT256_ERR_ARG;
    dim3 grid((unsigned)nblk, (unsigned)batch, 1);
    const unsigned threads = 1u << (T - GB);
    const size_t smem = (size_t)8 << T;
    launch_pass(ctx->hf.gl, ctx->use_shift_twiddles ? ctx->k16[inverse ? 1 : 0] : 0, GB, grid, threads, smem, stream, p, p, A);
    ++ctx->launches;
    CTX_CUDA(cudaGetLastError());
  }
  return NTT256_OK;
}

int transform(ntt64_ctx* ctx, void* d_data, int log_n, int batch, bool inverse, cudaStream_t stream,
              const Fuse64* fuse = nullptr) {
  if (!ctx || !d_data || log_n < 0 || log_n > ctx->max_log_n || batch < 1) return NTT256_ERR_ARG;
  CTX_CUDA(cudaSetDevice(ctx->device));
  const bool scale = inverse && !(fuse && fuse->out_has_ninv);
  const u64 ninv_tab = ctx->hf.tab(ctx->hf.inv((1ull << log_n) % ctx->hf.p));
  if (log_n < 3) {
    // the end-pass companions of such sizes are applied by their callers (coset_*), never here
    if (log_n == 0) {
      return NTT256_OK;
    }
    TinyArgs64 A;
    std::memset(&A, 0, sizeof(A));
    A.f = ctx->fp;
    A.log_n = log_n;
    A.w4 = ctx->w4[inverse ? 1 : 0];
    A.do_scale = scale ? 1 : 0;
    A.scale = ninv_tab;
    A.batch = (unsigned long long)batch;
    const int threads = 128;
    const unsigned blocks = (unsigned)((A.batch + threads - 1) / threads);
    if (ctx->hf.gl) ntt64_tiny_kernel<true><<<blocks, threads, 0, stream>>>((u64*)d_data, A);
    else ntt64_tiny_kernel<false><<<blocks, threads, 0, stream>>>((u64*)d_data, A);
    ++ctx->launches;
    CTX_CUDA(cudaGetLastError());
    return NTT256_OK;
  }
  const Plan* plan = nullptr;
  const ntt64_ctx::PlanTables* tabs = nullptr;
  int rc = get_plan(ctx, log_n, batch, &plan, &tabs);
  if (rc) return rc;
  const int kSlice = 65535;
  for (int b0 = 0; b0 < batch; b0 += kSlice) {
    const int nb = std::min(kSlice, batch - b0);
    rc = run_plan64(ctx, *plan, *tabs, (u64*)d_data + ((size_t)b0 << log_n), 1ull << log_n, nb, inverse, scale ? log_n : -1,
                    stream, fuse, nullptr);
    if (rc) return rc;
  }
  return NTT256_OK;
}

// One local step of the multi-GPU four-step transform (kinds of plan.h FourStepKey); peer_recv != null: the last
// pass stores straight into the ranks' receive buffers.
int fourstep64(ntt64_ctx* ctx, int kind, void* d_shard, int log_n, int log_r, int log_p, int rank, cudaStream_t stream,
               const void* const* peer_recv = nullptr) {
  const bool inverse = kind >= 3;
  const bool first_step = (kind == 1 || kind == 3);
  if (!ctx || !d_shard || log_n > ctx->max_log_n || !ntt256::fourstep_args_ok(log_n, log_r, log_p, rank)) return NTT256_ERR_ARG;
  CTX_CUDA(cudaSetDevice(ctx->device));
  const ntt256::FourStepKey fk{kind, log_n, log_r, log_p, first_step ? rank : 0};
  const Plan* plan = nullptr;
  const ntt64_ctx::PlanTables* tabs = nullptr;
  int rc = get_plan(ctx, log_n, 1, &plan, &tabs, &fk);
  if (rc) return rc;
  ntt256::PeerSpec ps;
  if (peer_recv && log_p > 0) {
    if (log_p > 3) return NTT256_ERR_ARG;
    ps.ptrs = peer_recv;
    ps.bits = log_p;
    ps.shift = log_n - 2 * log_p;  // local index bits below the destination-rank field
    ps.rank = rank;
  }
  return run_plan64(ctx, *plan, *tabs, (u64*)d_shard, 0, 1, inverse, kind == 4 ? log_n : -1, stream, nullptr,
                    ps.bits ? &ps : nullptr);
}

int ensure_stage(ntt64_ctx* ctx, size_t bytes) {
  if (bytes > ctx->stage_bytes) {
    if (ctx->d_stage) cudaFree(ctx->d_stage);
    ctx->d_stage = nullptr;
    ctx->stage_bytes = 0;
    CTX_CUDA(cudaMalloc(&ctx->d_stage, bytes));
    ctx->stage_bytes = bytes;
  }
  return NTT256_OK;
}

// Host buffers in, host buffers out.  One transform (or a small call): H2D, passes, D2H on the default stream.
// A batch: chunks of ~32 MiB through a three-slot ring on three streams, so uplink, kernels and downlink work on
// different chunks at the same time.
int transform_host(ntt64_ctx* ctx, const void* h_in, void* h_out, int log_n, int batch, bool inverse) {
  if (!ctx || !h_in || !h_out || log_n < 0 || log_n > ctx->max_log_n || batch < 1) return NTT256_ERR_ARG;
  CTX_CUDA(cudaSetDevice(ctx->device));
  const size_t one = ((size_t)1 << log_n) * 8;
  const size_t bytes = one * batch;
  const size_t kChunkBytes = (size_t)32 << 20;
  const int per_chunk = (int)std::max<size_t>(1, kChunkBytes / one);
  if (batch < 2 * per_chunk || batch < 4) {
    int rc = ensure_stage(ctx, bytes);
    if (rc) return rc;
    CTX_CUDA(cudaMemcpyAsync(ctx->d_stage, h_in, bytes, cudaMemcpyHostToDevice, 0));
    rc = transform(ctx, ctx->d_stage, log_n, batch, inverse, 0);
    if (rc) return rc;
    CTX_CUDA(cudaMemcpyAsync(h_out, ctx->d_stage, bytes, cudaMemcpyDeviceToHost, 0));
    CTX_CUDA(cudaStreamSynchronize(0));
    return NTT256_OK;
  }
  int rc = ensure_stage(ctx, 3 * one * per_chunk);
  if (rc) return rc;
  return ntt256::host_ring(ctx, h_in, h_out, one, batch, per_chunk,
                   [&](void* d, int nb, cudaStream_t s) { return transform(ctx, d, log_n, nb, inverse, s); });
}

void free_coset(ntt64_ctx::Coset& c) {
  if (c.direct) cudaFree(c.direct);
  if (c.hi) cudaFree(c.hi);
  if (c.lo) cudaFree(c.lo);
}

// shift^j tables (optionally times 1/n), cached per (shift, log_n, with_ninv); at most four are kept
int get_coset(ntt64_ctx* ctx, u64 shift, int log_n, bool with_ninv, const ntt64_ctx::Coset** out, cudaStream_t stream) {
  for (auto& c : ctx->cosets)
    if (c.shift == shift && c.log_n == log_n && c.with_ninv == (with_ninv ? 1 : 0)) {
      *out = &c;
      return NTT256_OK;
    }
  if (ctx->cosets.size() >= 4) {
    cudaDeviceSynchronize();
    free_coset(ctx->cosets.front());
    ctx->cosets.erase(ctx->cosets.begin());
  }
  ntt64_ctx::Coset c;
  std::memset(&c, 0, sizeof(c));
  c.shift = shift;
  c.log_n = log_n;
  c.with_ninv = with_ninv ? 1 : 0;
  const u64 start = with_ninv ? ctx->hf.inv((1ull << log_n) % ctx->hf.p) : 1;
  int rc;
  if (((size_t)8 << log_n) <= ctx->coset_direct_max) {
    rc = build_power_table(ctx, &c.direct, shift, log_n, start, stream);
    if (rc) return rc;
  } else {
    c.lo_bits = log_n / 2;
    rc = build_power_table(ctx, &c.lo, shift, c.lo_bits, start, stream);
    if (rc) return rc;
    rc = build_power_table(ctx, &c.hi, pow2k(ctx->hf, shift, c.lo_bits), log_n - c.lo_bits, 1, stream);
    if (rc) {
      cudaFree(c.lo);
      return rc;
    }
  }
  ctx->cosets.push_back(c);
  *out = &ctx->cosets.back();
  return NTT256_OK;
}

int coset_scale(ntt64_ctx* ctx, void* d_data, int log_n, int batch, const ntt64_ctx::Coset& c, cudaStream_t stream) {
  const unsigned long long total = (unsigned long long)batch << log_n;
  const int threads = 256;
  const unsigned blocks = (unsigned)((total + threads - 1) / threads);
  const u64* hi = c.direct ? c.direct : c.hi;
  // a direct table is its own "hi" table with no low part: lo_bits = 0 and lo[0] = 1 is what om_lo[0] holds
  const u64* lo = c.direct ? ctx->om_lo : c.lo;
  const int lo_bits = c.direct ? 0 : c.lo_bits;
  if (ctx->hf.gl) coset_scale_kernel64<true><<<blocks, threads, 0, stream>>>((u64*)d_data, total, log_n, lo_bits, hi, lo, ctx->fp);
  else coset_scale_kernel64<false><<<blocks, threads, 0, stream>>>((u64*)d_data, total, log_n, lo_bits, hi, lo, ctx->fp);
  ++ctx->launches;
  CTX_CUDA(cudaGetLastError());
  return NTT256_OK;
}

}  // namespace

extern "C" {

int ntt64_ctx_create(ntt64_ctx** out, int device, uint64_t p, uint64_t g, int max_log_n) {
  if (!out || max_log_n < 1 || max_log_n > 32) return NTT256_ERR_ARG;
  *out = nullptr;
  if (!(p & 1) || p < 3) return NTT256_ERR_FIELD;
  if (p != kGoldilocksP && (p >> 63)) return NTT256_ERR_FIELD;  // the Montgomery instance needs p < 2^63
  if (g == 0 || g >= p) return NTT256_ERR_FIELD;
  if (max_log_n < 64 && ((p - 1) & ((1ull << max_log_n) - 1ull))) return NTT256_ERR_FIELD;
  ntt64_ctx* ctx = new (std::nothrow) ntt64_ctx();
  if (!ctx) return NTT256_ERR_NOMEM;
  ctx->device = device;
  ctx->max_log_n = max_log_n;
  ctx->hf.init(p);
  ctx->fp.p = p;
  ctx->fp.np = ctx->hf.np;
  const HostField64& hf = ctx->hf;
  // w_N = g^((p-1)/N) must have exact order N
  ctx->wN = hf.pow(g, (p - 1) >> max_log_n);
  ctx->wN_inv = hf.inv(ctx->wN);
  if (pow2k(hf, ctx->wN, max_log_n - 1) != p - 1) {
    delete ctx;
    return NTT256_ERR_FIELD;
  }
  ctx->hi_bits = std::min(max_log_n, kHiBits64);
  ctx->lo_bits = max_log_n - ctx->hi_bits;
  ctx->loc_bits = std::min(max_log_n, kLocBits64);
  std::memset(ctx->c16, 0, sizeof(ctx->c16));
  // constants of the register rounds: powers of w_16 (of w_2^max_log_n when the context is smaller: only the
  // powers that exist are ever indexed, scaled accordingly)
  for (int dir = 0; dir < 2; ++dir) {
    const u64 w = dir ? ctx->wN_inv : ctx->wN;
    if (max_log_n >= 4) {
      const u64 w16 = pow2k(hf, w, max_log_n - 4);
      u64 a = w16;
      for (int k = 0; k < 7; ++k) {
        ctx->c16[dir][k] = hf.tab(a);
        a = hf.mul(a, w16);
      }
    } else {
      // w_16^e with e a multiple of 16 / 2^max_log_n
      const int step = 16 >> max_log_n;
      u64 a = w;
      for (int e = step; e <= 7; e += step) {
        ctx->c16[dir][e - 1] = hf.tab(a);
        a = hf.mul(a, w);
      }
    }
    ctx->w4[dir] = max_log_n >= 2 ? hf.tab(pow2k(hf, w, max_log_n - 2)) : 0;
    // Goldilocks: is w_16 = 2^(12 k) for one of the k the kernels are compiled for?
    if (hf.gl && max_log_n >= 4) {
      const u64 w16 = pow2k(hf, w, max_log_n - 4);
      for (int k : {13, 3})
        if (hf.pow(2, (u64)(12 * k)) == w16) ctx->k16[dir] = k;
    }
  }
  auto fail = [&](int code) {
    ntt64_ctx_destroy(ctx);
    return code;
  };
  if (cudaSetDevice(device) != cudaSuccess) return fail(NTT256_ERR_CUDA);
  if (!set_smem_limits(hf.gl)) return fail(NTT256_ERR_CUDA);
  int rc = build_power_table(ctx, &ctx->om_hi, pow2k(hf, ctx->wN, ctx->lo_bits), ctx->hi_bits);
  if (rc) return fail(rc);
  rc = build_power_table(ctx, &ctx->om_lo, ctx->wN, ctx->lo_bits);
  if (rc) return fail(rc);
  rc = build_power_table(ctx, &ctx->loc, pow2k(hf, ctx->wN, max_log_n - ctx->loc_bits), ctx->loc_bits);
  if (rc) return fail(rc);
  *out = ctx;
  return NTT256_OK;
}

void ntt64_ctx_destroy(ntt64_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  if (ctx->om_hi) cudaFree(ctx->om_hi);
  if (ctx->om_lo) cudaFree(ctx->om_lo);
  if (ctx->loc) cudaFree(ctx->loc);
  for (auto& kv : ctx->loc_scaled) cudaFree(kv.second);
  free_plan_tables(ctx);
  for (auto& c : ctx->cosets) free_coset(c);
  if (ctx->s_h2d) cudaStreamDestroy(ctx->s_h2d);
  if (ctx->s_d2h) cudaStreamDestroy(ctx->s_d2h);
  if (ctx->s_comp) cudaStreamDestroy(ctx->s_comp);
  for (int i = 0; i < 3; ++i) {
    if (ctx->ev_in[i]) cudaEventDestroy(ctx->ev_in[i]);
    if (ctx->ev_done[i]) cudaEventDestroy(ctx->ev_done[i]);
    if (ctx->ev_out[i]) cudaEventDestroy(ctx->ev_out[i]);
  }
  if (ctx->d_stage) cudaFree(ctx->d_stage);
  delete ctx;
}

int ntt64_ctx_set_tuning(ntt64_ctx* ctx, int tile_bits, int coalesce_bits) {
  if (!ctx || tile_bits < 3 || tile_bits > kMaxTile64 || coalesce_bits < 0 || coalesce_bits > 6 || coalesce_bits > tile_bits)
    return NTT256_ERR_ARG;
  ctx->tile_bits = tile_bits;
  ctx->coalesce_bits = coalesce_bits;
  ctx->auto_shape = false;
  cudaSetDevice(ctx->device);
  cudaDeviceSynchronize();
  ctx->plans.clear();  // loc_scaled tables depend on log_n only and stay valid
  free_plan_tables(ctx);
  return NTT256_OK;
}

int ntt64_ctx_set_coset_table_limit(ntt64_ctx* ctx, size_t max_bytes) {
  if (!ctx) return NTT256_ERR_ARG;
  cudaSetDevice(ctx->device);
  cudaDeviceSynchronize();
  for (auto& c : ctx->cosets) free_coset(c);
  ctx->cosets.clear();
  ctx->coset_direct_max = max_bytes;
  return NTT256_OK;
}

int ntt64_ctx_set_shift_twiddles(ntt64_ctx* ctx, int enable) {
  if (!ctx) return NTT256_ERR_ARG;
  ctx->use_shift_twiddles = enable != 0;
  return NTT256_OK;
}

int ntt64_forward(ntt64_ctx* ctx, void* d_data, int log_n, ntt256_stream stream) {
  return transform(ctx, d_data, log_n, 1, false, (cudaStream_t)stream);
}
int ntt64_inverse(ntt64_ctx* ctx, void* d_data, int log_n, ntt256_stream stream) {
  return transform(ctx, d_data, log_n, 1, true, (cudaStream_t)stream);
}
int ntt64_forward_batch(ntt64_ctx* ctx, void* d_data, int log_n, int batch, ntt256_stream stream) {
  return transform(ctx, d_data, log_n, batch, false, (cudaStream_t)stream);
}
int ntt64_inverse_batch(ntt64_ctx* ctx, void* d_data, int log_n, int batch, ntt256_stream stream) {
  return transform(ctx, d_data, log_n, batch, true, (cudaStream_t)stream);
}
int ntt64_forward_host(ntt64_ctx* ctx, const void* h_in, void* h_out, int log_n, int batch) {
  return transform_host(ctx, h_in, h_out, log_n, batch, false);
}
int ntt64_inverse_host(ntt64_ctx* ctx, const void* h_in, void* h_out, int log_n, int batch) {
  return transform_host(ctx, h_in, h_out, log_n, batch, true);
}

int ntt64_fourstep_step1(ntt64_ctx* ctx, void* d_shard, int log_n, int log_r, int log_p, int rank, ntt256_stream stream) {
  return fourstep64(ctx, 1, d_shard, log_n, log_r, log_p, rank, (cudaStream_t)stream);
}
int ntt64_fourstep_step1_peer(ntt64_ctx* ctx, void* d_shard, const void* const* peer_recv, int log_n, int log_r, int log_p,
                              int rank, ntt256_stream stream) {
  if (!peer_recv) return NTT256_ERR_ARG;
  return fourstep64(ctx, 1, d_shard, log_n, log_r, log_p, rank, (cudaStream_t)stream, peer_recv);
}
int ntt64_fourstep_step2(ntt64_ctx* ctx, void* d_shard, int log_n, int log_r, int log_p, int rank, ntt256_stream stream) {
  return fourstep64(ctx, 2, d_shard, log_n, log_r, log_p, rank, (cudaStream_t)stream);
}
int ntt64_fourstep_inverse_step1(ntt64_ctx* ctx, void* d_shard, int log_n, int log_r, int log_p, int rank, ntt256_stream stream) {
  return fourstep64(ctx, 3, d_shard, log_n, log_r, log_p, rank, (cudaStream_t)stream);
}
int ntt64_fourstep_inverse_step1_peer(ntt64_ctx* ctx, void* d_shard, const void* const* peer_recv, int log_n, int log_r,
                                      int log_p, int rank, ntt256_stream stream) {
  if (!peer_recv) return NTT256_ERR_ARG;
  return fourstep64(ctx, 3, d_shard, log_n, log_r, log_p, rank, (cudaStream_t)stream, peer_recv);
}
int ntt64_fourstep_inverse_step2(ntt64_ctx* ctx, void* d_shard, int log_n, int log_r, int log_p, int rank, ntt256_stream stream) {
  return fourstep64(ctx, 4, d_shard, log_n, log_r, log_p, rank, (cudaStream_t)stream);
}

int ntt64_coset_forward(ntt64_ctx* ctx, void* d_data, int log_n, int batch, uint64_t shift, ntt256_stream stream) {
  if (!ctx || !d_data || log_n < 0 || log_n > ctx->max_log_n || batch < 1 || shift == 0 || shift >= ctx->hf.p)
    return NTT256_ERR_ARG;
  CTX_CUDA(cudaSetDevice(ctx->device));
  if (log_n == 0) return NTT256_OK;
  const ntt64_ctx::Coset* c = nullptr;
  int rc = get_coset(ctx, shift, log_n, false, &c, (cudaStream_t)stream);
  if (rc) return rc;
  if (c->direct && log_n >= 3) {
    Fuse64 fz;
    fz.in_scale = c->direct;  // shift^j multiplies the elements as the first pass loads them
    return transform(ctx, d_data, log_n, batch, false, (cudaStream_t)stream, &fz);
  }
  rc = coset_scale(ctx, d_data, log_n, batch, *c, (cudaStream_t)stream);
  if (rc) return rc;
  return transform(ctx, d_data, log_n, batch, false, (cudaStream_t)stream);
}

int ntt64_coset_inverse(ntt64_ctx* ctx, void* d_data, int log_n, int batch, uint64_t shift, ntt256_stream stream) {
  if (!ctx || !d_data || log_n < 0 || log_n > ctx->max_log_n || batch < 1 || shift == 0 || shift >= ctx->hf.p)
    return NTT256_ERR_ARG;
  CTX_CUDA(cudaSetDevice(ctx->device));
  if (log_n == 0) return NTT256_OK;
  const u64 sinv = ctx->hf.inv(shift);
  const ntt64_ctx::Coset* c = nullptr;
  int rc = get_coset(ctx, sinv, log_n, true, &c, (cudaStream_t)stream);
  if (rc) return rc;
  Fuse64 fz;
  fz.out_has_ninv = true;  // the table is (1/n) * shift^-k: the transform itself does not scale
  if (c->direct && log_n >= 3) {
    fz.out_scale = c->direct;
    return transform(ctx, d_data, log_n, batch, true, (cudaStream_t)stream, &fz);
  }
  rc = transform(ctx, d_data, log_n, batch, true, (cudaStream_t)stream, &fz);
  if (rc) return rc;
  return coset_scale(ctx, d_data, log_n, batch, *c, (cudaStream_t)stream);
}

int ntt64_vec_mul(ntt64_ctx* ctx, void* d_out, const void* d_a, const void* d_b, unsigned long long n, ntt256_stream stream) {
  if (!ctx || !d_out || !d_a || !d_b) return NTT256_ERR_ARG;
  if (n == 0) return NTT256_OK;
  CTX_CUDA(cudaSetDevice(ctx->device));
  const int threads = 256;
  const unsigned blocks = (unsigned)((n + threads - 1) / threads);
  if (ctx->hf.gl)
    vec_mul_kernel64<true><<<blocks, threads, 0, (cudaStream_t)stream>>>((u64*)d_out, (const u64*)d_a, (const u64*)d_b, n, ctx->fp, ctx->hf.r2);
  else
    vec_mul_kernel64<false><<<blocks, threads, 0, (cudaStream_t)stream>>>((u64*)d_out, (const u64*)d_a, (const u64*)d_b, n, ctx->fp, ctx->hf.r2);
  ++ctx->launches;
  CTX_CUDA(cudaGetLastError());
  return NTT256_OK;
}

unsigned long long ntt64_launch_count(const ntt64_ctx* ctx) { return ctx ? ctx->launches : 0; }
const char* ntt64_last_cuda_error(const ntt64_ctx* ctx) { return ctx ? ctx->last_err.c_str() : ""; }

size_t ntt64_plan_describe_fourstep(int step, int log_n, int log_r, int log_p, int rank, int max_log_n, char* buf,
                                    size_t buf_len) {
  ntt64_ctx tmp;  // defaults only
  std::string s = "{\"error\": \"no plan\"}";
  if (step >= 1 && step <= 4 && ntt256::fourstep_args_ok(log_n, log_r, log_p, rank) && log_n <= max_log_n) {
    const bool first = (step == 1 || step == 3);
    const ntt256::FourStepKey fk{step, log_n, log_r, log_p, first ? rank : 0};
    int T, gb, g;
    choose_shape(&tmp, log_n - log_p, 1, &T, &gb, &g);
    Plan plan;
    if (plan_for(log_n, T, gb, g, max_log_n, &plan, &fk)) s = plan.describe();
  }
  if (buf && buf_len) {
    size_t m = std::min(buf_len - 1, s.size());
    std::memcpy(buf, s.data(), m);
    buf[m] = 0;
  }
  return s.size() + 1;
}

size_t ntt64_plan_describe(int log_n, int batch, int max_log_n, char* buf, size_t buf_len) {
  ntt64_ctx tmp;  // defaults only: tile / coalescing / automatic shapes
  int T, gb, g;
  std::string s = "{\"error\": \"no plan\"}";
  if (log_n >= 3 && log_n <= max_log_n && batch >= 1) {
    choose_shape(&tmp, log_n, batch, &T, &gb, &g);
    Plan plan;
    if (plan_for(log_n, T, gb, g, max_log_n, &plan)) s = plan.describe();
  }
  if (buf && buf_len) {
    size_t m = std::min(buf_len - 1, s.size());
    std::memcpy(buf, s.data(), m);
    buf[m] = 0;
  }
  return s.size() + 1;
}

}  // extern "C"
