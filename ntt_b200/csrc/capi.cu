// C ABI of libntt256.so (include/ntt256.h): context, device twiddle tables, plan cache and the
// launch sequences.  Host-side counterpart of the reference drivers NTT_GPU_Naive / NTT_GZKP
// (/root/reference/src/big-num.cu:106-170, 260-353) -- minus their per-call table build,
// bit-reversal list and printing.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>
#include <map>
#include <algorithm>
#include <new>
#include <tuple>
#include <string>
#include <vector>

#include "../../include/ntt256.h"
#include "hostfield.h"
#include "aux_kernels.cuh"
#include "pass_launch.h"
#include "host_common.h"
#include "pass_tables.h"
#include "plan.h"

using namespace ntt256;

namespace {

constexpr int kHiBitsMax = 18;       // w^(i<<lo) table: 2^18 entries = 8 MiB, stays in L2
constexpr int kHiBitsMaxLarge = 20;  // contexts for >= 2^26 points: 32 MiB, one two-factor pass less (2^28: -6 %)
constexpr int kLocBitsMax = 12;

struct PlanKey {
  int kind;  // 0 = whole transform, 1 = four-step step 1, 2 = four-step step 2
  int log_n, log_r, log_p, rank;
  int tile_bits = 0;   // 0 = the context's tile_bits
  int group_bits = 3;  // 3 = eight elements per thread; 2 = four (small, latency-bound transforms)
  int coalesce = -1;   // -1 = the context's coalesce_bits
  bool operator<(const PlanKey& o) const {
    return std::tie(kind, log_n, log_r, log_p, rank, tile_bits, group_bits, coalesce) <
           std::tie(o.kind, o.log_n, o.log_r, o.log_p, o.rank, o.tile_bits, o.group_bits, o.coalesce);
  }
};

}  // namespace

struct ntt256_ctx {
  int device = 0;
  int max_log_n = 0;
  HostField hf;
  FieldParams fp;
  U256 g_mont;       // generator, Montgomery form
  U256 wN, wN_inv;   // w_{2^max_log_n} and its inverse, Montgomery form
  uint4* om_hi = nullptr;
  uint4* om_lo = nullptr;
  uint4* loc = nullptr;
  int hi_bits = 0, lo_bits = 0, loc_bits = 0;
  u32 c8[2][3][8];   // [direction][k-1] = w_8^(+-k), Montgomery form
  u32 c8p[2][3][8];  // the same constants plain, and their Shoup companions (shoup contexts)
  u32 c8q[2][3][8];
  bool shoup = false;  // p < 2^254: twiddle tables hold (w, wq) pairs and the passes multiply with shoup_mul
  u32 w4[2][8];
  int tile_bits = 11;
  int coalesce_bits = 2;
  U256 ninv[64];                          // Montgomery form of 2^-k, filled on first use
  bool ninv_ok[64] = {};
  bool dense_pass4 = true;                // four elements per thread: three CTAs of <= 256 threads per SM
  int force_group_bits = 0;               // 0: automatic; 1..3: elements per thread = 2^k for every plan
  long long small_limit = 1ll << 17;      // batch*n at or below this: four (half of it: two) elements per thread
  bool auto_tile = true;                  // shrink the tile for small problems so the grid fills the chip
  bool zero_copy_host = true;             // *_host entry points: first/last pass access pinned host memory directly
  // coset shift / pointwise product inside the first / last pass: 1 always, 0 never, -1 (default) for calls of
  // at most 2^17 elements.  The companions cost one more product per element wherever they run; on this
  // multiplier-bound path the separate element-wise kernel (full occupancy, HBM at 80 %) is as fast as the
  // fused form from ~2^18 elements on, and the fused form only saves a launch (profiles/r02_companions.jsonl).
  int fuse_companions = -1;
  // programmatic dependent launch between the pass kernels of a stream (a pass becomes resident, and computes its
  // indices, while its predecessor drains its last wave)
  bool use_pdl = true;
  bool const_field = false;      // the modulus is BN254 Fr: the throughput kernel with compile-time field constants applies
  bool use_const_field = true;   // testing knob
  std::map<int, uint4*> loc_scaled;       // log_n -> (1/n) * loc[] : folds the inverse scaling into a round twiddle
  // streamed pass-twiddle arrays, keyed by (plan id, pass index, direction)
  struct TwStream {
    uint4* tab = nullptr;
    size_t bytes = 0;
    unsigned long long last_use = 0;  // value of tw_clock at the last launch that read it
  };
  std::map<std::tuple<unsigned long long, int, int>, TwStream> tw_streams;
  size_t tw_stream_limit = (size_t)1 << 30;   // largest array built automatically (bytes); 0 = never;
                                              // set to 1/10 of the device's memory at creation
  size_t tw_stream_budget = (size_t)2 << 30;  // all cached arrays together; least recently used ones are dropped
                                              // to make room (set to 1/8 of the device's memory at creation)
  size_t tw_stream_bytes = 0;
  unsigned long long tw_clock = 0;
  unsigned long long next_plan_id = 1;
  struct CosetTables {
    U256 shift;
    int log_n, lo_bits;
    uint4 *hi, *lo;
    uint4* direct;  // s^j for every j < 2^log_n in the context's twiddle format (32 or 64 B each), or null
    bool direct_shoup;
  };
  std::vector<CosetTables> cosets;        // small cache of shift^j tables (forward and inverse shifts)
  cudaStream_t s_h2d = nullptr, s_d2h = nullptr, s_comp = nullptr;   // pipelined host batches
  cudaEvent_t ev_in[3] = {nullptr, nullptr, nullptr}, ev_done[3] = {nullptr, nullptr, nullptr},
              ev_out[3] = {nullptr, nullptr, nullptr};
  std::map<PlanKey, Plan> plans;
  std::map<unsigned long long, PassTables> plan_tables;  // by Plan::id: the index tables of its passes (pass_tables.h)
  void* d_stage = nullptr;
  size_t stage_bytes = 0;
  // rank synchronisation folded into the four-step passes (ntt256_ctx_set_peer_sync)
  unsigned long long sync_flags[8] = {};
  int sync_nranks = 0, sync_rank = 0;
  unsigned long long sync_epoch = 0;
  unsigned long long sync_timeout_ns = 60ull * 1000000000ull;
  unsigned int* sync_counter = nullptr;
  unsigned long long launches = 0;
  std::string last_err;
};

namespace {

#define CTX_CUDA(call)                                   \
  do {                                                   \
    cudaError_t e_ = (call);                             \
    if (e_ != cudaSuccess) {                             \
      ctx->last_err = std::string(#call) + ": " + cudaGetErrorString(e_); \
      return NTT256_ERR_CUDA;                            \
    }                                                    \
  } while (0)

U256 root_of_unity(const HostField& hf, const U256& g_mont, int log_n) {
  U256 pm1, e;
  u256_sub(pm1, hf.p, u256_small(1));
  e = u256_shr(pm1, log_n);  // (p-1)/n, exact: checked at context creation
  return hf.pow(g_mont, e);
}

U256 pow2k(const HostField& hf, U256 b, int k) {  // b^(2^k)
  for (int i = 0; i < k; ++i) b = hf.mul(b, b);
  return b;
}

void to_u32(const U256& a, u32 out[8]) { u256_to_limbs(a, out); }

// shoup_entries: 64-byte (w, wq) entries for the pass kernels of a shoup context; else 32-byte Montgomery form.
int build_power_table(ntt256_ctx* ctx, uint4** out, const U256& base_mont, int bits, bool shoup_entries = false) {
  const unsigned long long count = 1ull << bits;
  CTX_CUDA(cudaMalloc((void**)out, count * (shoup_entries ? 64 : 32)));
  std::vector<u32> pw(8 * 64, 0);
  U256 b = base_mont;
  for (int k = 0; k < std::max(bits, 1); ++k) {
    to_u32(b, &pw[8 * k]);
    b = ctx->hf.mul(b, b);
  }
  uint4* d_pw = nullptr;
  auto fail = [&](const char* what) {
    if (d_pw) cudaFree(d_pw);
    cudaFree(*out);
    *out = nullptr;
    ctx->last_err = what;
    return NTT256_ERR_CUDA;
  };
  if (cudaMalloc((void**)&d_pw, pw.size() * 4) != cudaSuccess) return fail("cudaMalloc (power seeds)");
  if (cudaMemcpy(d_pw, pw.data(), pw.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess) return fail("cudaMemcpy (power seeds)");
  Fe one, pinv;
  to_u32(ctx->hf.one, one.v);
  to_u32(ctx->hf.pinv, pinv.v);
  const int threads = 128;
  const unsigned blocks = (unsigned)((count + threads - 1) / threads);
  twiddle_powers_kernel<<<blocks, threads>>>(*out, d_pw, bits, 0ull, count, ctx->fp, one, shoup_entries ? 1 : 0, pinv);
  ++ctx->launches;
  if (cudaGetLastError() != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) return fail("twiddle_powers_kernel");
  cudaFree(d_pw);
  return NTT256_OK;
}

void drop_tw_streams(ntt256_ctx* ctx) {
  for (auto& kv : ctx->tw_streams) cudaFree(kv.second.tab);
  ctx->tw_streams.clear();
  ctx->tw_stream_bytes = 0;
}

// Make room for a new streamed twiddle array of `need` bytes: least recently used arrays go first.  cudaFree
// waits for the device, so an array still being read by an earlier launch is not pulled from under it.
void evict_tw_streams(ntt256_ctx* ctx, size_t need) {
  while (!ctx->tw_streams.empty() && ctx->tw_stream_bytes + need > ctx->tw_stream_budget) {
    auto victim = ctx->tw_streams.begin();
    for (auto it = ctx->tw_streams.begin(); it != ctx->tw_streams.end(); ++it)
      if (it->second.last_use < victim->second.last_use) victim = it;
    cudaFree(victim->second.tab);
    ctx->tw_stream_bytes -= victim->second.bytes;
    ctx->tw_streams.erase(victim);
  }
}

void free_twiddle_tables(ntt256_ctx* ctx) {
  if (ctx->om_hi) cudaFree(ctx->om_hi);
  if (ctx->om_lo) cudaFree(ctx->om_lo);
  if (ctx->loc) cudaFree(ctx->loc);
  ctx->om_hi = ctx->om_lo = ctx->loc = nullptr;
  for (auto& kv : ctx->loc_scaled) cudaFree(kv.second);
  ctx->loc_scaled.clear();
  drop_tw_streams(ctx);
}

// The three w_N power tables of the pass kernels, in the context's twiddle representation.
int build_twiddle_tables(ntt256_ctx* ctx) {
  int rc = build_power_table(ctx, &ctx->om_hi, pow2k(ctx->hf, ctx->wN, ctx->lo_bits), ctx->hi_bits, ctx->shoup);
  if (rc) return rc;
  rc = build_power_table(ctx, &ctx->om_lo, ctx->wN, ctx->lo_bits, ctx->shoup);
  if (rc) return rc;
  return build_power_table(ctx, &ctx->loc, pow2k(ctx->hf, ctx->wN, ctx->max_log_n - ctx->loc_bits), ctx->loc_bits,
                           ctx->shoup);
}

PlanRequest build_request(const PlanKey& key, int tile_bits, int coalesce_bits, int max_log_n) {
  return fourstep_request(FourStepKey{key.kind, key.log_n, key.log_r, key.log_p, key.rank}, tile_bits, coalesce_bits, max_log_n);
}

int get_plan(ntt256_ctx* ctx, const PlanKey& key, const Plan** out) {
  auto it = ctx->plans.find(key);
  if (it == ctx->plans.end()) {
    PlanRequest req = build_request(key, key.tile_bits ? key.tile_bits : ctx->tile_bits,
                                    key.coalesce >= 0 ? key.coalesce : ctx->coalesce_bits, ctx->max_log_n);
    req.group_bits = key.group_bits;
    Plan plan;
    bool ok = make_plan(req, &plan);
    // very small tiles with wide coalescing leave too few free tile bits: relax the run length
    // first, then allow more passes, then grow the tile
    for (int g = req.coalesce_bits - 1; !ok && g >= 0; --g) {
      req.coalesce_bits = g;
      ok = make_plan(req, &plan, 12);
    }
    const int max_tile = (key.group_bits == 3) ? kMaxTileBits : 11;  // 1024 / 512 threads per CTA at most
    if (req.tile_bits > max_tile) return NTT256_ERR_PLAN;
    for (int t = req.tile_bits + 1; !ok && t <= max_tile; ++t) {
      req.tile_bits = t;
      ok = make_plan(req, &plan, 12);
    }
    if (!ok) return NTT256_ERR_PLAN;
    plan.id = ctx->next_plan_id++;
    PassTables pt;
    if (cudaSetDevice(ctx->device) != cudaSuccess || build_pass_tables(plan, &pt) != cudaSuccess) {
      ctx->last_err = "index tables of a plan (cudaMalloc / cudaMemcpy)";
      return NTT256_ERR_CUDA;
    }
    ctx->plan_tables.emplace(plan.id, pt);
    it = ctx->plans.emplace(key, plan).first;
  }
  *out = &it->second;
  return NTT256_OK;
}

// The pass kernel for 2^group_bits elements per thread; FZ instances only for launches that carry companions.
// const_field: the context's modulus is the one ntt_pass_bn254_kernel is compiled for
void launch_pass(bool shoup, int group_bits, bool dense, dim3 grid, unsigned threads, size_t smem, cudaStream_t stream,
                 const uint4* in, uint4* out, const PassArgs& A, bool pdl, bool const_field = false) {
  const bool fz = A.in_scale || A.out_scale || A.out_mul;
  if (shoup && const_field && !fz && group_bits == 3) {
    launch_pass_bn254(grid, threads, smem, stream, in, out, A, pdl);
    return;
  }
  if (shoup) {
    if (fz) launch_pass_inst<true, true>(group_bits, dense, grid, threads, smem, stream, in, out, A, pdl);
    else launch_pass_inst<true, false>(group_bits, dense, grid, threads, smem, stream, in, out, A, pdl);
  } else {
    if (fz) launch_pass_inst<false, true>(group_bits, dense, grid, threads, smem, stream, in, out, A, pdl);
    else launch_pass_inst<false, false>(group_bits, dense, grid, threads, smem, stream, in, out, A, pdl);
  }
}

// (2^k)^-1 mod p in Montgomery form (a host inversion is ~10 us: computed once per k)
const U256& inv_pow2(ntt256_ctx* ctx, int k) {
  if (!ctx->ninv_ok[k]) {
    ctx->ninv[k] = ctx->hf.inv(ctx->hf.to_mont(u256_small(1ull << k)));
    ctx->ninv_ok[k] = true;
  }
  return ctx->ninv[k];
}

// sync: 0 none; 1 the first pass waits for every rank's epoch; 2 the last pass publishes this rank's epoch
int run_plan(ntt256_ctx* ctx, const Plan& plan, void* d_data, unsigned long long batch_stride, int batch, bool inverse,
             int scale_log_n, cudaStream_t stream, const PeerSpec* peer = nullptr, const IoSpec* io = nullptr,
             int sync = 0, const FuseSpec* fuse = nullptr) {
  PassArgs A;
  std::memset(&A, 0, sizeof(A));
  A.f = ctx->fp;
  A.om_hi = ctx->om_hi;
  A.om_lo = ctx->om_lo;
  A.loc = ctx->loc;
  A.hi_bits = ctx->hi_bits;
  A.lo_bits = ctx->lo_bits;
  A.loc_bits = ctx->loc_bits;
  A.inverse = inverse ? 1 : 0;
  const bool sh = ctx->shoup;
  for (int k = 0; k < 3; ++k) {  // by-value constants as table entries: w, then its Shoup companion
    std::memcpy(A.c8[k], sh ? ctx->c8p[inverse ? 1 : 0][k] : ctx->c8[inverse ? 1 : 0][k], 32);
    if (sh) std::memcpy(A.c8[k] + 8, ctx->c8q[inverse ? 1 : 0][k], 32);
  }
  if (sh) {
    to_u32(ctx->hf.r2, A.r2);
    to_u32(ctx->hf.pinv, A.pinv);
  }
  A.batch_stride = batch_stride;
  A.scaled_round = -1;
  const uint4* loc_scaled = nullptr;
  if (scale_log_n >= 0) {
    const U256& ninv = inv_pow2(ctx, scale_log_n);
    if (sh) {
      const U256 plain = ctx->hf.from_mont(ninv);
      to_u32(plain, A.scale);
      to_u32(ctx->hf.shoup_companion(plain), A.scale + 8);
    } else {
      to_u32(ninv, A.scale);
    }
    if (plan.passes.back().nrounds >= 2) {
      // fold 1/n into the twiddle table of the last pass's first round (all 8 elements of that
      // round are multiplied, the trivial one by (1/n)*R instead of being skipped)
      auto it = ctx->loc_scaled.find(scale_log_n);
      if (it == ctx->loc_scaled.end()) {
        uint4* tab = nullptr;
        const unsigned long long count = 1ull << ctx->loc_bits;
        CTX_CUDA(cudaMalloc((void**)&tab, count * (sh ? 64 : 32)));
        Fe c, r2, pinv;
        to_u32(ninv, c.v);
        to_u32(ctx->hf.r2, r2.v);
        to_u32(ctx->hf.pinv, pinv.v);
        table_scale_kernel<<<(unsigned)((count + 127) / 128), 128, 0, stream>>>(tab, ctx->loc, count, ctx->fp, c,
                                                                               sh ? 1 : 0, r2, pinv);
        ++ctx->launches;
        CTX_CUDA(cudaGetLastError());
        CTX_CUDA(cudaStreamSynchronize(stream));  // one-off: later calls may use other streams
        it = ctx->loc_scaled.emplace(scale_log_n, tab).first;
      }
      loc_scaled = it->second;
    }
  }
  const PassTables& tabs = ctx->plan_tables.at(plan.id);
  for (size_t i = 0; i < plan.passes.size(); ++i) {
    A.pd = plan.passes[i];
    A.thr = tabs.thr[i];
    A.blk = tabs.blk[i];
    A.scaled_round = -1;
    A.loc_scaled = nullptr;
    if (scale_log_n >= 0 && i + 1 == plan.passes.size()) {
      if (loc_scaled) {
        A.scaled_round = 0;
        A.loc_scaled = loc_scaled;
      } else {
        A.pd.scale_out = 1;
      }
    }
    // streamed twiddles for a pass that would otherwise need the two-factor product
    A.tw_stream = nullptr;
    A.tw_gen_out = nullptr;
    if (A.pd.has_tw && A.pd.tw_lshift < ctx->lo_bits && ctx->tw_stream_limit > 0) {
      const size_t tbytes = (size_t)(sh ? 64 : 32) << (A.pd.T + A.pd.nblk_bits);
      if (tbytes <= ctx->tw_stream_limit && tbytes <= ctx->tw_stream_budget) {
        auto tkey = std::make_tuple(plan.id, (int)i, inverse ? 1 : 0);
        auto ts = ctx->tw_streams.find(tkey);
        if (ts == ctx->tw_streams.end()) {
          uint4* tab = nullptr;
          evict_tw_streams(ctx, tbytes);
          if (cudaMalloc((void**)&tab, tbytes) == cudaSuccess) {
            PassArgs G = A;
            G.tw_gen_out = tab;
            G.peer_bits = 0;
            to_u32(sh ? u256_small(1) : ctx->hf.one, G.one);
            const int Tg = G.pd.T;
            dim3 ggrid((unsigned)(1ull << G.pd.nblk_bits), 1, 1);
            launch_pass(sh, plan.group_bits, ctx->dense_pass4, ggrid, 1u << (Tg - plan.group_bits), (size_t)32 << Tg, stream,
                        (const uint4*)d_data, (uint4*)d_data, G, false, ctx->const_field && ctx->use_const_field);
            ++ctx->launches;
            CTX_CUDA(cudaGetLastError());
            // one-off: the array may be read from other streams by later calls
            CTX_CUDA(cudaStreamSynchronize(stream));
            ntt256_ctx::TwStream rec;
            rec.tab = tab;
            rec.bytes = tbytes;
            ts = ctx->tw_streams.emplace(tkey, rec).first;
            ctx->tw_stream_bytes += tbytes;
          } else {
            cudaGetLastError();  // not enough memory: keep the two-lookup path
          }
        }
        if (ts != ctx->tw_streams.end()) {
          ts->second.last_use = ++ctx->tw_clock;
          A.tw_stream = ts->second.tab;
        }
      }
    }
    if (peer && i + 1 == plan.passes.size()) {
      for (int r = 0; r < (1 << peer->bits); ++r) A.peer[r] = (unsigned long long)peer->ptrs[r];
      A.peer_bits = peer->bits;
      A.peer_shift = peer->shift;
      A.peer_rank = peer->rank;
    }
    A.in_scale = (fuse && i == 0) ? fuse->in_scale : nullptr;
    A.out_scale = (fuse && i + 1 == plan.passes.size()) ? fuse->out_scale : nullptr;
    A.out_mul = (fuse && i + 1 == plan.passes.size()) ? fuse->out_mul : nullptr;
    A.sync_mode = 0;
    if (sync && ctx->sync_nranks > 1) {
      if ((sync & 1) && i == 0) A.sync_mode |= 1;
      if ((sync & 2) && i + 1 == plan.passes.size()) A.sync_mode |= 2;
      if (A.sync_mode) {
        std::memcpy(A.sync_flags, ctx->sync_flags, sizeof(A.sync_flags));
        A.sync_epoch = ctx->sync_epoch;
        A.sync_timeout_ns = ctx->sync_timeout_ns;
        A.sync_counter = ctx->sync_counter;
        A.sync_nranks = ctx->sync_nranks;
        A.sync_rank = ctx->sync_rank;
      }
    }
    const int T = A.pd.T;
    const int GB = plan.group_bits;
    const unsigned threads = 1u << (T - GB);
    const size_t smem = (size_t)32 << T;
    const unsigned long long nblk = 1ull << A.pd.nblk_bits;
    if (nblk > 0x7fffffffull || batch > 65535) return NTT256_ERR_ARG;
    dim3 grid((unsigned)nblk, (unsigned)batch, 1);
    // first pass may read from `src_first`, last pass may write to `dst_last` (host-mapped buffers)
    const uint4* pin = (i == 0 && io && io->src_first) ? (const uint4*)io->src_first : (const uint4*)d_data;
    uint4* pout = (i + 1 == plan.passes.size() && io && io->dst_last) ? (uint4*)io->dst_last : (uint4*)d_data;
    launch_pass(sh, GB, ctx->dense_pass4, grid, threads, smem, stream, pin, pout, A, ctx->use_pdl,
                ctx->const_field && ctx->use_const_field);
    ++ctx->launches;
    CTX_CUDA(cudaGetLastError());
  }
  return NTT256_OK;
}

int run_tiny(ntt256_ctx* ctx, void* d_data, int log_n, unsigned long long batch, bool inverse, cudaStream_t stream) {
  if (log_n == 0) return NTT256_OK;
  TinyArgs A;
  std::memset(&A, 0, sizeof(A));
  A.f = ctx->fp;
  A.log_n = log_n;
  std::memcpy(A.w4, ctx->w4[inverse ? 1 : 0], sizeof(A.w4));
  A.do_scale = inverse ? 1 : 0;
  if (inverse) {
    to_u32(inv_pow2(ctx, log_n), A.scale);
  }
  A.batch = batch;
  const int threads = 128;
  ntt_tiny_kernel<<<(unsigned)((batch + threads - 1) / threads), threads, 0, stream>>>((uint4*)d_data, A);
  ++ctx->launches;
  CTX_CUDA(cudaGetLastError());
  return NTT256_OK;
}

int transform(ntt256_ctx* ctx, void* d_data, int log_n, int batch, bool inverse, cudaStream_t stream,
              const IoSpec* io = nullptr, const FuseSpec* fuse = nullptr) {
  if (!ctx || !d_data || log_n < 0 || log_n > ctx->max_log_n || batch < 1) return NTT256_ERR_ARG;
  CTX_CUDA(cudaSetDevice(ctx->device));
  if (log_n < 3) return run_tiny(ctx, d_data, log_n, (unsigned long long)batch, inverse, stream);
  // Kernel and tile choice by the number of elements E in the call (profiles/r01_small_sweep.md):
  // large calls are throughput-bound and want eight elements per thread and 2^11-element tiles;
  // below ~2^21 elements smaller tiles balance the 148 SMs better; below ~2^17 the call is
  // latency-bound (few warps per SM, each a serial chain of 256-bit multiplies), so the work is cut
  // finer: four, then two elements per thread in 2^9 / 2^8-element tiles, no coalescing constraint
  // (the data lives in L2).
  const Plan* plan = nullptr;
  PlanKey key{0, log_n, 0, 0, 0};
  if (ctx->auto_tile) {
    const long long E = (long long)batch << log_n;
    if (E <= ctx->small_limit / 2) {
      key.group_bits = 1;
      key.tile_bits = std::min(8, log_n);
      key.coalesce = 0;
    } else if (E <= ctx->small_limit) {
      key.group_bits = 2;
      key.tile_bits = std::min(9, log_n);
      key.coalesce = 0;
    } else if (E > (1ll << 19) && E <= (1ll << 20) && log_n >= 10 && ctx->dense_pass4) {
      // calls of about 2^20 elements: four elements per thread in 2^10-element tiles on the three-CTAs-per-SM instance
      // (24 warps per SM): 2-5 % ahead of eight per thread here, level with it below (profiles/r02_shapes_mid.jsonl)
      key.group_bits = 2;
      key.tile_bits = 10;
    } else if (E <= (1ll << 20)) {
      key.tile_bits = std::min(std::min(10, ctx->tile_bits), log_n);
    } else if (log_n < ctx->tile_bits) {
      key.tile_bits = log_n;
    }
    if (key.tile_bits == ctx->tile_bits && key.group_bits == 3) key.tile_bits = 0;
  }
  if (ctx->force_group_bits) key.group_bits = ctx->force_group_bits;
  int rc = get_plan(ctx, key, &plan);
  if (rc) return rc;
  // batches larger than one grid dimension are issued in slices
  const int kSlice = 65535;
  for (int b0 = 0; b0 < batch; b0 += kSlice) {
    const int nb = std::min(kSlice, batch - b0);
    const size_t off = ((size_t)b0 << log_n) * 32;
    char* p = (char*)d_data + off;
    // out-of-place ends (zero-copy host buffers) follow the slice: same offset in every buffer
    IoSpec io_slice;
    if (io) {
      io_slice.src_first = io->src_first ? (const char*)io->src_first + off : nullptr;
      io_slice.dst_last = io->dst_last ? (char*)io->dst_last + off : nullptr;
    }
    FuseSpec fuse_slice;
    if (fuse) {
      fuse_slice = *fuse;
      if (fuse->out_mul) fuse_slice.out_mul = fuse->out_mul + 2 * ((size_t)b0 << log_n);  // per call, not per transform
    }
    rc = run_plan(ctx, *plan, p, 1ull << log_n, nb, inverse, inverse ? log_n : -1, stream, nullptr,
                  io ? &io_slice : nullptr, 0, fuse ? &fuse_slice : nullptr);
    if (rc) return rc;
  }
  return NTT256_OK;
}

int ensure_stage(ntt256_ctx* ctx, size_t bytes) {
  if (bytes > ctx->stage_bytes) {
    if (ctx->d_stage) cudaFree(ctx->d_stage);
    ctx->d_stage = nullptr;
    ctx->stage_bytes = 0;
    CTX_CUDA(cudaMalloc(&ctx->d_stage, bytes));
    ctx->stage_bytes = bytes;
  }
  return NTT256_OK;
}

// Host buffers in, host buffers out.  A single transform is H2D -> passes -> D2H.  A batch is
// cut into chunks that flow through a 3-slot ring on three streams, so the PCIe uplink, the
// kernels and the downlink work on different chunks at the same time (pinned host memory is
// needed for the copies to be asynchronous; pageable memory still gives correct results).
int transform_host(ntt256_ctx* ctx, const void* h_in, void* h_out, int log_n, int batch, bool inverse) {
  if (!ctx || !h_in || !h_out || log_n < 0 || log_n > ctx->max_log_n || batch < 1) return NTT256_ERR_ARG;
  CTX_CUDA(cudaSetDevice(ctx->device));
  const size_t one = ((size_t)1 << log_n) * 32;
  const size_t bytes = one * batch;
  const size_t kChunkBytes = (size_t)32 << 20;
  int per_chunk = (int)std::max<size_t>(1, kChunkBytes / one);
  if (batch < 2 * per_chunk || batch < 4) {
    int rc = ensure_stage(ctx, bytes);
    if (rc) return rc;
    // Pinned (mapped) host buffers: the first pass loads straight from host memory and the last
    // pass stores straight to it, so the PCIe transfers overlap those passes.
    if (log_n >= 3 && ctx->zero_copy_host) {
      cudaPointerAttributes ai, ao;
      void *din = nullptr, *dout = nullptr;
      bool ok = cudaPointerGetAttributes(&ai, h_in) == cudaSuccess && ai.type == cudaMemoryTypeHost &&
                cudaPointerGetAttributes(&ao, h_out) == cudaSuccess && ao.type == cudaMemoryTypeHost &&
                cudaHostGetDevicePointer(&din, (void*)h_in, 0) == cudaSuccess &&
                cudaHostGetDevicePointer(&dout, h_out, 0) == cudaSuccess;
      cudaGetLastError();
      if (ok) {
        IoSpec io;
        io.src_first = din;
        io.dst_last = dout;
        rc = transform(ctx, ctx->d_stage, log_n, batch, inverse, 0, &io);
        if (rc) return rc;
        CTX_CUDA(cudaStreamSynchronize(0));
        return NTT256_OK;
      }
    }
    CTX_CUDA(cudaMemcpyAsync(ctx->d_stage, h_in, bytes, cudaMemcpyHostToDevice, 0));
    rc = transform(ctx, ctx->d_stage, log_n, batch, inverse, 0);
    if (rc) return rc;
    CTX_CUDA(cudaMemcpyAsync(h_out, ctx->d_stage, bytes, cudaMemcpyDeviceToHost, 0));
    CTX_CUDA(cudaStreamSynchronize(0));
    return NTT256_OK;
  }
  int rc = ensure_stage(ctx, 3 * one * per_chunk);
  if (rc) return rc;
  return host_ring(ctx, h_in, h_out, one, batch, per_chunk,
                   [&](void* d, int nb, cudaStream_t s) { return transform(ctx, d, log_n, nb, inverse, s); });
}

// Multiplied-out shift^j tables trade 32 B/element of memory and HBM traffic for one of the two
// products of the scaling pass (2^24: 0.53 -> 0.3 ms).  One table may take 1 GiB (n <= 2^25), all
// cached ones together 2 GiB; larger transforms keep the two small factored tables.
constexpr size_t kCosetDirectMax = (size_t)2 << 30;     // 2^25 points in Montgomery format, 2^24 as Shoup pairs
constexpr size_t kCosetDirectBudget = (size_t)4 << 30;

void drop_oldest_coset(ntt256_ctx* ctx) {
  auto& c = ctx->cosets.front();
  cudaFree(c.hi);
  cudaFree(c.lo);
  if (c.direct) cudaFree(c.direct);
  ctx->cosets.erase(ctx->cosets.begin());
}

int get_coset_tables(ntt256_ctx* ctx, const U256& shift_plain, int log_n, const ntt256_ctx::CosetTables** out) {
  for (auto& c : ctx->cosets)
    if (c.log_n == log_n && u256_cmp(c.shift, shift_plain) == 0) {
      *out = &c;
      return NTT256_OK;
    }
  if (ctx->cosets.size() >= 8) drop_oldest_coset(ctx);
  ntt256_ctx::CosetTables t;
  t.shift = shift_plain;
  t.log_n = log_n;
  t.lo_bits = log_n / 2;
  t.hi = t.lo = nullptr;
  U256 sm = ctx->hf.to_mont(shift_plain);
  int rc = build_power_table(ctx, &t.lo, sm, t.lo_bits);
  if (rc) return rc;
  rc = build_power_table(ctx, &t.hi, pow2k(ctx->hf, sm, t.lo_bits), log_n - t.lo_bits);
  if (rc) {
    cudaFree(t.lo);
    return rc;
  }
  t.direct = nullptr;
  t.direct_shoup = ctx->shoup;
  const size_t dbytes = (size_t)(ctx->shoup ? 64 : 32) << log_n;
  if (dbytes <= kCosetDirectMax) {
    auto cached = [&] {
      size_t sum = 0;
      for (auto& c : ctx->cosets)
        if (c.direct) sum += (size_t)(c.direct_shoup ? 64 : 32) << c.log_n;
      return sum;
    };
    while (!ctx->cosets.empty() && cached() + dbytes > kCosetDirectBudget) drop_oldest_coset(ctx);
    if (cudaMalloc((void**)&t.direct, dbytes) == cudaSuccess) {
      const unsigned long long count = 1ull << log_n;
      Fe pinv;
      to_u32(ctx->hf.pinv, pinv.v);
      coset_table_kernel<<<(unsigned)((count + 127) / 128), 128>>>(t.direct, count, t.lo_bits, t.hi, t.lo, ctx->fp,
                                                                   ctx->shoup ? 1 : 0, pinv);
      ++ctx->launches;
      if (cudaGetLastError() != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) {  // one-off sync
        cudaFree(t.direct);
        cudaFree(t.hi);
        cudaFree(t.lo);
        ctx->last_err = "coset_table_kernel failed";
        return NTT256_ERR_CUDA;
      }
    } else {
      cudaGetLastError();  // no memory to spare: keep the factored tables
      t.direct = nullptr;
    }
  }
  ctx->cosets.push_back(t);
  *out = &ctx->cosets.back();
  return NTT256_OK;
}

int coset_scale(ntt256_ctx* ctx, void* d_data, int log_n, int batch, const U256& shift_plain, cudaStream_t stream) {
  const ntt256_ctx::CosetTables* t = nullptr;
  int rc = get_coset_tables(ctx, shift_plain, log_n, &t);
  if (rc) return rc;
  const unsigned long long n = (unsigned long long)batch << log_n;
  const int threads = 128;
  if (t->direct)
    coset_scale_direct_kernel<<<(unsigned)((n + threads - 1) / threads), threads, 0, stream>>>(
        (uint4*)d_data, n, log_n, t->direct, ctx->fp, t->direct_shoup ? 1 : 0);
  else
    coset_scale_kernel<<<(unsigned)((n + threads - 1) / threads), threads, 0, stream>>>((uint4*)d_data, n, log_n,
                                                                                      t->lo_bits, t->hi, t->lo, ctx->fp);
  ++ctx->launches;
  CTX_CUDA(cudaGetLastError());
  return NTT256_OK;
}

}  // namespace

extern "C" {

// 2^229 < p and 5p + 2^226 < 2^256: what the Shoup / wide-lazy kernels need (BN254 Fr: 2^256 / p = 5.29).  The
// lower bound keeps the slack of the approximate corrections (up to 2^224) below p, so that a - b + 3p > 0.
static bool shoup_capable(const U256& P) {
  if (u256_bitlen(P) > 254 || u256_bitlen(P) < 230) return false;
  U256 p2, p4, p5, t;
  u256_add(p2, P, P);
  if (u256_add(p4, p2, p2) || u256_add(p5, p4, P)) return false;
  return u256_add(t, p5, U256{{0, 0, 0, 1ull << 34}}) == 0;
}

int ntt256_ctx_create(ntt256_ctx** out, int device, const uint32_t p[8], const uint32_t g[8], int max_log_n) {
  if (!out || !p || !g || max_log_n < 1 || max_log_n > 32) return NTT256_ERR_ARG;
  *out = nullptr;
  U256 P = u256_from_limbs(p), G = u256_from_limbs(g);
  if (!(P.w[0] & 1) || u256_bitlen(P) > 255 || u256_bitlen(P) < 2) return NTT256_ERR_FIELD;
  if (u256_is_zero(G) || u256_cmp(G, P) >= 0) return NTT256_ERR_FIELD;
  U256 pm1;
  u256_sub(pm1, P, u256_small(1));
  for (int i = 0; i < max_log_n; ++i)
    if (u256_bit(pm1, i)) return NTT256_ERR_FIELD;
  ntt256_ctx* ctx = new (std::nothrow) ntt256_ctx();
  if (!ctx) return NTT256_ERR_NOMEM;
  ctx->device = device;
  ctx->max_log_n = max_log_n;
  ctx->hf.init(P);
  std::memset(&ctx->fp, 0, sizeof(ctx->fp));
  u256_to_limbs(P, ctx->fp.p);
  U256 P2;
  u256_add(P2, P, P);
  u256_to_limbs(P2, ctx->fp.p2);
  {
    U256 pbar;
    u256_sub(pbar, u256_small(0), P);  // 2^256 - p
    u256_to_limbs(pbar, ctx->fp.pbar);
    U256 p2bar;
    u256_sub(p2bar, u256_small(0), P2);  // 2^256 - 2p
    u256_to_limbs(p2bar, ctx->fp.p2bar);
  }
  ctx->fp.np0 = (u32)ctx->hf.np0;
  // Shoup twiddles come with the wide-lazy add/sub policy (field_wl.cuh): sums reach 4p + 2^225 and differences
  // a - b + 3p reach 5p + 2^224, which must fit 256 bits; shoup_mul's own remainder stays below 3p
  ctx->shoup = shoup_capable(P);
  if (ctx->shoup) {
    U256 p3;
    u256_add(p3, P2, P);
    u256_to_limbs(p3, ctx->fp.kp);
  }
  ctx->g_mont = ctx->hf.to_mont(G);
  ctx->wN = root_of_unity(ctx->hf, ctx->g_mont, max_log_n);
  ctx->wN_inv = ctx->hf.inv(ctx->wN);
  // w_N must have exact order N (g is a generator / non-residue): w_N^(N/2) == -1
  {
    U256 h = pow2k(ctx->hf, ctx->wN, max_log_n - 1), m1;
    m1 = ctx->hf.sub(u256_small(0), ctx->hf.one);
    if (u256_cmp(h, m1) != 0) {
      delete ctx;
      return NTT256_ERR_FIELD;
    }
  }
  ctx->hi_bits = std::min(max_log_n, max_log_n >= 26 ? kHiBitsMaxLarge : kHiBitsMax);
  ctx->lo_bits = max_log_n - ctx->hi_bits;
  ctx->loc_bits = std::min(max_log_n, kLocBitsMax);
  std::memset(ctx->c8, 0, sizeof(ctx->c8));
  std::memset(ctx->w4, 0, sizeof(ctx->w4));
  if (max_log_n >= 3) {
    U256 w8 = pow2k(ctx->hf, ctx->wN, max_log_n - 3);
    U256 w8i = pow2k(ctx->hf, ctx->wN_inv, max_log_n - 3);
    U256 a = w8, b = w8i;
    for (int k = 0; k < 3; ++k) {
      to_u32(a, ctx->c8[0][k]);
      to_u32(b, ctx->c8[1][k]);
      a = ctx->hf.mul(a, w8);
      b = ctx->hf.mul(b, w8i);
    }
  }
  if (max_log_n >= 2) {
    to_u32(pow2k(ctx->hf, ctx->wN, max_log_n - 2), ctx->w4[0]);
    to_u32(pow2k(ctx->hf, ctx->wN_inv, max_log_n - 2), ctx->w4[1]);
    if (max_log_n < 3) {  // radix-4 rounds use c8[1] = w_4
      std::memcpy(ctx->c8[0][1], ctx->w4[0], 32);
      std::memcpy(ctx->c8[1][1], ctx->w4[1], 32);
    }
  }
  // plain copies and Shoup companions of the by-value constants
  std::memset(ctx->c8p, 0, sizeof(ctx->c8p));
  std::memset(ctx->c8q, 0, sizeof(ctx->c8q));
  for (int d = 0; d < 2; ++d)
    for (int k = 0; k < 3; ++k) {
      const U256 plain = ctx->hf.from_mont(u256_from_limbs(ctx->c8[d][k]));
      to_u32(plain, ctx->c8p[d][k]);
      to_u32(ctx->hf.shoup_companion(plain), ctx->c8q[d][k]);
    }
  int rc = NTT256_OK;
  auto fail = [&](int code) {
    ntt256_ctx_destroy(ctx);
    return code;
  };
  if (cudaSetDevice(device) != cudaSuccess) return fail(NTT256_ERR_CUDA);
  {
    // a streamed first-pass twiddle array is as large as the transform itself (twice with Shoup pairs)
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess && total_b) {
      // one array may take a tenth of the device (2^28 points as Shoup pairs = 16 GiB on a 180 GB B200: 58.0 ms
      // instead of 62.7 ms with two lookups, profiles/r02_large_1gpu.jsonl), all of them together a fifth
      ctx->tw_stream_limit = total_b / 10;
      ctx->tw_stream_budget = total_b / 5;
    }
  }
  ctx->const_field = pass_bn254_params_match(ctx->fp);
  if (!pass_bn254_set_smem(32 << kMaxTileBits)) return fail(NTT256_ERR_CUDA);
  const int big = 32 << kMaxTileBits, small = 32 << 11;
  if (!pass_inst_set_smem<false, false>(big, small) || !pass_inst_set_smem<false, true>(big, small) ||
      !pass_inst_set_smem<true, false>(big, small) || !pass_inst_set_smem<true, true>(big, small))
    return fail(NTT256_ERR_CUDA);
  rc = build_twiddle_tables(ctx);
  if (rc) return fail(rc);
  *out = ctx;
  return NTT256_OK;
}

int ntt256_ctx_create_named(ntt256_ctx** out, int device, int field, int max_log_n) {
  static const uint32_t bn254_p[8] = {0xf0000001u, 0x43e1f593u, 0x79b97091u, 0x2833e848u,
                                      0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
  static const uint32_t bls_p[8] = {0x00000001u, 0xffffffffu, 0xfffe5bfeu, 0x53bda402u,
                                    0x09a1d805u, 0x3339d808u, 0x299d7d48u, 0x73eda753u};
  uint32_t g[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (field == NTT256_FIELD_BN254_FR) {
    g[0] = 5;
    return ntt256_ctx_create(out, device, bn254_p, g, max_log_n);
  }
  if (field == NTT256_FIELD_BLS12_381_FR) {
    g[0] = 7;
    return ntt256_ctx_create(out, device, bls_p, g, max_log_n);
  }
  return NTT256_ERR_ARG;
}

void ntt256_ctx_destroy(ntt256_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  free_twiddle_tables(ctx);
  for (auto& c : ctx->cosets) {
    cudaFree(c.hi);
    cudaFree(c.lo);
    if (c.direct) cudaFree(c.direct);
  }
  if (ctx->s_h2d) cudaStreamDestroy(ctx->s_h2d);
  if (ctx->s_d2h) cudaStreamDestroy(ctx->s_d2h);
  if (ctx->s_comp) cudaStreamDestroy(ctx->s_comp);
  for (int i = 0; i < 3; ++i) {
    if (ctx->ev_in[i]) cudaEventDestroy(ctx->ev_in[i]);
    if (ctx->ev_done[i]) cudaEventDestroy(ctx->ev_done[i]);
    if (ctx->ev_out[i]) cudaEventDestroy(ctx->ev_out[i]);
  }
  if (ctx->d_stage) cudaFree(ctx->d_stage);
  if (ctx->sync_counter) cudaFree(ctx->sync_counter);
  for (auto& kv : ctx->plan_tables) free_pass_tables(&kv.second);
  delete ctx;
}

int ntt256_ctx_set_tuning(ntt256_ctx* ctx, int tile_bits, int coalesce_bits) {
  if (!ctx || tile_bits < 3 || tile_bits > kMaxTileBits || coalesce_bits < 0 || coalesce_bits > 6 ||
      coalesce_bits > tile_bits)
    return NTT256_ERR_ARG;
  ctx->tile_bits = tile_bits;
  ctx->coalesce_bits = coalesce_bits;
  ctx->auto_tile = false;
  cudaSetDevice(ctx->device);
  cudaDeviceSynchronize();
  ctx->plans.clear();
  for (auto& kv : ctx->plan_tables) free_pass_tables(&kv.second);
  ctx->plan_tables.clear();
  drop_tw_streams(ctx);
  return NTT256_OK;
}

int ntt256_forward(ntt256_ctx* ctx, void* d_data, int log_n, ntt256_stream stream) {
  return transform(ctx, d_data, log_n, 1, false, (cudaStream_t)stream);
}
int ntt256_inverse(ntt256_ctx* ctx, void* d_data, int log_n, ntt256_stream stream) {
  return transform(ctx, d_data, log_n, 1, true, (cudaStream_t)stream);
}
int ntt256_forward_batch(ntt256_ctx* ctx, void* d_data, int log_n, int batch, ntt256_stream stream) {
  return transform(ctx, d_data, log_n, batch, false, (cudaStream_t)stream);
}
int ntt256_inverse_batch(ntt256_ctx* ctx, void* d_data, int log_n, int batch, ntt256_stream stream) {
  return transform(ctx, d_data, log_n, batch, true, (cudaStream_t)stream);
}
int ntt256_forward_host(ntt256_ctx* ctx, const void* h_in, void* h_out, int log_n, int batch) {
  return transform_host(ctx, h_in, h_out, log_n, batch, false);
}
int ntt256_inverse_host(ntt256_ctx* ctx, const void* h_in, void* h_out, int log_n, int batch) {
  return transform_host(ctx, h_in, h_out, log_n, batch, true);
}

int ntt256_vec_mul(ntt256_ctx* ctx, void* d_out, const void* d_a, const void* d_b, unsigned long long n, int mont_form,
                   ntt256_stream stream) {
  if (!ctx || !d_out || !d_a || !d_b || (mont_form != 0 && mont_form != 1)) return NTT256_ERR_ARG;
  if (n == 0) return NTT256_OK;
  CTX_CUDA(cudaSetDevice(ctx->device));
  Fe r2;
  to_u32(ctx->hf.r2, r2.v);
  const int threads = 128;
  vec_mul_kernel<<<(unsigned)((n + threads - 1) / threads), threads, 0, (cudaStream_t)stream>>>(
      (uint4*)d_out, (const uint4*)d_a, (const uint4*)d_b, n, mont_form, ctx->fp, r2);
  ++ctx->launches;
  CTX_CUDA(cudaGetLastError());
  return NTT256_OK;
}

static int vec_scale(ntt256_ctx* ctx, void* d_data, unsigned long long n, const U256& c_mont, cudaStream_t stream) {
  if (!ctx || !d_data) return NTT256_ERR_ARG;
  if (n == 0) return NTT256_OK;
  CTX_CUDA(cudaSetDevice(ctx->device));
  Fe c;
  to_u32(c_mont, c.v);
  const int threads = 128;
  vec_scale_kernel<<<(unsigned)((n + threads - 1) / threads), threads, 0, stream>>>((uint4*)d_data, n, ctx->fp, c);
  ++ctx->launches;
  CTX_CUDA(cudaGetLastError());
  return NTT256_OK;
}

int ntt256_to_montgomery(ntt256_ctx* ctx, void* d_data, unsigned long long n, ntt256_stream stream) {
  if (!ctx) return NTT256_ERR_ARG;
  return vec_scale(ctx, d_data, n, ctx->hf.r2, (cudaStream_t)stream);  // montmul(R^2, x) = x*R
}
int ntt256_from_montgomery(ntt256_ctx* ctx, void* d_data, unsigned long long n, ntt256_stream stream) {
  if (!ctx) return NTT256_ERR_ARG;
  return vec_scale(ctx, d_data, n, u256_small(1), (cudaStream_t)stream);  // montmul(1, x) = x/R
}

static bool fuse_here(const ntt256_ctx* ctx, int log_n, int batch) {
  if (log_n < 3) return false;
  if (ctx->fuse_companions >= 0) return ctx->fuse_companions != 0;
  return ((long long)batch << log_n) <= (1ll << 17);
}
int ntt256_coset_forward(ntt256_ctx* ctx, void* d_data, int log_n, int batch, const uint32_t shift[8],
                         ntt256_stream stream) {
  if (!ctx || !d_data || !shift || log_n < 0 || log_n > ctx->max_log_n || batch < 1) return NTT256_ERR_ARG;
  U256 s = u256_from_limbs(shift);
  if (u256_is_zero(s) || u256_cmp(s, ctx->hf.p) >= 0) return NTT256_ERR_ARG;
  CTX_CUDA(cudaSetDevice(ctx->device));
  // shift^j multiplies the elements as the first pass loads them (one more product on values already in
  // registers, no extra trip through HBM) whenever the multiplied-out table exists
  const ntt256_ctx::CosetTables* t = nullptr;
  int rc = get_coset_tables(ctx, s, log_n, &t);
  if (rc) return rc;
  if (t->direct && fuse_here(ctx, log_n, batch)) {
    FuseSpec fz;
    fz.in_scale = t->direct;
    return transform(ctx, d_data, log_n, batch, false, (cudaStream_t)stream, nullptr, &fz);
  }
  rc = coset_scale(ctx, d_data, log_n, batch, s, (cudaStream_t)stream);
  if (rc) return rc;
  return transform(ctx, d_data, log_n, batch, false, (cudaStream_t)stream);
}

int ntt256_coset_inverse(ntt256_ctx* ctx, void* d_data, int log_n, int batch, const uint32_t shift[8],
                         ntt256_stream stream) {
  if (!ctx || !d_data || !shift || log_n < 0 || log_n > ctx->max_log_n || batch < 1) return NTT256_ERR_ARG;
  U256 s = u256_from_limbs(shift);
  if (u256_is_zero(s) || u256_cmp(s, ctx->hf.p) >= 0) return NTT256_ERR_ARG;
  CTX_CUDA(cudaSetDevice(ctx->device));
  U256 sinv = ctx->hf.from_mont(ctx->hf.inv(ctx->hf.to_mont(s)));
  const ntt256_ctx::CosetTables* t = nullptr;
  int rc = get_coset_tables(ctx, sinv, log_n, &t);
  if (rc) return rc;
  if (t->direct && fuse_here(ctx, log_n, batch)) {
    FuseSpec fz;
    fz.out_scale = t->direct;  // shift^-k multiplies the outputs of the last pass before they are stored
    return transform(ctx, d_data, log_n, batch, true, (cudaStream_t)stream, nullptr, &fz);
  }
  rc = transform(ctx, d_data, log_n, batch, true, (cudaStream_t)stream);
  if (rc) return rc;
  return coset_scale(ctx, d_data, log_n, batch, sinv, (cudaStream_t)stream);
}

static int transform_mul(ntt256_ctx* ctx, void* d_data, const void* d_mul_mont, int log_n, int batch, bool inverse,
                         cudaStream_t stream) {
  if (!ctx || !d_data || !d_mul_mont || log_n < 0 || log_n > ctx->max_log_n || batch < 1) return NTT256_ERR_ARG;
  CTX_CUDA(cudaSetDevice(ctx->device));
  if (fuse_here(ctx, log_n, batch)) {
    FuseSpec fz;
    fz.out_mul = (const uint4*)d_mul_mont;
    return transform(ctx, d_data, log_n, batch, inverse, stream, nullptr, &fz);
  }
  int rc = transform(ctx, d_data, log_n, batch, inverse, stream);
  if (rc) return rc;
  return ntt256_vec_mul(ctx, d_data, d_data, d_mul_mont, (unsigned long long)batch << log_n, 1, (ntt256_stream)stream);
}
int ntt256_forward_mul(ntt256_ctx* ctx, void* d_data, const void* d_mul_mont, int log_n, int batch, ntt256_stream stream) {
  return transform_mul(ctx, d_data, d_mul_mont, log_n, batch, false, (cudaStream_t)stream);
}
int ntt256_inverse_mul(ntt256_ctx* ctx, void* d_data, const void* d_mul_mont, int log_n, int batch, ntt256_stream stream) {
  return transform_mul(ctx, d_data, d_mul_mont, log_n, batch, true, (cudaStream_t)stream);
}
int ntt256_ctx_set_fused_companions(ntt256_ctx* ctx, int enable) {
  if (!ctx) return NTT256_ERR_ARG;
  if (enable < -1 || enable > 1) return NTT256_ERR_ARG;
  ctx->fuse_companions = enable;
  return NTT256_OK;
}

static int fourstep(ntt256_ctx* ctx, int kind, void* d_shard, int log_n, int log_r, int log_p, int rank,
                    cudaStream_t stream, const void* const* peer_recv = nullptr, const IoSpec* io = nullptr) {
  const bool inverse = kind >= 3;
  const bool first_step = (kind == 1 || kind == 3);
  const int scale_log_n = (kind == 4) ? log_n : -1;
  if (!ctx || !d_shard || log_n > ctx->max_log_n || !fourstep_args_ok(log_n, log_r, log_p, rank)) return NTT256_ERR_ARG;
  CTX_CUDA(cudaSetDevice(ctx->device));
  const Plan* plan = nullptr;
  // step 2 does not depend on the rank (no foreign time bits remain)
  PlanKey key{kind, log_n, log_r, log_p, first_step ? rank : 0};
  if (ctx->force_group_bits) key.group_bits = ctx->force_group_bits;
  int rc = get_plan(ctx, key, &plan);
  if (rc) return rc;
  // ranks armed with ntt256_ctx_set_peer_sync: step 1 with peer stores ends by publishing a new epoch,
  // step 2 begins by waiting for it (from every rank)
  const bool armed = ctx->sync_nranks == (1 << log_p) && log_p > 0;
  if (peer_recv && log_p > 0) {
    if (log_p > 3) return NTT256_ERR_ARG;
    PeerSpec ps;
    ps.ptrs = peer_recv;
    ps.bits = log_p;
    ps.shift = log_n - 2 * log_p;  // local index bits below the destination-rank field
    ps.rank = rank;
    if (armed) ++ctx->sync_epoch;
    return run_plan(ctx, *plan, d_shard, 0, 1, inverse, scale_log_n, stream, &ps, io, armed ? 2 : 0);
  }
  return run_plan(ctx, *plan, d_shard, 0, 1, inverse, scale_log_n, stream, nullptr, io,
                  (armed && !first_step) ? 1 : 0);
}

int ntt256_fourstep_inverse_step1(ntt256_ctx* ctx, void* d_shard, int log_n, int log_r, int log_p, int rank,
                                  ntt256_stream stream) {
  return fourstep(ctx, 3, d_shard, log_n, log_r, log_p, rank, (cudaStream_t)stream);
}
int ntt256_fourstep_inverse_step1_peer(ntt256_ctx* ctx, void* d_shard, const void* const* peer_recv, int log_n,
                                       int log_r, int log_p, int rank, ntt256_stream stream) {
  if (!peer_recv) return NTT256_ERR_ARG;
  return fourstep(ctx, 3, d_shard, log_n, log_r, log_p, rank, (cudaStream_t)stream, peer_recv);
}
int ntt256_fourstep_inverse_step2(ntt256_ctx* ctx, void* d_shard, int log_n, int log_r, int log_p, int rank,
                                  ntt256_stream stream) {
  return fourstep(ctx, 4, d_shard, log_n, log_r, log_p, rank, (cudaStream_t)stream);
}

int ntt256_fourstep_step1_peer(ntt256_ctx* ctx, void* d_shard, const void* const* peer_recv, int log_n, int log_r,
                               int log_p, int rank, ntt256_stream stream) {
  if (!peer_recv) return NTT256_ERR_ARG;
  return fourstep(ctx, 1, d_shard, log_n, log_r, log_p, rank, (cudaStream_t)stream, peer_recv);
}

int ntt256_fourstep_step1(ntt256_ctx* ctx, void* d_shard, int log_n, int log_r, int log_p, int rank,
                          ntt256_stream stream) {
  return fourstep(ctx, 1, d_shard, log_n, log_r, log_p, rank, (cudaStream_t)stream);
}
int ntt256_fourstep_step2(ntt256_ctx* ctx, void* d_shard, int log_n, int log_r, int log_p, int rank,
                          ntt256_stream stream) {
  return fourstep(ctx, 2, d_shard, log_n, log_r, log_p, rank, (cudaStream_t)stream);
}

// ---------------------------------------------------------------- single-process multi-GPU driver
struct ntt256_mgpu {
  int P = 0, log_p = 0, max_log_n = 0;
  std::vector<int> dev;
  std::vector<ntt256_ctx*> ctx;
  std::vector<void*> recv;
  size_t recv_bytes = 0;
  std::vector<cudaStream_t> stream;
  std::vector<cudaEvent_t> ev1, ev2;
  bool ev2_valid = false;
};

void ntt256_mgpu_destroy(ntt256_mgpu* m) {
  if (!m) return;
  for (int i = 0; i < (int)m->dev.size(); ++i) {
    cudaSetDevice(m->dev[i]);
    cudaDeviceSynchronize();
    if (i < (int)m->recv.size() && m->recv[i]) cudaFree(m->recv[i]);
    if (i < (int)m->stream.size() && m->stream[i]) cudaStreamDestroy(m->stream[i]);
    if (i < (int)m->ev1.size() && m->ev1[i]) cudaEventDestroy(m->ev1[i]);
    if (i < (int)m->ev2.size() && m->ev2[i]) cudaEventDestroy(m->ev2[i]);
    if (i < (int)m->ctx.size() && m->ctx[i]) ntt256_ctx_destroy(m->ctx[i]);
  }
  delete m;
}

int ntt256_mgpu_create(ntt256_mgpu** out, int ngpus, const int* devices, int field, int max_log_n) {
  if (!out || ngpus < 1 || ngpus > 8 || (ngpus & (ngpus - 1))) return NTT256_ERR_ARG;
  *out = nullptr;
  ntt256_mgpu* m = new (std::nothrow) ntt256_mgpu();
  if (!m) return NTT256_ERR_NOMEM;
  m->P = ngpus;
  while ((1 << m->log_p) < ngpus) ++m->log_p;
  m->max_log_n = max_log_n;
  m->dev.resize(ngpus);
  m->ctx.assign(ngpus, nullptr);
  m->recv.assign(ngpus, nullptr);
  m->stream.assign(ngpus, nullptr);
  m->ev1.assign(ngpus, nullptr);
  m->ev2.assign(ngpus, nullptr);
  for (int i = 0; i < ngpus; ++i) m->dev[i] = devices ? devices[i] : i;
  for (int i = 0; i < ngpus; ++i) {
    int rc = ntt256_ctx_create_named(&m->ctx[i], m->dev[i], field, max_log_n);
    if (rc) {
      ntt256_mgpu_destroy(m);
      return rc;
    }
    if (cudaSetDevice(m->dev[i]) != cudaSuccess ||
        cudaStreamCreateWithFlags(&m->stream[i], cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&m->ev1[i], cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&m->ev2[i], cudaEventDisableTiming) != cudaSuccess) {
      ntt256_mgpu_destroy(m);
      return NTT256_ERR_CUDA;
    }
    for (int j = 0; j < ngpus; ++j)
      if (j != i && ntt256_enable_peer_access(m->dev[i], m->dev[j]) != NTT256_OK) {
        ntt256_mgpu_destroy(m);
        return NTT256_ERR_CUDA;
      }
  }
  *out = m;
  return NTT256_OK;
}

static int mgpu_transform(ntt256_mgpu* m, void* const* d_shards, int log_n, bool inverse) {
  if (!m || !d_shards || log_n > m->max_log_n) return NTT256_ERR_ARG;
  const int P = m->P, lp = m->log_p;
  const int lr = (log_n + 1) / 2;
  if (!fourstep_args_ok(log_n, lr, lp, 0)) return NTT256_ERR_ARG;
  for (int i = 0; i < P; ++i)
    if (!d_shards[i]) return NTT256_ERR_ARG;
  const size_t bytes = (size_t)32 << (log_n - lp);
  if (bytes > m->recv_bytes) {
    for (int i = 0; i < P; ++i) {
      if (cudaSetDevice(m->dev[i]) != cudaSuccess) return NTT256_ERR_CUDA;
      cudaDeviceSynchronize();
      if (m->recv[i]) cudaFree(m->recv[i]);
      m->recv[i] = nullptr;
      if (cudaMalloc(&m->recv[i], bytes) != cudaSuccess) return NTT256_ERR_NOMEM;
    }
    m->recv_bytes = bytes;
    m->ev2_valid = false;
  }
  const int k1 = inverse ? 3 : 1, k2 = inverse ? 4 : 2;
  // phase 1: every device runs its local passes; the last one stores into the peers' receive buffers
  for (int i = 0; i < P; ++i) {
    if (cudaSetDevice(m->dev[i]) != cudaSuccess) return NTT256_ERR_CUDA;
    if (m->ev2_valid)  // the previous transform's phase 2 may still be reading the receive buffers
      for (int j = 0; j < P; ++j) cudaStreamWaitEvent(m->stream[i], m->ev2[j], 0);
    int rc = fourstep(m->ctx[i], k1, d_shards[i], log_n, lr, lp, i, m->stream[i], P > 1 ? m->recv.data() : nullptr);
    if (rc) return rc;
    if (cudaEventRecord(m->ev1[i], m->stream[i]) != cudaSuccess) return NTT256_ERR_CUDA;
  }
  // phase 2: after every device's phase 1 has finished, local passes on the received data; the
  // last pass writes the result into the caller's shard buffer
  for (int i = 0; i < P; ++i) {
    if (cudaSetDevice(m->dev[i]) != cudaSuccess) return NTT256_ERR_CUDA;
    for (int j = 0; j < P; ++j) cudaStreamWaitEvent(m->stream[i], m->ev1[j], 0);
    int rc;
    if (P > 1) {
      IoSpec io;
      io.dst_last = d_shards[i];
      rc = fourstep(m->ctx[i], k2, m->recv[i], log_n, lr, lp, i, m->stream[i], nullptr, &io);
    } else {
      rc = fourstep(m->ctx[i], k2, d_shards[i], log_n, lr, lp, i, m->stream[i]);
    }
    if (rc) return rc;
    if (cudaEventRecord(m->ev2[i], m->stream[i]) != cudaSuccess) return NTT256_ERR_CUDA;
  }
  m->ev2_valid = true;
  return NTT256_OK;
}

int ntt256_mgpu_forward(ntt256_mgpu* m, void* const* d_shards, int log_n) { return mgpu_transform(m, d_shards, log_n, false); }
int ntt256_mgpu_inverse(ntt256_mgpu* m, void* const* d_shards, int log_n) { return mgpu_transform(m, d_shards, log_n, true); }

int ntt256_mgpu_forward_batch(ntt256_mgpu* m, void* const* d_shards, int log_n, int batch_per_gpu) {
  if (!m || !d_shards || batch_per_gpu < 1) return NTT256_ERR_ARG;
  for (int i = 0; i < m->P; ++i) {
    int rc = transform(m->ctx[i], d_shards[i], log_n, batch_per_gpu, false, m->stream[i]);
    if (rc) return rc;
  }
  return NTT256_OK;
}

int ntt256_mgpu_synchronize(ntt256_mgpu* m) {
  if (!m) return NTT256_ERR_ARG;
  for (int i = 0; i < m->P; ++i) {
    if (cudaSetDevice(m->dev[i]) != cudaSuccess) return NTT256_ERR_CUDA;
    if (cudaStreamSynchronize(m->stream[i]) != cudaSuccess) return NTT256_ERR_CUDA;
  }
  return NTT256_OK;
}

int ntt256_ipc_alloc(int device, size_t bytes, void** d_ptr, unsigned char handle[64]) {
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  if (!d_ptr || !handle || !bytes) return NTT256_ERR_ARG;
  if (cudaSetDevice(device) != cudaSuccess) return NTT256_ERR_CUDA;
  if (cudaMalloc(d_ptr, bytes) != cudaSuccess) return NTT256_ERR_NOMEM;
  cudaIpcMemHandle_t h;
  if (cudaIpcGetMemHandle(&h, *d_ptr) != cudaSuccess) {
    cudaFree(*d_ptr);
    *d_ptr = nullptr;
    return NTT256_ERR_CUDA;
  }
  std::memcpy(handle, &h, 64);
  return NTT256_OK;
}

int ntt256_ipc_open(int device, const unsigned char handle[64], void** d_ptr) {
  if (!d_ptr || !handle) return NTT256_ERR_ARG;
  if (cudaSetDevice(device) != cudaSuccess) return NTT256_ERR_CUDA;
  cudaIpcMemHandle_t h;
  std::memcpy(&h, handle, 64);
  // opened with the CONSUMER's device current: the mapping is a peer mapping usable by its kernels
  if (cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) return NTT256_ERR_CUDA;
  return NTT256_OK;
}

int ntt256_ipc_close(int device, void* d_ptr) {
  if (cudaSetDevice(device) != cudaSuccess) return NTT256_ERR_CUDA;
  return cudaIpcCloseMemHandle(d_ptr) == cudaSuccess ? NTT256_OK : NTT256_ERR_CUDA;
}

int ntt256_ipc_free(int device, void* d_ptr) {
  if (cudaSetDevice(device) != cudaSuccess) return NTT256_ERR_CUDA;
  return cudaFree(d_ptr) == cudaSuccess ? NTT256_OK : NTT256_ERR_CUDA;
}

int ntt256_ctx_set_stream_twiddle_limit(ntt256_ctx* ctx, size_t max_bytes) {
  if (!ctx) return NTT256_ERR_ARG;
  cudaSetDevice(ctx->device);
  cudaDeviceSynchronize();
  ctx->tw_stream_limit = max_bytes;
  drop_tw_streams(ctx);
  return NTT256_OK;
}

int ntt256_ctx_set_stream_twiddle_budget(ntt256_ctx* ctx, size_t total_bytes) {
  if (!ctx) return NTT256_ERR_ARG;
  cudaSetDevice(ctx->device);
  ctx->tw_stream_budget = total_bytes;
  evict_tw_streams(ctx, 0);
  return NTT256_OK;
}

size_t ntt256_ctx_table_bytes(const ntt256_ctx* ctx, size_t* streamed_bytes) {
  if (!ctx) return 0;
  const size_t e = ctx->shoup ? 64 : 32;
  size_t fixed = ((size_t)e << ctx->hi_bits) + ((size_t)e << ctx->lo_bits) + ((size_t)e << ctx->loc_bits);
  fixed += ctx->loc_scaled.size() * ((size_t)e << ctx->loc_bits);
  size_t coset = 0;
  for (auto& c : ctx->cosets) {
    coset += (size_t)64 << c.lo_bits;  // two factored tables
    if (c.direct) coset += (size_t)(c.direct_shoup ? 64 : 32) << c.log_n;
  }
  if (streamed_bytes) *streamed_bytes = ctx->tw_stream_bytes;
  return fixed + coset + ctx->tw_stream_bytes;
}

int ntt256_ctx_set_peer_sync(ntt256_ctx* ctx, void* const* flag_bufs, int nranks, int rank, double timeout_s) {
  return set_peer_sync_impl(ctx, flag_bufs, nranks, rank, timeout_s);
}

int ntt256_ctx_set_twiddle_mode(ntt256_ctx* ctx, int shoup) {
  if (!ctx || (shoup != 0 && shoup != 1)) return NTT256_ERR_ARG;
  if (shoup && !shoup_capable(ctx->hf.p)) return NTT256_ERR_FIELD;
  if ((shoup != 0) == ctx->shoup) return NTT256_OK;
  CTX_CUDA(cudaSetDevice(ctx->device));
  CTX_CUDA(cudaDeviceSynchronize());
  free_twiddle_tables(ctx);
  while (!ctx->cosets.empty()) drop_oldest_coset(ctx);
  ctx->shoup = shoup != 0;
  return build_twiddle_tables(ctx);
}

int ntt256_ctx_get_twiddle_mode(const ntt256_ctx* ctx) { return ctx && ctx->shoup ? 1 : 0; }

int ntt256_ctx_set_elements_per_thread(ntt256_ctx* ctx, int log2_elements) {
  if (!ctx || log2_elements < 0 || log2_elements > 3) return NTT256_ERR_ARG;
  ctx->force_group_bits = log2_elements;
  return NTT256_OK;
}

int ntt256_ctx_set_dependent_launch(ntt256_ctx* ctx, int enable) {
  if (!ctx) return NTT256_ERR_ARG;
  ctx->use_pdl = enable != 0;
  return NTT256_OK;
}

int ntt256_ctx_set_constant_field_kernel(ntt256_ctx* ctx, int enable) {
  if (!ctx) return NTT256_ERR_ARG;
  ctx->use_const_field = enable != 0;
  return NTT256_OK;
}

int ntt256_ctx_set_dense_pass4(ntt256_ctx* ctx, int enable) {
  if (!ctx) return NTT256_ERR_ARG;
  ctx->dense_pass4 = enable != 0;
  return NTT256_OK;
}

int ntt256_ctx_set_small_transform_limit(ntt256_ctx* ctx, long long max_elements) {
  if (!ctx || max_elements < 0) return NTT256_ERR_ARG;
  ctx->small_limit = max_elements;
  return NTT256_OK;
}

namespace {
unsigned long long g_barrier_timeout_ns = 60ull * 1000000000ull;
struct PeerBarrierArgs {
  unsigned long long flags[8];  // every rank's flag buffer, as mapped in this process
  unsigned long long epoch;
  unsigned long long timeout_ns;
  int nranks, rank;
};

// Thread r: publish `epoch` in rank r's slot [rank], then wait for rank r's epoch in the local slot [r].
__global__ void peer_barrier_kernel(const __grid_constant__ PeerBarrierArgs A) {
  // programmatic dependent launch: the pass kernel behind this barrier may become resident now (it waits for this
  // kernel's completion before it touches data); this kernel itself waits for its predecessor -- step 1's last
  // pass and all its peer stores -- before it publishes anything
  grid_dependency_launch_dependents();
  grid_dependency_wait();
  const int r = threadIdx.x;
  if (r >= A.nranks) return;
  constexpr int kSlotBase = kSyncSlotBase;  // flags start at byte 256 (the first words carry the IPC check stamp)
  __threadfence_system();
  volatile unsigned long long* theirs = (volatile unsigned long long*)A.flags[r] + kSlotBase + A.rank;
  *theirs = A.epoch;
  volatile unsigned long long* mine = (volatile unsigned long long*)A.flags[A.rank] + kSlotBase + r;
  const unsigned long long t0 = global_timer_ns();
  while (*mine < A.epoch) {
    if (global_timer_ns() - t0 > A.timeout_ns) {
      // a peer died or is very late: leave a mark the host can read (ntt256_peer_sync_status) and stop
      // waiting -- no trap, the context and its IPC mappings stay usable
      ((volatile unsigned long long*)A.flags[A.rank])[kSyncErrSlot] = A.epoch;
      break;
    }
  }
  __threadfence_system();
}
}  // namespace

int ntt256_peer_barrier(int device, void* const* flag_bufs, int nranks, int rank, unsigned long long epoch,
                        ntt256_stream stream) {
  if (!flag_bufs || nranks < 1 || nranks > 8 || rank < 0 || rank >= nranks || epoch == 0) return NTT256_ERR_ARG;
  if (cudaSetDevice(device) != cudaSuccess) return NTT256_ERR_CUDA;
  PeerBarrierArgs A;
  std::memset(&A, 0, sizeof(A));
  for (int r = 0; r < nranks; ++r) {
    if (!flag_bufs[r]) return NTT256_ERR_ARG;
    A.flags[r] = (unsigned long long)flag_bufs[r];
  }
  A.epoch = epoch;
  A.timeout_ns = g_barrier_timeout_ns;
  A.nranks = nranks;
  A.rank = rank;
  {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(1, 1, 1);
    cfg.blockDim = dim3(32, 1, 1);
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, peer_barrier_kernel, A);
  }
  return cudaGetLastError() == cudaSuccess ? NTT256_OK : NTT256_ERR_CUDA;
}

int ntt256_set_peer_barrier_timeout(double seconds) {
  if (!(seconds > 0)) return NTT256_ERR_ARG;
  g_barrier_timeout_ns = (unsigned long long)(seconds * 1e9);
  return NTT256_OK;
}

int ntt256_peer_sync_status(int device, const void* flag_buf, unsigned long long* timed_out_epoch) {
  if (!flag_buf || !timed_out_epoch) return NTT256_ERR_ARG;
  if (cudaSetDevice(device) != cudaSuccess) return NTT256_ERR_CUDA;
  unsigned long long v = 0;
  if (cudaMemcpy(&v, (const char*)flag_buf + 8 * kSyncErrSlot, 8, cudaMemcpyDeviceToHost) != cudaSuccess)
    return NTT256_ERR_CUDA;
  *timed_out_epoch = v;
  return v ? NTT256_ERR_CUDA : NTT256_OK;
}

int ntt256_enable_peer_access(int device, int peer_device) {
  if (device == peer_device) return NTT256_OK;
  int can = 0;
  if (cudaSetDevice(device) != cudaSuccess) return NTT256_ERR_CUDA;
  if (cudaDeviceCanAccessPeer(&can, device, peer_device) != cudaSuccess || !can) return NTT256_ERR_CUDA;
  cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
  if (e == cudaErrorPeerAccessAlreadyEnabled) {
    cudaGetLastError();
    return NTT256_OK;
  }
  return e == cudaSuccess ? NTT256_OK : NTT256_ERR_CUDA;
}

size_t ntt256_plan_describe_ex(int log_n, int max_log_n, int tile_bits, int coalesce_bits, int group_bits, char* buf,
                               size_t buf_len);
size_t ntt256_plan_describe(int log_n, int max_log_n, int tile_bits, int coalesce_bits, char* buf, size_t buf_len) {
  return ntt256_plan_describe_ex(log_n, max_log_n, tile_bits, coalesce_bits, 3, buf, buf_len);
}
size_t ntt256_plan_describe_ex(int log_n, int max_log_n, int tile_bits, int coalesce_bits, int group_bits, char* buf,
                               size_t buf_len) {
  PlanRequest req;
  req.group_bits = group_bits;
  req.log_n = log_n;
  req.max_log_n = max_log_n;
  req.tile_bits = tile_bits;
  req.coalesce_bits = coalesce_bits;
  Plan plan;
  std::string s;
  if (log_n < 3 || !make_plan(req, &plan))
    s = "{\"error\": \"no plan\"}";
  else
    s = plan.describe();
  if (buf && buf_len) {
    size_t n = std::min(buf_len - 1, s.size());
    std::memcpy(buf, s.data(), n);
    buf[n] = 0;
  }
  return s.size() + 1;
}

size_t ntt256_plan_describe_fourstep(int step, int log_n, int log_r, int log_p, int rank, int max_log_n, int tile_bits,
                                     int coalesce_bits, char* buf, size_t buf_len) {
  std::string s;
  Plan plan;
  if (step < 1 || step > 4 || !fourstep_args_ok(log_n, log_r, log_p, rank) || max_log_n < log_n ||
      !make_plan(build_request(PlanKey{step, log_n, log_r, log_p, rank}, tile_bits, coalesce_bits, max_log_n), &plan))
    s = "{\"error\": \"no plan\"}";
  else
    s = plan.describe();
  if (buf && buf_len) {
    size_t n = std::min(buf_len - 1, s.size());
    std::memcpy(buf, s.data(), n);
    buf[n] = 0;
  }
  return s.size() + 1;
}

unsigned long long ntt256_launch_count(const ntt256_ctx* ctx) { return ctx ? ctx->launches : 0; }

const char* ntt256_strerror(int status) {
  switch (status) {
    case NTT256_OK: return "ok";
    case NTT256_ERR_ARG: return "invalid argument";
    case NTT256_ERR_FIELD: return "unsupported field (need odd p < 2^255, 2^max_log_n | p-1, g of full 2-power order)";
    case NTT256_ERR_CUDA: return "CUDA error";
    case NTT256_ERR_NOMEM: return "out of memory";
    case NTT256_ERR_PLAN: return "no pass plan";
    default: return "unknown status";
  }
}
const char* ntt256_last_cuda_error(const ntt256_ctx* ctx) { return ctx ? ctx->last_err.c_str() : ""; }
const char* ntt256_version(void) { return "ntt256-b200 0.1 (sm_100a)"; }

}  // extern "C"
