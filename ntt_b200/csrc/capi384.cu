// C ABI of the 384-bit transforms (include/ntt384.h): the reference's `bits` template parameter
// (/root/reference/src/big-num.cu:57,67,172,260) instantiated at 384 -- fields such as BLS12-377 Fq
// (= BW6-761 Fr; 377 bits, 2-adicity 46).  Same planner and pass structure as the 256-bit library
// (plan.cpp; the kernels are the 12-limb instances of the one pass body, ntt_kernels_n.cuh), field layer
// field_n.cuh.  SURVEY.md par. 8f row 4: forward / inverse / batches, the element-wise product, host buffers
// (zero-copy end passes over pinned memory), coset transforms and pointwise factors fused into the end passes,
// and the local steps of the multi-GPU four-step transform with the exchange fused into the stores.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstring>
#include <map>
#include <new>
#include <string>
#include <tuple>
#include <vector>

#include "../../include/ntt384.h"
#include "host_common.h"
#include "hostfield_n.h"
#include "ntt_kernels_n.cuh"
#include "pass_tables.h"
#include "plan.h"

using namespace ntt256;

namespace {
constexpr int NL = 12;
constexpr int W = 6;
typedef UIntN<W> U384;
constexpr int kTileBits384 = 9;    // 24 KiB of shared memory, 128 threads (sweep: profiles/r01_ntt384.jsonl)
constexpr int kCoalesce384 = 1;    // runs of >= 2 elements = 96 bytes
constexpr int kHiBits384 = 16;     // 2^16 x 48 B = 3 MiB
constexpr int kLocBits384 = 12;
}  // namespace

struct ntt384_ctx {
  int device = 0;
  int max_log_n = 0;
  HostFieldN<W> hf;
  FieldParamsN<NL> fp;
  U384 wN, wN_inv;
  uint4 *om_hi = nullptr, *om_lo = nullptr, *loc = nullptr;
  int hi_bits = 0, lo_bits = 0, loc_bits = 0;
  u32 c8[2][3][NL];
  u32 w4[2][NL];
  U384 ninv[64];
  bool ninv_ok[64] = {};
  int tile_bits = kTileBits384, coalesce_bits = kCoalesce384;
  std::map<std::tuple<int, int, int, int, int>, Plan> plans;  // by (kind, log_n, log_r, log_p, rank); kind 0 = whole transform
  std::map<unsigned long long, PassTables> plan_tables;          // by Plan::id: the index tables of its passes
  unsigned long long next_plan_id = 1;
  std::map<int, uint4*> loc_scaled;  // log_n -> (1/n) * loc[]
  struct Coset {
    U384 shift;  // plain
    int log_n;
    uint4* direct;  // shift^j, j < 2^log_n, Montgomery form
  };
  std::vector<Coset> cosets;
  void* d_stage = nullptr;
  size_t stage_bytes = 0;
  // rank synchronisation folded into the four-step passes (ntt384_ctx_set_peer_sync)
  unsigned long long sync_flags[8] = {};
  int sync_nranks = 0, sync_rank = 0;
  unsigned long long sync_epoch = 0;
  unsigned long long sync_timeout_ns = 60ull * 1000000000ull;
  unsigned int* sync_counter = nullptr;
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

U384 pow2k(const HostFieldN<W>& hf, U384 b, int k) {
  for (int i = 0; i < k; ++i) b = hf.mul(b, b);
  return b;
}

int build_power_table(ntt384_ctx* ctx, uint4** out, const U384& base_mont, int bits) {
  const unsigned long long count = 1ull << bits;
  CTX_CUDA(cudaMalloc((void**)out, count * 4 * NL));
  std::vector<u32> pw((size_t)NL * 64, 0);
  U384 b = base_mont;
  for (int k = 0; k < std::max(bits, 1); ++k) {
    un_to_limbs<W>(b, &pw[(size_t)NL * k]);
    b = ctx->hf.mul(b, b);
  }
  uint4* d_pw = nullptr;
  CTX_CUDA(cudaMalloc((void**)&d_pw, pw.size() * 4));
  CTX_CUDA(cudaMemcpy(d_pw, pw.data(), pw.size() * 4, cudaMemcpyHostToDevice));
  FeN<NL> one;
  un_to_limbs<W>(ctx->hf.one, one.v);
  const int threads = 128;
  twiddle_powers_kernel_n<NL><<<(unsigned)((count + threads - 1) / threads), threads>>>(*out, d_pw, bits, count, ctx->fp, one);
  ++ctx->launches;
  CTX_CUDA(cudaGetLastError());
  CTX_CUDA(cudaDeviceSynchronize());
  CTX_CUDA(cudaFree(d_pw));
  return NTT256_OK;
}

int get_plan(ntt384_ctx* ctx, const FourStepKey& key, const Plan** out) {
  auto k = std::make_tuple(key.kind, key.log_n, key.log_r, key.log_p, key.rank);
  auto it = ctx->plans.find(k);
  if (it == ctx->plans.end()) {
    const int nloc = key.kind == 0 ? key.log_n : key.log_n - key.log_p;
    PlanRequest req = fourstep_request(key, std::min(ctx->tile_bits, nloc), 0, ctx->max_log_n);
    req.coalesce_bits = std::min(ctx->coalesce_bits, req.tile_bits);
    req.group_bits = 2;
    Plan plan;
    bool ok = make_plan(req, &plan);
    for (int g = req.coalesce_bits - 1; !ok && g >= 0; --g) {
      req.coalesce_bits = g;
      ok = make_plan(req, &plan, 12);
    }
    for (int t = req.tile_bits + 1; !ok && t <= 11; ++t) {
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
    it = ctx->plans.emplace(k, plan).first;
  }
  *out = &it->second;
  return NTT256_OK;
}

const U384& inv_pow2(ntt384_ctx* ctx, int k) {
  if (!ctx->ninv_ok[k]) {
    ctx->ninv[k] = ctx->hf.inv(ctx->hf.to_mont(un_small<W>(1ull << k)));
    ctx->ninv_ok[k] = true;
  }
  return ctx->ninv[k];
}

// sync: bit 0 the first pass waits for every rank's epoch; bit 1 the last pass publishes this rank's epoch
int run_plan(ntt384_ctx* ctx, const Plan& plan, void* d_data, unsigned long long batch_stride, int batch, bool inverse,
             int scale_log_n, cudaStream_t stream, const PeerSpec* peer = nullptr, const IoSpec* io = nullptr, int sync = 0,
             const FuseSpec* fuse = nullptr) {
  PassArgsN<NL> A;
  std::memset(&A, 0, sizeof(A));
  A.f = ctx->fp;
  A.om_hi = ctx->om_hi;
  A.om_lo = ctx->om_lo;
  A.loc = ctx->loc;
  A.hi_bits = ctx->hi_bits;
  A.lo_bits = ctx->lo_bits;
  A.loc_bits = ctx->loc_bits;
  A.inverse = inverse ? 1 : 0;
  std::memcpy(A.c8, ctx->c8[inverse ? 1 : 0], sizeof(A.c8));
  A.batch_stride = batch_stride;
  A.scaled_round = -1;
  const uint4* loc_scaled = nullptr;
  if (scale_log_n >= 0) {
    un_to_limbs<W>(inv_pow2(ctx, scale_log_n), A.scale);
    if (plan.passes.back().nrounds >= 2) {
      auto it = ctx->loc_scaled.find(scale_log_n);
      if (it == ctx->loc_scaled.end()) {
        uint4* tab = nullptr;
        const unsigned long long count = 1ull << ctx->loc_bits;
        CTX_CUDA(cudaMalloc((void**)&tab, count * 4 * NL));
        FeN<NL> c;
        un_to_limbs<W>(inv_pow2(ctx, scale_log_n), c.v);
        table_scale_kernel_n<NL><<<(unsigned)((count + 127) / 128), 128, 0, stream>>>(tab, ctx->loc, count, ctx->fp, c);
        ++ctx->launches;
        CTX_CUDA(cudaGetLastError());
        CTX_CUDA(cudaStreamSynchronize(stream));  // one-off: later calls may use other streams
        it = ctx->loc_scaled.emplace(scale_log_n, tab).first;
      }
      loc_scaled = it->second;
    }
  }
  const bool fz = fuse && (fuse->in_scale || fuse->out_scale || fuse->out_mul);
  const PassTables& tabs = ctx->plan_tables.at(plan.id);
  for (size_t i = 0; i < plan.passes.size(); ++i) {
    const bool last_pass = (i + 1 == plan.passes.size());
    A.pd = plan.passes[i];
    A.thr = tabs.thr[i];
    A.blk = tabs.blk[i];
    A.scaled_round = -1;
    A.loc_scaled = nullptr;
    if (scale_log_n >= 0 && last_pass) {
      if (loc_scaled) {
        A.scaled_round = 0;  // 1/n folded into the first round's twiddles of the last pass
        A.loc_scaled = loc_scaled;
      } else {
        A.pd.scale_out = 1;
      }
    }
    A.peer_bits = 0;
    if (peer && last_pass) {
      for (int r = 0; r < (1 << peer->bits); ++r) A.peer[r] = (unsigned long long)peer->ptrs[r];
      A.peer_bits = peer->bits;
      A.peer_shift = peer->shift;
      A.peer_rank = peer->rank;
    }
    A.in_scale = (fuse && i == 0) ? fuse->in_scale : nullptr;
    A.out_scale = (fuse && last_pass) ? fuse->out_scale : nullptr;
    A.out_mul = (fuse && last_pass) ? fuse->out_mul : nullptr;
    A.sync_mode = 0;
    if (sync && ctx->sync_nranks > 1) {
      if ((sync & 1) && i == 0) A.sync_mode |= 1;
      if ((sync & 2) && last_pass) A.sync_mode |= 2;
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
    const unsigned long long nblk = 1ull << A.pd.nblk_bits;
    if (nblk > 0x7fffffffull || batch > 65535) return NTT256_ERR_ARG;
    dim3 grid((unsigned)nblk, (unsigned)batch, 1);
    const uint4* pin = (i == 0 && io && io->src_first) ? (const uint4*)io->src_first : (const uint4*)d_data;
    uint4* pout = (last_pass && io && io->dst_last) ? (uint4*)io->dst_last : (uint4*)d_data;
    const unsigned threads = 1u << (T - 2);
    const size_t smem = (size_t)(4 * NL) << T;
    {
      // programmatic stream serialization (see grid_dependency_wait in ntt_pass_generic.cuh): the pass becomes resident
      // while its predecessor drains and waits before it touches data
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
      if (fz) cudaLaunchKernelEx(&cfg, ntt384_pass_kernel<true>, pin, (uint4*)pout, A);
      else cudaLaunchKernelEx(&cfg, ntt384_pass_kernel<false>, pin, (uint4*)pout, A);
    }
    ++ctx->launches;
    CTX_CUDA(cudaGetLastError());
  }
  return NTT256_OK;
}

int transform(ntt384_ctx* ctx, void* d_data, int log_n, int batch, bool inverse, cudaStream_t stream,
              const IoSpec* io = nullptr, const FuseSpec* fuse = nullptr) {
  if (!ctx || !d_data || log_n < 0 || log_n > ctx->max_log_n || batch < 1) return NTT256_ERR_ARG;
  CTX_CUDA(cudaSetDevice(ctx->device));
  if (log_n == 0) return NTT256_OK;
  if (log_n < 3) {
    if (io || fuse) return NTT256_ERR_ARG;  // callers keep such sizes on the plain path
    TinyArgsN<NL> A;
    std::memset(&A, 0, sizeof(A));
    A.f = ctx->fp;
    A.log_n = log_n;
    std::memcpy(A.w4, ctx->w4[inverse ? 1 : 0], sizeof(A.w4));
    A.do_scale = inverse ? 1 : 0;
    if (inverse) un_to_limbs<W>(inv_pow2(ctx, log_n), A.scale);
    A.batch = (unsigned long long)batch;
    const int threads = 128;
    ntt_tiny_kernel_n<NL><<<(unsigned)((A.batch + threads - 1) / threads), threads, 0, stream>>>((uint4*)d_data, A);
    ++ctx->launches;
    CTX_CUDA(cudaGetLastError());
    return NTT256_OK;
  }
  const Plan* plan = nullptr;
  int rc = get_plan(ctx, FourStepKey{0, log_n, 0, 0, 0}, &plan);
  if (rc) return rc;
  const int kSlice = 65535;
  for (int b0 = 0; b0 < batch; b0 += kSlice) {
    const int nb = std::min(kSlice, batch - b0);
    const size_t off = ((size_t)b0 << log_n) * 4 * NL;
    IoSpec io_slice;
    if (io) {
      io_slice.src_first = io->src_first ? (const char*)io->src_first + off : nullptr;
      io_slice.dst_last = io->dst_last ? (char*)io->dst_last + off : nullptr;
    }
    FuseSpec fuse_slice;
    if (fuse) {
      fuse_slice = *fuse;
      if (fuse->out_mul) fuse_slice.out_mul = fuse->out_mul + (size_t)(NL / 4) * ((size_t)b0 << log_n);
    }
    rc = run_plan(ctx, *plan, (char*)d_data + off, 1ull << log_n, nb, inverse, inverse ? log_n : -1, stream, nullptr,
                  io ? &io_slice : nullptr, 0, fuse ? &fuse_slice : nullptr);
    if (rc) return rc;
  }
  return NTT256_OK;
}

int fourstep(ntt384_ctx* ctx, int kind, void* d_shard, int log_n, int log_r, int log_p, int rank, cudaStream_t stream,
             const void* const* peer_recv = nullptr) {
  const bool inverse = kind >= 3;
  const bool first_step = (kind == 1 || kind == 3);
  const int scale_log_n = (kind == 4) ? log_n : -1;
  if (!ctx || !d_shard || log_n > ctx->max_log_n || !fourstep_args_ok(log_n, log_r, log_p, rank)) return NTT256_ERR_ARG;
  CTX_CUDA(cudaSetDevice(ctx->device));
  const Plan* plan = nullptr;
  int rc = get_plan(ctx, FourStepKey{kind, log_n, log_r, log_p, first_step ? rank : 0}, &plan);
  if (rc) return rc;
  const bool armed = ctx->sync_nranks == (1 << log_p) && log_p > 0;
  if (peer_recv && log_p > 0) {
    if (log_p > 3) return NTT256_ERR_ARG;
    PeerSpec ps;
    ps.ptrs = peer_recv;
    ps.bits = log_p;
    ps.shift = log_n - 2 * log_p;  // local index bits below the destination-rank field
    ps.rank = rank;
    if (armed) ++ctx->sync_epoch;
    return run_plan(ctx, *plan, d_shard, 0, 1, inverse, scale_log_n, stream, &ps, nullptr, armed ? 2 : 0);
  }
  return run_plan(ctx, *plan, d_shard, 0, 1, inverse, scale_log_n, stream, nullptr, nullptr, (armed && !first_step) ? 1 : 0);
}

int ensure_stage(ntt384_ctx* ctx, size_t bytes) {
  if (bytes > ctx->stage_bytes) {
    if (ctx->d_stage) cudaFree(ctx->d_stage);
    ctx->d_stage = nullptr;
    ctx->stage_bytes = 0;
    CTX_CUDA(cudaMalloc(&ctx->d_stage, bytes));
    ctx->stage_bytes = bytes;
  }
  return NTT256_OK;
}

// Host buffers in, host buffers out.  Pinned (mapped) buffers: the first pass loads straight from host memory and
// the last pass stores straight to it, so the PCIe transfers overlap those passes; otherwise staged copies.
int transform_host(ntt384_ctx* ctx, const void* h_in, void* h_out, int log_n, int batch, bool inverse) {
  if (!ctx || !h_in || !h_out || log_n < 0 || log_n > ctx->max_log_n || batch < 1) return NTT256_ERR_ARG;
  CTX_CUDA(cudaSetDevice(ctx->device));
  const size_t bytes = (((size_t)batch) << log_n) * 4 * NL;
  int rc = ensure_stage(ctx, bytes);
  if (rc) return rc;
  if (log_n >= 3) {
    cudaPointerAttributes ai, ao;
    void *din = nullptr, *dout = nullptr;
    const bool ok = cudaPointerGetAttributes(&ai, h_in) == cudaSuccess && ai.type == cudaMemoryTypeHost &&
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

// shift^j, j < 2^log_n, multiplied out on the device (Montgomery form), cached per (shift, log_n); two kept
int get_coset(ntt384_ctx* ctx, const U384& shift, int log_n, const uint4** out) {
  for (auto& c : ctx->cosets)
    if (c.log_n == log_n && un_cmp(c.shift, shift) == 0) {
      *out = c.direct;
      return NTT256_OK;
    }
  if (ctx->cosets.size() >= 2) {
    cudaDeviceSynchronize();
    cudaFree(ctx->cosets.front().direct);
    ctx->cosets.erase(ctx->cosets.begin());
  }
  ntt384_ctx::Coset c;
  c.shift = shift;
  c.log_n = log_n;
  c.direct = nullptr;
  int rc = build_power_table(ctx, &c.direct, ctx->hf.to_mont(shift), log_n);
  if (rc) return rc;
  ctx->cosets.push_back(c);
  *out = c.direct;
  return NTT256_OK;
}

}  // namespace

extern "C" {

int ntt384_ctx_create(ntt384_ctx** out, int device, const uint32_t p[12], const uint32_t g[12], int max_log_n) {
  if (!out || !p || !g || max_log_n < 1 || max_log_n > 30) return NTT256_ERR_ARG;
  *out = nullptr;
  U384 P = un_from_limbs<W>(p), G = un_from_limbs<W>(g);
  if (!(P.w[0] & 1) || un_bitlen(P) > 32 * NL - 1 || un_bitlen(P) < 2) return NTT256_ERR_FIELD;
  if (un_is_zero(G) || un_cmp(G, P) >= 0) return NTT256_ERR_FIELD;
  U384 pm1;
  un_sub(pm1, P, un_small<W>(1));
  for (int i = 0; i < max_log_n; ++i)
    if (un_bit(pm1, i)) return NTT256_ERR_FIELD;
  ntt384_ctx* ctx = new (std::nothrow) ntt384_ctx();
  if (!ctx) return NTT256_ERR_NOMEM;
  ctx->device = device;
  ctx->max_log_n = max_log_n;
  ctx->hf.init(P);
  std::memset(&ctx->fp, 0, sizeof(ctx->fp));
  un_to_limbs<W>(P, ctx->fp.p);
  U384 P2;
  un_add(P2, P, P);
  un_to_limbs<W>(P2, ctx->fp.p2);
  ctx->fp.np0 = (u32)ctx->hf.np0;
  // w_N = g^((p-1)/N), must have exact order N
  U384 gm = ctx->hf.to_mont(G);
  ctx->wN = ctx->hf.pow(gm, un_shr(pm1, max_log_n));
  ctx->wN_inv = ctx->hf.inv(ctx->wN);
  {
    U384 h = pow2k(ctx->hf, ctx->wN, max_log_n - 1);
    U384 m1 = ctx->hf.sub(un_small<W>(0), ctx->hf.one);
    if (un_cmp(h, m1) != 0) {
      delete ctx;
      return NTT256_ERR_FIELD;
    }
  }
  ctx->hi_bits = std::min(max_log_n, kHiBits384);
  ctx->lo_bits = max_log_n - ctx->hi_bits;
  ctx->loc_bits = std::min(max_log_n, kLocBits384);
  std::memset(ctx->c8, 0, sizeof(ctx->c8));
  std::memset(ctx->w4, 0, sizeof(ctx->w4));
  if (max_log_n >= 3) {
    U384 w8 = pow2k(ctx->hf, ctx->wN, max_log_n - 3), w8i = pow2k(ctx->hf, ctx->wN_inv, max_log_n - 3);
    U384 a = w8, b = w8i;
    for (int k = 0; k < 3; ++k) {
      un_to_limbs<W>(a, ctx->c8[0][k]);
      un_to_limbs<W>(b, ctx->c8[1][k]);
      a = ctx->hf.mul(a, w8);
      b = ctx->hf.mul(b, w8i);
    }
  }
  if (max_log_n >= 2) {
    un_to_limbs<W>(pow2k(ctx->hf, ctx->wN, max_log_n - 2), ctx->w4[0]);
    un_to_limbs<W>(pow2k(ctx->hf, ctx->wN_inv, max_log_n - 2), ctx->w4[1]);
    if (max_log_n < 3) {
      std::memcpy(ctx->c8[0][1], ctx->w4[0], 4 * NL);
      std::memcpy(ctx->c8[1][1], ctx->w4[1], 4 * NL);
    }
  }
  auto fail = [&](int code) {
    ntt384_ctx_destroy(ctx);
    return code;
  };
  if (cudaSetDevice(device) != cudaSuccess) return fail(NTT256_ERR_CUDA);
  if (cudaFuncSetAttribute(ntt384_pass_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (4 * NL) << 11) != cudaSuccess ||
      cudaFuncSetAttribute(ntt384_pass_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (4 * NL) << 11) != cudaSuccess)
    return fail(NTT256_ERR_CUDA);
  int rc = build_power_table(ctx, &ctx->om_hi, pow2k(ctx->hf, ctx->wN, ctx->lo_bits), ctx->hi_bits);
  if (rc) return fail(rc);
  rc = build_power_table(ctx, &ctx->om_lo, ctx->wN, ctx->lo_bits);
  if (rc) return fail(rc);
  rc = build_power_table(ctx, &ctx->loc, pow2k(ctx->hf, ctx->wN, max_log_n - ctx->loc_bits), ctx->loc_bits);
  if (rc) return fail(rc);
  *out = ctx;
  return NTT256_OK;
}

void ntt384_ctx_destroy(ntt384_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  if (ctx->om_hi) cudaFree(ctx->om_hi);
  if (ctx->om_lo) cudaFree(ctx->om_lo);
  if (ctx->loc) cudaFree(ctx->loc);
  for (auto& kv : ctx->loc_scaled) cudaFree(kv.second);
  for (auto& c : ctx->cosets) cudaFree(c.direct);
  for (auto& kv : ctx->plan_tables) free_pass_tables(&kv.second);
  if (ctx->d_stage) cudaFree(ctx->d_stage);
  if (ctx->sync_counter) cudaFree(ctx->sync_counter);
  delete ctx;
}

int ntt384_ctx_set_tuning(ntt384_ctx* ctx, int tile_bits, int coalesce_bits) {
  if (!ctx || tile_bits < 3 || tile_bits > 11 || coalesce_bits < 0 || coalesce_bits > 6 || coalesce_bits > tile_bits)
    return NTT256_ERR_ARG;
  ctx->tile_bits = tile_bits;
  ctx->coalesce_bits = coalesce_bits;
  cudaSetDevice(ctx->device);
  cudaDeviceSynchronize();
  ctx->plans.clear();  // loc_scaled tables depend on log_n only and stay valid
  for (auto& kv : ctx->plan_tables) free_pass_tables(&kv.second);
  ctx->plan_tables.clear();
  return NTT256_OK;
}

int ntt384_forward_batch(ntt384_ctx* ctx, void* d_data, int log_n, int batch, ntt256_stream stream) {
  return transform(ctx, d_data, log_n, batch, false, (cudaStream_t)stream);
}
int ntt384_inverse_batch(ntt384_ctx* ctx, void* d_data, int log_n, int batch, ntt256_stream stream) {
  return transform(ctx, d_data, log_n, batch, true, (cudaStream_t)stream);
}
int ntt384_forward(ntt384_ctx* ctx, void* d_data, int log_n, ntt256_stream stream) {
  return transform(ctx, d_data, log_n, 1, false, (cudaStream_t)stream);
}
int ntt384_inverse(ntt384_ctx* ctx, void* d_data, int log_n, ntt256_stream stream) {
  return transform(ctx, d_data, log_n, 1, true, (cudaStream_t)stream);
}

int ntt384_forward_host(ntt384_ctx* ctx, const void* h_in, void* h_out, int log_n, int batch) {
  return transform_host(ctx, h_in, h_out, log_n, batch, false);
}
int ntt384_inverse_host(ntt384_ctx* ctx, const void* h_in, void* h_out, int log_n, int batch) {
  return transform_host(ctx, h_in, h_out, log_n, batch, true);
}

int ntt384_coset_forward(ntt384_ctx* ctx, void* d_data, int log_n, int batch, const uint32_t shift[12], ntt256_stream stream) {
  if (!ctx || !d_data || !shift || log_n < 3 || log_n > ctx->max_log_n || batch < 1) return NTT256_ERR_ARG;
  const U384 s = un_from_limbs<W>(shift);
  if (un_is_zero(s) || un_cmp(s, ctx->hf.p) >= 0) return NTT256_ERR_ARG;
  CTX_CUDA(cudaSetDevice(ctx->device));
  FuseSpec fz;
  int rc = get_coset(ctx, s, log_n, &fz.in_scale);  // shift^j multiplies the elements as the first pass loads them
  if (rc) return rc;
  return transform(ctx, d_data, log_n, batch, false, (cudaStream_t)stream, nullptr, &fz);
}
int ntt384_coset_inverse(ntt384_ctx* ctx, void* d_data, int log_n, int batch, const uint32_t shift[12], ntt256_stream stream) {
  if (!ctx || !d_data || !shift || log_n < 3 || log_n > ctx->max_log_n || batch < 1) return NTT256_ERR_ARG;
  const U384 s = un_from_limbs<W>(shift);
  if (un_is_zero(s) || un_cmp(s, ctx->hf.p) >= 0) return NTT256_ERR_ARG;
  CTX_CUDA(cudaSetDevice(ctx->device));
  const U384 sinv = ctx->hf.from_mont(ctx->hf.inv(ctx->hf.to_mont(s)));
  FuseSpec fz;
  int rc = get_coset(ctx, sinv, log_n, &fz.out_scale);  // shift^-k multiplies the outputs before the last pass stores them
  if (rc) return rc;
  return transform(ctx, d_data, log_n, batch, true, (cudaStream_t)stream, nullptr, &fz);
}

static int transform_mul(ntt384_ctx* ctx, void* d_data, const void* d_mul_mont, int log_n, int batch, bool inverse,
                         cudaStream_t stream) {
  if (!ctx || !d_data || !d_mul_mont || log_n < 3 || log_n > ctx->max_log_n || batch < 1) return NTT256_ERR_ARG;
  CTX_CUDA(cudaSetDevice(ctx->device));
  FuseSpec fz;
  fz.out_mul = (const uint4*)d_mul_mont;
  return transform(ctx, d_data, log_n, batch, inverse, stream, nullptr, &fz);
}
int ntt384_forward_mul(ntt384_ctx* ctx, void* d_data, const void* d_mul_mont, int log_n, int batch, ntt256_stream stream) {
  return transform_mul(ctx, d_data, d_mul_mont, log_n, batch, false, (cudaStream_t)stream);
}
int ntt384_inverse_mul(ntt384_ctx* ctx, void* d_data, const void* d_mul_mont, int log_n, int batch, ntt256_stream stream) {
  return transform_mul(ctx, d_data, d_mul_mont, log_n, batch, true, (cudaStream_t)stream);
}

int ntt384_fourstep_step1(ntt384_ctx* ctx, void* d_shard, int log_n, int log_r, int log_p, int rank, ntt256_stream stream) {
  return fourstep(ctx, 1, d_shard, log_n, log_r, log_p, rank, (cudaStream_t)stream);
}
int ntt384_fourstep_step1_peer(ntt384_ctx* ctx, void* d_shard, const void* const* peer_recv, int log_n, int log_r,
                               int log_p, int rank, ntt256_stream stream) {
  if (!peer_recv) return NTT256_ERR_ARG;
  return fourstep(ctx, 1, d_shard, log_n, log_r, log_p, rank, (cudaStream_t)stream, peer_recv);
}
int ntt384_fourstep_step2(ntt384_ctx* ctx, void* d_shard, int log_n, int log_r, int log_p, int rank, ntt256_stream stream) {
  return fourstep(ctx, 2, d_shard, log_n, log_r, log_p, rank, (cudaStream_t)stream);
}
int ntt384_fourstep_inverse_step1(ntt384_ctx* ctx, void* d_shard, int log_n, int log_r, int log_p, int rank,
                                  ntt256_stream stream) {
  return fourstep(ctx, 3, d_shard, log_n, log_r, log_p, rank, (cudaStream_t)stream);
}
int ntt384_fourstep_inverse_step1_peer(ntt384_ctx* ctx, void* d_shard, const void* const* peer_recv, int log_n,
                                       int log_r, int log_p, int rank, ntt256_stream stream) {
  if (!peer_recv) return NTT256_ERR_ARG;
  return fourstep(ctx, 3, d_shard, log_n, log_r, log_p, rank, (cudaStream_t)stream, peer_recv);
}
int ntt384_fourstep_inverse_step2(ntt384_ctx* ctx, void* d_shard, int log_n, int log_r, int log_p, int rank,
                                  ntt256_stream stream) {
  return fourstep(ctx, 4, d_shard, log_n, log_r, log_p, rank, (cudaStream_t)stream);
}

int ntt384_ctx_set_peer_sync(ntt384_ctx* ctx, void* const* flag_bufs, int nranks, int rank, double timeout_s) {
  return set_peer_sync_impl(ctx, flag_bufs, nranks, rank, timeout_s);
}

int ntt384_vec_mul(ntt384_ctx* ctx, void* d_out, const void* d_a, const void* d_b, unsigned long long n, int mont_form,
                   ntt256_stream stream) {
  if (!ctx || !d_out || !d_a || !d_b) return NTT256_ERR_ARG;
  if (n == 0) return NTT256_OK;
  CTX_CUDA(cudaSetDevice(ctx->device));
  FeN<NL> r2;
  un_to_limbs<W>(ctx->hf.r2, r2.v);
  const int threads = 128;
  vec_mul_kernel_n<NL><<<(unsigned)((n + threads - 1) / threads), threads, 0, (cudaStream_t)stream>>>(
      (uint4*)d_out, (const uint4*)d_a, (const uint4*)d_b, n, mont_form, ctx->fp, r2);
  ++ctx->launches;
  CTX_CUDA(cudaGetLastError());
  return NTT256_OK;
}

unsigned long long ntt384_launch_count(const ntt384_ctx* ctx) { return ctx ? ctx->launches : 0; }
const char* ntt384_last_cuda_error(const ntt384_ctx* ctx) { return ctx ? ctx->last_err.c_str() : ""; }

size_t ntt384_plan_describe_fourstep(int step, int log_n, int log_r, int log_p, int rank, int max_log_n, char* buf,
                                     size_t buf_len) {
  std::string s = "{\"error\": \"no plan\"}";
  if (step >= 1 && step <= 4 && fourstep_args_ok(log_n, log_r, log_p, rank) && log_n <= max_log_n) {
    const bool first = (step == 1 || step == 3);
    const int nloc = log_n - log_p;
    // the request get_plan() makes for this piece with the default tuning
    PlanRequest req = fourstep_request(FourStepKey{step, log_n, log_r, log_p, first ? rank : 0}, std::min(kTileBits384, nloc), 0,
                                       max_log_n);
    req.coalesce_bits = std::min(kCoalesce384, req.tile_bits);
    req.group_bits = 2;
    Plan plan;
    bool ok = make_plan(req, &plan);
    for (int g = req.coalesce_bits - 1; !ok && g >= 0; --g) {
      req.coalesce_bits = g;
      ok = make_plan(req, &plan, 12);
    }
    for (int t = req.tile_bits + 1; !ok && t <= 11; ++t) {
      req.tile_bits = t;
      ok = make_plan(req, &plan, 12);
    }
    if (ok) s = plan.describe();
  }
  if (buf && buf_len) {
    size_t m = std::min(buf_len - 1, s.size());
    std::memcpy(buf, s.data(), m);
    buf[m] = 0;
  }
  return s.size() + 1;
}

size_t ntt384_plan_describe(int log_n, int max_log_n, char* buf, size_t buf_len) {
  PlanRequest req;
  req.log_n = log_n;
  req.tile_bits = std::min(kTileBits384, log_n);
  req.coalesce_bits = std::min(kCoalesce384, req.tile_bits);
  req.max_log_n = max_log_n;
  req.group_bits = 2;
  Plan plan;
  bool ok = make_plan(req, &plan);
  for (int g = req.coalesce_bits - 1; !ok && g >= 0; --g) {
    req.coalesce_bits = g;
    ok = make_plan(req, &plan, 12);
  }
  std::string s = ok ? plan.describe() : std::string("{\"error\": \"no plan\"}");
  if (buf && buf_len) {
    size_t m = std::min(buf_len - 1, s.size());
    std::memcpy(buf, s.data(), m);
    buf[m] = 0;
  }
  return s.size() + 1;
}

}  // extern "C"
