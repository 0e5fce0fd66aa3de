/* libntt256 -- B200-native (sm_100a) forward/inverse NTT over 256-bit prime fields.
 *
 * Drop-in boundary for the 256-bit NTT path of tie-pilot-qxw/NTT.  The reference has no
 * library API: its boundary is the C++ host drivers called from main(),
 *
 *   template<uint tpi,uint bits> void NTT_GZKP(cgbn_mem_t<bits> data[], uint len, uint2 reverse[],
 *       uint reverse_len, cgbn_mem_t<bits> prime, cgbn_mem_t<bits> omega, uint B, uint G)
 *                                                     (/root/reference/src/big-num.cu:260-353)
 *   template<uint tpi,uint bits> void NTT_GPU_Naive(...)              (big-num.cu:106-170)
 *
 * whose conventions this header keeps:
 *   - `data` is a DEVICE pointer owned by the caller, n = 2^log_n elements, transformed IN PLACE,
 *     natural order in and out;
 *   - an element is cgbn_mem_t<256> (include/cgbn/cgbn_cuda.h:51-55): 8 little-endian u32 limbs,
 *     32 bytes, canonical residue in [0, p), NOT in Montgomery form (on a little-endian host the
 *     same bytes are the 4 x u64 limb layout of arkworks / blst / bellman field elements; data in
 *     Montgomery form can be passed as is -- the transform is linear, see ntt256_vec_mul for the
 *     one operation where the form matters).  CONTRACT: every input element must be a canonical
 *     residue (< p), as cgbn_mem_t data in the reference is; nothing re-reduces the caller's data,
 *     and values >= p give unspecified (not merely unreduced) results;
 *   - `prime` and `omega` are 8-limb host values; `omega` is the multiplicative GENERATOR g and
 *     the callee derives w_n = g^((p-1)/n) (big-num.cu:128-134, 292-296);
 *   - X[k] = sum_j x[j] * w_n^(jk) mod p (CPU oracle NTT(), big-num.cu:37-55).
 * What it drops: the bit-reversal pair list (`reverse`, `reverse_len`: no permutation pass exists
 * here), the tuning knobs (B, G: chosen by the plan), the per-call twiddle table build
 * (big-num.cu:298-309: tables live in the context), printing and asserts (status codes instead).
 * What it adds: inverse transform (the reference's commented intent, big-num.cu:470-477),
 * batches, streams, and the multi-GPU four-step pieces (SURVEY.md par. 8e).
 *
 * Threading: a context is bound to one device and is not re-entrant.  Calls are asynchronous
 * on the given stream unless the name ends in _sync or takes host buffers -- with one exception:
 * the FIRST call of a (size, batch shape, direction) builds what it needs lazily (the pass plan and
 * its index tables, the 1/n-scaled round table of an inverse, a streamed first-pass twiddle array)
 * and synchronises once while doing so; such a call must not be made inside a CUDA graph capture.
 * Later calls of the same shape only launch kernels and are capturable.
 * There is NO CPU fallback: every entry point fails with NTT256_ERR_CUDA if no sm_100 device
 * can run it.
 */
#ifndef NTT256_H_
#define NTT256_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ntt256_ctx ntt256_ctx;
typedef void* ntt256_stream;  /* cudaStream_t; NULL = default stream */

enum {
  NTT256_OK = 0,
  NTT256_ERR_ARG = 1,         /* null pointer, log_n out of range, bad field id */
  NTT256_ERR_FIELD = 2,       /* p even, p >= 2^255, or 2^max_log_n does not divide p-1 */
  NTT256_ERR_CUDA = 3,        /* CUDA runtime error; ntt256_last_cuda_error() has the text */
  NTT256_ERR_NOMEM = 4,
  NTT256_ERR_PLAN = 5         /* no pass plan found (internal) */
};

enum { NTT256_FIELD_BN254_FR = 0, NTT256_FIELD_BLS12_381_FR = 1 };

/* Context for an arbitrary odd modulus p < 2^255 and generator g (both 8 LE u32 limbs), able to
 * transform up to 2^max_log_n points.  Replaces the `prime`/`omega` by-value arguments of
 * NTT_GZKP (big-num.cu:260-261) and the host twiddle precompute inside it (big-num.cu:278-309):
 * tables are built once, on the device. */
int ntt256_ctx_create(ntt256_ctx** out, int device, const uint32_t p[8], const uint32_t g[8], int max_log_n);
/* Same for the two named fields of BASELINE.json (BN254 Fr with g = 5, BLS12-381 Fr with g = 7). */
int ntt256_ctx_create_named(ntt256_ctx** out, int device, int field, int max_log_n);
void ntt256_ctx_destroy(ntt256_ctx* ctx);

/* Plan tuning (optional): tile_bits in [3,12] (shared-memory tile = 2^tile_bits elements),
 * coalesce_bits in [0,6] (every global access run is >= 2^coalesce_bits elements).  Cached
 * plans are discarded. */
int ntt256_ctx_set_tuning(ntt256_ctx* ctx, int tile_bits, int coalesce_bits);

/* The first pass of a transform longer than 2^18 points needs twiddles w_n^e with a 2^log_n-wide
 * exponent range: either two table lookups and one extra product per element, or one coalesced
 * read from an n-entry array laid out like the pass's output (32*n bytes per (size, direction),
 * built on first use, kept in the context; 64*n bytes with Shoup pairs).  Arrays up to max_bytes are
 * built automatically (default: 1/10 of the device's memory, i.e. every n <= 2^28 on a 180 GB B200;
 * when the allocation fails the two-lookup path is used); 0 selects the two-lookup path always. */
int ntt256_ctx_set_stream_twiddle_limit(ntt256_ctx* ctx, size_t max_bytes);
/* All streamed arrays of a context together stay below total_bytes (default: 1/5 of the device's memory):
 * before a new one is built the least recently used ones are freed (cudaFree waits for the device, so a
 * launch that still reads one is never cut off); a transform whose array was dropped rebuilds it at its
 * next call.  ntt256_ctx_table_bytes returns the device memory the context holds in twiddle / coset
 * tables right now and, through *streamed_bytes (may be NULL), the streamed arrays' share of it. */
int ntt256_ctx_set_stream_twiddle_budget(ntt256_ctx* ctx, size_t total_bytes);
size_t ntt256_ctx_table_bytes(const ntt256_ctx* ctx, size_t* streamed_bytes);
/* The pass kernels of a transform are launched with programmatic stream serialization (default on): pass i+1
 * becomes resident and does its index arithmetic while pass i drains its last wave, and touches no data before
 * pass i has completed.  enable = 0 restores plain stream-ordered launches (same results). */
int ntt256_ctx_set_dependent_launch(ntt256_ctx* ctx, int enable);
/* BN254 Fr contexts run the throughput kernel in an instance whose field constants are compile-time immediates
 * (default on; same bits).  enable = 0 keeps them on the generic instance (testing). */
int ntt256_ctx_set_constant_field_kernel(ntt256_ctx* ctx, int enable);
/* Calls with batch * 2^log_n <= max_elements (default 2^17) are latency-bound and run on the
 * four-elements-per-thread kernel, those of at most half that on the two-elements-per-thread
 * kernel (more, shorter threads); 0 keeps every call on the throughput kernel. */
int ntt256_ctx_set_small_transform_limit(ntt256_ctx* ctx, long long max_elements);
/* Tuning/testing: force the pass kernel that keeps 2^log2_elements (1, 2 or 3) elements per thread
 * for ntt256_forward/inverse[_batch]; 0 restores the automatic choice.  Tiles are limited to 2^11
 * elements below 3. */
int ntt256_ctx_set_elements_per_thread(ntt256_ctx* ctx, int log2_elements);
/* Four-elements-per-thread passes run on a kernel instance compiled for three resident CTAs per SM (<= 85
 * registers, 24 warps per SM) when the CTA has <= 256 threads; 0 selects the 128-register instance. */
int ntt256_ctx_set_dense_pass4(ntt256_ctx* ctx, int enable);
/* Twiddle representation of the pass kernels.  1 (the default whenever p < 2^254, e.g. BN254 Fr): tables
 * hold 64-byte pairs (w, floor(w 2^256 / p)) and a twiddle product is a Shoup/Barrett multiplication
 * (99 wide + 16 low 32-bit multiplies, csrc/field_shoup.cuh); 0 (the only choice for 255-bit moduli such as
 * BLS12-381 Fr): 32-byte Montgomery-form entries and the Montgomery product (128 + 8).  Both give the same
 * bits; switching rebuilds the tables (synchronises the device).  NTT256_ERR_FIELD if 1 is asked of a field
 * with p >= 2^254. */
int ntt256_ctx_set_twiddle_mode(ntt256_ctx* ctx, int shoup);
int ntt256_ctx_get_twiddle_mode(const ntt256_ctx* ctx);

/* Forward NTT of 2^log_n points, in place on device memory, natural order in and out.
 * == NTT_GZKP<8,256>(data, 1<<log_n, ..., prime, omega, B, G)  (big-num.cu:458). */
int ntt256_forward(ntt256_ctx* ctx, void* d_data, int log_n, ntt256_stream stream);
/* Inverse NTT including the 1/n scaling (big-num.cu:470-477: NTT with inv(root), times inv(length)). */
int ntt256_inverse(ntt256_ctx* ctx, void* d_data, int log_n, ntt256_stream stream);
/* `batch` independent transforms stored back to back (transform b at d_data + b * 32 * 2^log_n). */
int ntt256_forward_batch(ntt256_ctx* ctx, void* d_data, int log_n, int batch, ntt256_stream stream);
int ntt256_inverse_batch(ntt256_ctx* ctx, void* d_data, int log_n, int batch, ntt256_stream stream);

/* Reference-like synchronous convenience: HOST buffer in, HOST buffer out (may alias), the
 * library stages through its own device buffer: H2D, transform, D2H, synchronise.  Batches are
 * streamed in chunks (uplink, kernels and downlink overlap; pass pinned memory for that).
 * Mirrors main() of big-num.cu:424-460 (cudaMemcpy H2D, NTT_GZKP, cudaMemcpy D2H). */
int ntt256_forward_host(ntt256_ctx* ctx, const void* h_in, void* h_out, int log_n, int batch);
int ntt256_inverse_host(ntt256_ctx* ctx, const void* h_in, void* h_out, int log_n, int batch);

/* ---- companions of the transform (SURVEY.md par. 8f: the prover steps either side of an NTT;
 * the reference has none of them -- its kernels were lifted out of that context,
 * /root/reference/src/GZKP-NTT.cu:321-386).  All on device memory, asynchronous on `stream`.
 *
 * Element-wise product of two arrays of n elements.  mont_form = 0: canonical residues in and out
 * (out = a*b mod p); mont_form = 1: Montgomery-form data in and out (out = a*b*2^-256 mod p, i.e.
 * the Montgomery product, one multiplication instead of two).  d_out may alias d_a or d_b. */
int ntt256_vec_mul(ntt256_ctx* ctx, void* d_out, const void* d_a, const void* d_b, unsigned long long n, int mont_form,
                   ntt256_stream stream);
/* In-place conversion of n elements to / from Montgomery form (x*2^256 mod p).  The transforms
 * themselves are linear and therefore work on Montgomery-form data unchanged. */
int ntt256_to_montgomery(ntt256_ctx* ctx, void* d_data, unsigned long long n, ntt256_stream stream);
int ntt256_from_montgomery(ntt256_ctx* ctx, void* d_data, unsigned long long n, ntt256_stream stream);
/* Coset transforms: forward X[k] = sum_j (shift^j * x[j]) * w_n^(jk); inverse is its exact
 * inverse (x[j] = shift^-j * INTT(X)[j]).  shift: 8 LE u32 limbs, 0 < shift < p. */
int ntt256_coset_forward(ntt256_ctx* ctx, void* d_data, int log_n, int batch, const uint32_t shift[8],
                         ntt256_stream stream);
int ntt256_coset_inverse(ntt256_ctx* ctx, void* d_data, int log_n, int batch, const uint32_t shift[8],
                         ntt256_stream stream);
/* Transform, then multiply output k by d_mul_mont[k] -- a caller-owned device vector of batch * 2^log_n
 * elements in MONTGOMERY form (ntt256_to_montgomery), so that the result is the plain product
 * NTT(x)[k] * b[k] (resp. iNTT(x)[k] * b[k]): the pointwise-product step that follows a transform in a
 * polynomial multiplication, fused into the stores of the last pass (no extra trip through HBM). */
int ntt256_forward_mul(ntt256_ctx* ctx, void* d_data, const void* d_mul_mont, int log_n, int batch, ntt256_stream stream);
int ntt256_inverse_mul(ntt256_ctx* ctx, void* d_data, const void* d_mul_mont, int log_n, int batch, ntt256_stream stream);
/* Where the companions run.  1: the coset shift multiplies elements inside the first pass (forward) / last pass
 * (inverse) and the pointwise product inside the last pass; 0: separate element-wise kernels before / after the
 * passes; -1 (default): fused for calls of at most 2^17 elements (saves a launch), separate above (one more
 * product per element costs the same wherever it runs on this multiplier-bound path, and the element-wise
 * kernel runs at full occupancy).  Same bits either way. */
int ntt256_ctx_set_fused_companions(ntt256_ctx* ctx, int enable);

/* Multi-GPU four-step pieces (SURVEY.md par. 8e).  One process per GPU; the all-to-all between
 * the two calls is done by the caller (torch.distributed / NCCL, see ntt_b200/dist.py).
 *   n = 2^log_n = R*C, R = 2^log_r rows, C = n/R columns, P = 2^log_p ranks.
 *   Input shard of rank r: x[j1*C + j2] for j2 in [r*C/P, (r+1)*C/P), packed row-major R x (C/P).
 *   step1: R-point column NTTs + twiddle w_n^(j2*k1), in place.
 *   exchange: rank r sends rows k1 in [s*R/P, (s+1)*R/P) to rank s (contiguous chunks).
 *   step2: on the received buffer laid out [src rank][k1 local][j2 local], C-point row NTTs.
 *          Output shard of rank r: X[k1 + R*k2] for k1 in block r, stored as P blocks of
 *          (R/P) x (C/P): element index = ((k2 mod P)*(R/P) + k1_local)*(C/P) + (k2 div P)
 *          (each index field is transformed where it lies, which keeps step 2 to one pass for
 *          C <= 4096; ntt_b200.dist.gather_output() turns the shards into natural order). */
int ntt256_fourstep_step1(ntt256_ctx* ctx, void* d_shard, int log_n, int log_r, int log_p, int rank,
                          ntt256_stream stream);
int ntt256_fourstep_step2(ntt256_ctx* ctx, void* d_shard, int log_n, int log_r, int log_p, int rank,
                          ntt256_stream stream);
/* step 1 with the exchange FUSED into its last pass: instead of storing the tile back in place,
 * the butterfly kernel stores every output element straight into the receive buffer of its
 * destination rank over NVLink.  peer_recv[s] = rank s's receive buffer (2^(log_n-log_p) elements)
 * as mapped into THIS process (CUDA IPC / peer access; peer_recv[rank] is the local one).
 * The caller only has to barrier before step 2 reads the receive buffer.  log_p <= 3. */
int ntt256_fourstep_step1_peer(ntt256_ctx* ctx, void* d_shard, const void* const* peer_recv, int log_n, int log_r,
                               int log_p, int rank, ntt256_stream stream);

/* Distributed INVERSE (includes the 1/n): consumes the output shards of the forward four-step
 * (same log_r) and returns its input shards, i.e. it undoes steps 2, exchange, 1 with the same
 * three-phase structure (local passes, exchange of contiguous chunks / fused peer stores, local
 * passes); forward followed by inverse is the identity on every rank's shard. */
int ntt256_fourstep_inverse_step1(ntt256_ctx* ctx, void* d_shard, int log_n, int log_r, int log_p, int rank,
                                  ntt256_stream stream);
int ntt256_fourstep_inverse_step1_peer(ntt256_ctx* ctx, void* d_shard, const void* const* peer_recv, int log_n,
                                       int log_r, int log_p, int rank, ntt256_stream stream);
int ntt256_fourstep_inverse_step2(ntt256_ctx* ctx, void* d_shard, int log_n, int log_r, int log_p, int rank,
                                  ntt256_stream stream);

/* ---- single-process multi-GPU driver (the C-ABI list of SURVEY.md par. 8b): one host thread
 * drives ngpus (power of two, <= 8) peer-connected devices; no NCCL, no second process.  Shards
 * use the four-step layouts above (default log_r = ceil(log_n/2)); the exchange is the fused one
 * (phase 1's last kernel stores into the peers' receive buffers), phases are ordered with
 * cross-device events, and phase 2's last kernel writes the result back into d_shards[i].
 * Calls are asynchronous on internal streams: ntt256_mgpu_synchronize() waits for them.
 * devices == NULL selects devices 0 .. ngpus-1. */
typedef struct ntt256_mgpu ntt256_mgpu;
int ntt256_mgpu_create(ntt256_mgpu** out, int ngpus, const int* devices, int field, int max_log_n);
int ntt256_mgpu_forward(ntt256_mgpu* m, void* const* d_shards, int log_n);
int ntt256_mgpu_inverse(ntt256_mgpu* m, void* const* d_shards, int log_n);  /* undoes ntt256_mgpu_forward */
int ntt256_mgpu_forward_batch(ntt256_mgpu* m, void* const* d_shards, int log_n, int batch_per_gpu);
int ntt256_mgpu_synchronize(ntt256_mgpu* m);
void ntt256_mgpu_destroy(ntt256_mgpu* m);

/* Exchange buffers for the fused step: cudaMalloc'ed by the library so that a CUDA IPC handle
 * (64 bytes) describes exactly this buffer; the other ranks open it with THEIR device current. */
int ntt256_ipc_alloc(int device, size_t bytes, void** d_ptr, unsigned char handle[64]);
int ntt256_ipc_open(int device, const unsigned char handle[64], void** d_ptr);
int ntt256_ipc_close(int device, void* d_ptr);
int ntt256_ipc_free(int device, void* d_ptr);

/* Stream-ordered barrier between the ranks of a fused four-step transform, through peer memory
 * (no host involvement, no NCCL launch): flag_bufs[r] is rank r's flag buffer as mapped in THIS
 * process (an ntt256_ipc_alloc'ed buffer of >= 4 KiB, zero-filled from byte 256 on, never written
 * by anything else; flag_bufs[rank] is the local one).  One tiny kernel on `stream` publishes
 * `epoch` in every peer's buffer (after a system-scope fence, so the peer stores of the kernels
 * before it on the stream are visible first) and waits until every peer has published an epoch
 * >= `epoch` here.  Epochs must increase by one per call, identically on every rank.  A peer that
 * does not arrive within the timeout (default 60 s, ntt256_set_peer_barrier_timeout) makes the waiter
 * record the epoch in its flag buffer and give up -- no trap, the context and the IPC mappings stay
 * usable; ntt256_peer_sync_status reports it after a synchronisation.  nranks <= 8. */
int ntt256_peer_barrier(int device, void* const* flag_bufs, int nranks, int rank, unsigned long long epoch,
                        ntt256_stream stream);
int ntt256_set_peer_barrier_timeout(double seconds);
/* NTT256_OK if no waiter of this rank's flag buffer ever timed out; else NTT256_ERR_CUDA and the epoch it
 * was waiting for in *timed_out_epoch (the results of that transform are invalid).  Synchronous. */
int ntt256_peer_sync_status(int device, const void* flag_buf, unsigned long long* timed_out_epoch);
/* The same synchronisation folded INTO the four-step kernels, no barrier launch at all: once a context is
 * armed with every rank's flag buffer, the last pass of ntt256_fourstep[_inverse]_step1_peer ends by
 * publishing a new epoch from its last CTA (after system-scope fences behind the peer stores), and the first
 * pass of the following ntt256_fourstep[_inverse]_step2 begins with every CTA waiting for all ranks' epochs.
 * Calls must alternate step1_peer / step2 identically on every rank.  flag_bufs == NULL or nranks <= 1
 * disarms.  timeout_s <= 0 keeps the current timeout (default 60 s). */
int ntt256_ctx_set_peer_sync(ntt256_ctx* ctx, void* const* flag_bufs, int nranks, int rank, double timeout_s);
/* Lets kernels running on `device` store into memory of `peer_device` (NVLink peer access);
 * needed once per peer before ntt256_fourstep_step1_peer.  Already-enabled is not an error. */
int ntt256_enable_peer_access(int device, int peer_device);

/* Introspection (host only, no CUDA call): JSON description of the pass plan that
 * ntt256_forward would run for 2^log_n points.  Returns the number of bytes needed
 * (including the NUL); writes at most buf_len bytes. */
size_t ntt256_plan_describe(int log_n, int max_log_n, int tile_bits, int coalesce_bits, char* buf, size_t buf_len);
/* group_bits = 3: the throughput kernel (eight elements per thread); 2: the four-elements-per-thread
 * kernel the library picks for small, latency-bound transforms. */
size_t ntt256_plan_describe_ex(int log_n, int max_log_n, int tile_bits, int coalesce_bits, int group_bits, char* buf,
                               size_t buf_len);
/* Same for one local step of the four-step decomposition on `rank`: 1, 2 = forward steps,
 * 3, 4 = steps 1 and 2 of the distributed inverse. */
size_t ntt256_plan_describe_fourstep(int step, int log_n, int log_r, int log_p, int rank, int max_log_n, int tile_bits,
                                     int coalesce_bits, char* buf, size_t buf_len);
/* Number of kernel launches issued by this context so far (bench.py's gpu_launches). */
unsigned long long ntt256_launch_count(const ntt256_ctx* ctx);

const char* ntt256_strerror(int status);
const char* ntt256_last_cuda_error(const ntt256_ctx* ctx);
const char* ntt256_version(void);

#ifdef __cplusplus
}
#endif
#endif /* NTT256_H_ */
