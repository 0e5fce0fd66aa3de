/* libntt256 -- 64-bit companion API: forward / inverse NTT over prime fields held in 64-bit containers
 * (one little-endian u64 per element, canonical residues in [0, p)).
 *
 * The reference's drivers and kernels are templates over the container width,
 *   template<uint tpi, uint bits> void NTT_GZKP(cgbn_mem_t<bits> data[], uint len, ...)
 *                                                (/root/reference/src/big-num.cu:57, 67, 172, 260)
 * and its CPU transform runs in 64-bit containers,
 *   void NTT(long long data[], long long reverse[], long long len, long long omega)
 *                                                (/root/reference/src/big-num.cu:37-55, P = 469762049, g = 3)
 * This header is the bits = 64 instance of the drop-in boundary declared in ntt256.h (SURVEY.md par. 8f row
 * 4, second half), with the same conventions: `d_data` is a device pointer to n = 2^log_n elements,
 * transformed in place, natural order in and out, X[k] = sum_j x[j] * w_n^(jk) mod p with
 * w_n = g^((p-1)/n) for the generator g given at context creation (big-num.cu:43: gap = omega^((P-1)/(2 stride)));
 * calls are asynchronous on `stream`; status codes are those of ntt256.h.  No CPU fallback.
 *
 * Two field kinds behind the same entry points (csrc/field64.cuh):
 *   - p = 2^64 - 2^32 + 1 (Goldilocks; g = 7 is the usual generator): special reduction, no second multiply;
 *   - any other odd p < 2^63 with 2^max_log_n | p - 1 (e.g. the reference's 469762049, which makes this
 *     library's output directly comparable with the reference's own CPU NTT): Montgomery products.
 */
#ifndef NTT64_H_
#define NTT64_H_

#include "ntt256.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ntt64_ctx ntt64_ctx;

#define NTT64_GOLDILOCKS_P 0xFFFFFFFF00000001ull

/* p as above, g a generator (checked: w_N^(N/2) == -1 for N = 2^max_log_n); max_log_n <= 32 and <= the
 * two-adicity of p - 1 (NTT256_ERR_FIELD otherwise). */
int ntt64_ctx_create(ntt64_ctx** out, int device, uint64_t p, uint64_t g, int max_log_n);
void ntt64_ctx_destroy(ntt64_ctx* ctx);
/* Plan tuning (optional): tile_bits in [3,14] (shared-memory tile = 2^tile_bits elements of 8 bytes; default
 * 12), coalesce_bits in [0,6] (global access runs of >= 2^coalesce_bits elements; default 4 = 128 bytes). */
int ntt64_ctx_set_tuning(ntt64_ctx* ctx, int tile_bits, int coalesce_bits);

/* Goldilocks only: every 64th root of unity of that field is a power of two, so with the usual generator the
 * constant twiddles inside a register round (powers of w_16) are applied by shifts instead of products.  On by
 * default whenever w_16 allows it; enable = 0 keeps them on the multiplier (testing: same bits either way). */
int ntt64_ctx_set_shift_twiddles(ntt64_ctx* ctx, int enable);

int ntt64_forward(ntt64_ctx* ctx, void* d_data, int log_n, ntt256_stream stream);
int ntt64_inverse(ntt64_ctx* ctx, void* d_data, int log_n, ntt256_stream stream); /* includes 1/n */
/* `batch` contiguous transforms of 2^log_n points each (independent polynomials; shard across GPUs by
 * giving each rank its slice of the batch -- no collective). */
int ntt64_forward_batch(ntt64_ctx* ctx, void* d_data, int log_n, int batch, ntt256_stream stream);
int ntt64_inverse_batch(ntt64_ctx* ctx, void* d_data, int log_n, int batch, ntt256_stream stream);
/* Host buffers in and out (h_in may equal h_out): copies + transform, synchronous.  Batches are cut into
 * chunks that flow through a three-slot ring so the two PCIe directions and the kernels overlap (pinned
 * buffers make the copies asynchronous; pageable ones still give the right answer). */
int ntt64_forward_host(ntt64_ctx* ctx, const void* h_in, void* h_out, int log_n, int batch);
int ntt64_inverse_host(ntt64_ctx* ctx, const void* h_in, void* h_out, int log_n, int batch);
/* Coset transforms: forward evaluates on shift * <w_n> (x[j] *= shift^j fused into the first pass's loads),
 * inverse undoes it (shift^-k fused into the last pass's stores together with 1/n).  0 < shift < p. */
int ntt64_coset_forward(ntt64_ctx* ctx, void* d_data, int log_n, int batch, uint64_t shift, ntt256_stream stream);
int ntt64_coset_inverse(ntt64_ctx* ctx, void* d_data, int log_n, int batch, uint64_t shift, ntt256_stream stream);

/* The multiplied-out shift^j table of a coset transform takes 8 * 2^log_n bytes (cached, four tables at most); above
 * max_bytes (default 256 MiB) two factored tables of 2^(log_n/2) entries and a separate element-wise pass are used
 * instead.  Drops the cached tables. */
int ntt64_ctx_set_coset_table_limit(ntt64_ctx* ctx, size_t max_bytes);

/* Multi-GPU four-step pieces, exactly as ntt256_fourstep_* (ntt256.h): 2^log_p ranks, n = 2^log_r rows x
 * 2^(log_n - log_r) columns, rank = column block; step 1, one exchange of contiguous row blocks (an all-to-all, or
 * none at all with *_peer, whose last pass stores every element straight into the destination rank's receive buffer
 * through peer-mapped pointers; ntt256_ipc_* and ntt256_peer_barrier are width-independent), step 2.  The inverse pair
 * consumes the forward's output shards and returns its input shards.  Goldilocks transforms of 2^28 .. 2^32 points
 * are the use: a shard of 2^32 / 8 elements is 4 GiB. */
int ntt64_fourstep_step1(ntt64_ctx* ctx, void* d_shard, int log_n, int log_r, int log_p, int rank, ntt256_stream stream);
int ntt64_fourstep_step1_peer(ntt64_ctx* ctx, void* d_shard, const void* const* peer_recv, int log_n, int log_r, int log_p,
                              int rank, ntt256_stream stream);
int ntt64_fourstep_step2(ntt64_ctx* ctx, void* d_shard, int log_n, int log_r, int log_p, int rank, ntt256_stream stream);
int ntt64_fourstep_inverse_step1(ntt64_ctx* ctx, void* d_shard, int log_n, int log_r, int log_p, int rank, ntt256_stream stream);
int ntt64_fourstep_inverse_step1_peer(ntt64_ctx* ctx, void* d_shard, const void* const* peer_recv, int log_n, int log_r,
                                      int log_p, int rank, ntt256_stream stream);
int ntt64_fourstep_inverse_step2(ntt64_ctx* ctx, void* d_shard, int log_n, int log_r, int log_p, int rank, ntt256_stream stream);

/* Element-wise product of n canonical elements, out = a*b mod p; out may alias a or b. */
int ntt64_vec_mul(ntt64_ctx* ctx, void* d_out, const void* d_a, const void* d_b, unsigned long long n,
                  ntt256_stream stream);

unsigned long long ntt64_launch_count(const ntt64_ctx* ctx);
const char* ntt64_last_cuda_error(const ntt64_ctx* ctx);
/* Host-only: JSON description of the pass plan for 2^log_n points with the default tuning and the shape
 * chosen for a call of `batch` such transforms (size returned, like ntt256_plan_describe). */
size_t ntt64_plan_describe(int log_n, int batch, int max_log_n, char* buf, size_t buf_len);
/* The same for one local step of the four-step decomposition (step 1 / 2 forward, 3 / 4 inverse) on `rank`. */
size_t ntt64_plan_describe_fourstep(int step, int log_n, int log_r, int log_p, int rank, int max_log_n, char* buf,
                                    size_t buf_len);

#ifdef __cplusplus
}
#endif
#endif
