/* libntt256 -- 384-bit companion API: forward / inverse NTT over prime fields held in 384-bit
 * containers (12 little-endian u32 limbs, 48 bytes, canonical residues in [0, p), p < 2^383), e.g.
 * BLS12-377 Fq = BW6-761 Fr (377 bits, 2^46 | p - 1).
 *
 * The reference's drivers and kernels are templates over the container width,
 *   template<uint tpi, uint bits> void NTT_GZKP(cgbn_mem_t<bits> data[], uint len, ...)
 *                                                (/root/reference/src/big-num.cu:57, 67, 172, 260)
 * instantiated there only at bits = 256 (big-num.cu:12); this header is the bits = 384 instance of
 * the drop-in boundary declared in ntt256.h, with the same conventions: `d_data` is a device
 * pointer, n = 2^log_n elements (cgbn_mem_t<384> layout, include/cgbn/cgbn_cuda.h:51-55),
 * transformed in place, natural order in and out, X[k] = sum_j x[j] * w_n^(jk) mod p with
 * w_n = g^((p-1)/n) for the generator g given at context creation; calls are asynchronous on
 * `stream`; status codes are those of ntt256.h.  SURVEY.md par. 8f row 4: batches, inverse with 1/n, the
 * element-wise product, host buffers, coset transforms / pointwise factors fused into the end passes, and the
 * local steps of the multi-GPU four-step transform (the kernels are the 12-limb instances of the same pass
 * body as the 256-bit ones, csrc/ntt_pass_generic.cuh).  No CPU fallback.
 */
#ifndef NTT384_H_
#define NTT384_H_

#include "ntt256.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ntt384_ctx ntt384_ctx;

/* p odd, p < 2^383, 2^max_log_n | p - 1, g a generator (checked: w_N^(N/2) == -1); max_log_n <= 30. */
int ntt384_ctx_create(ntt384_ctx** out, int device, const uint32_t p[12], const uint32_t g[12], int max_log_n);
void ntt384_ctx_destroy(ntt384_ctx* ctx);
/* Plan tuning (optional): tile_bits in [3,11] (shared-memory tile = 2^tile_bits elements of 48 bytes;
 * default 9), coalesce_bits in [0,6] (global access runs of >= 2^coalesce_bits elements; default 1). */
int ntt384_ctx_set_tuning(ntt384_ctx* ctx, int tile_bits, int coalesce_bits);

int ntt384_forward(ntt384_ctx* ctx, void* d_data, int log_n, ntt256_stream stream);
int ntt384_inverse(ntt384_ctx* ctx, void* d_data, int log_n, ntt256_stream stream); /* includes 1/n */
/* `batch` contiguous transforms of 2^log_n points each. */
int ntt384_forward_batch(ntt384_ctx* ctx, void* d_data, int log_n, int batch, ntt256_stream stream);
int ntt384_inverse_batch(ntt384_ctx* ctx, void* d_data, int log_n, int batch, ntt256_stream stream);

/* Host buffers in and out (h_in may equal h_out), synchronous.  Pinned buffers are read by the first pass and
 * written by the last pass directly (zero-copy over PCIe); pageable ones go through staged copies. */
int ntt384_forward_host(ntt384_ctx* ctx, const void* h_in, void* h_out, int log_n, int batch);
int ntt384_inverse_host(ntt384_ctx* ctx, const void* h_in, void* h_out, int log_n, int batch);

/* Coset transforms (log_n >= 3): forward = NTT of x[j] * shift^j, the factors applied as the first pass loads;
 * inverse undoes it (shift^-k applied before the last pass stores).  shift: 12 LE limbs, 0 < shift < p.  The
 * multiplied-out table (48 * 2^log_n bytes) is built on first use and cached (two tables). */
int ntt384_coset_forward(ntt384_ctx* ctx, void* d_data, int log_n, int batch, const uint32_t shift[12], ntt256_stream stream);
int ntt384_coset_inverse(ntt384_ctx* ctx, void* d_data, int log_n, int batch, const uint32_t shift[12], ntt256_stream stream);
/* Transform followed by the element-wise product with d_mul_mont (batch * 2^log_n Montgomery-form elements,
 * i.e. the plain product with the values they represent), fused into the last pass's stores. */
int ntt384_forward_mul(ntt384_ctx* ctx, void* d_data, const void* d_mul_mont, int log_n, int batch, ntt256_stream stream);
int ntt384_inverse_mul(ntt384_ctx* ctx, void* d_data, const void* d_mul_mont, int log_n, int batch, ntt256_stream stream);

/* Multi-GPU four-step pieces, exactly as ntt256_fourstep_* (ntt256.h): 2^log_p ranks, n = 2^log_r rows x
 * 2^(log_n - log_r) columns, rank = column block; step 1 (column transforms + fused twiddle), one exchange of
 * contiguous row blocks -- NCCL all-to-all, or none at all with *_peer, whose last pass stores every element
 * straight into the destination rank's receive buffer (peer-mapped pointers, one per rank; ntt256_ipc_*) --
 * then step 2.  The inverse pair consumes the forward's output shards and returns its input shards.
 * ntt384_ctx_set_peer_sync arms the rank synchronisation folded into those kernels (flag buffers as for
 * ntt256_ctx_set_peer_sync; NULL disarms). */
int ntt384_fourstep_step1(ntt384_ctx* ctx, void* d_shard, int log_n, int log_r, int log_p, int rank, ntt256_stream stream);
int ntt384_fourstep_step1_peer(ntt384_ctx* ctx, void* d_shard, const void* const* peer_recv, int log_n, int log_r,
                               int log_p, int rank, ntt256_stream stream);
int ntt384_fourstep_step2(ntt384_ctx* ctx, void* d_shard, int log_n, int log_r, int log_p, int rank, ntt256_stream stream);
int ntt384_fourstep_inverse_step1(ntt384_ctx* ctx, void* d_shard, int log_n, int log_r, int log_p, int rank,
                                  ntt256_stream stream);
int ntt384_fourstep_inverse_step1_peer(ntt384_ctx* ctx, void* d_shard, const void* const* peer_recv, int log_n,
                                       int log_r, int log_p, int rank, ntt256_stream stream);
int ntt384_fourstep_inverse_step2(ntt384_ctx* ctx, void* d_shard, int log_n, int log_r, int log_p, int rank,
                                  ntt256_stream stream);
int ntt384_ctx_set_peer_sync(ntt384_ctx* ctx, void* const* flag_bufs, int nranks, int rank, double timeout_s);

/* Element-wise product of n elements; mont_form = 0: canonical data (out = a*b mod p),
 * mont_form = 1: Montgomery-form data in and out (out = a*b*2^-384 mod p).  out may alias a or b. */
int ntt384_vec_mul(ntt384_ctx* ctx, void* d_out, const void* d_a, const void* d_b, unsigned long long n, int mont_form,
                   ntt256_stream stream);

unsigned long long ntt384_launch_count(const ntt384_ctx* ctx);
const char* ntt384_last_cuda_error(const ntt384_ctx* ctx);
/* Host-only: JSON description of the pass plan for 2^log_n points (size returned, like ntt256_plan_describe). */
size_t ntt384_plan_describe(int log_n, int max_log_n, char* buf, size_t buf_len);
/* The same for one local step of the four-step decomposition (step 1 / 2 forward, 3 / 4 inverse) on `rank`. */
size_t ntt384_plan_describe_fourstep(int step, int log_n, int log_r, int log_p, int rank, int max_log_n, char* buf,
                                     size_t buf_len);

#ifdef __cplusplus
}
#endif
#endif
