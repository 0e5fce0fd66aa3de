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
 * `stream`; status codes are those of ntt256.h.  SURVEY.md par. 8f row 4, first cut: one GPU,
 * batches, inverse with 1/n, element-wise product.  No CPU fallback.
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

/* Element-wise product of n elements; mont_form = 0: canonical data (out = a*b mod p),
 * mont_form = 1: Montgomery-form data in and out (out = a*b*2^-384 mod p).  out may alias a or b. */
int ntt384_vec_mul(ntt384_ctx* ctx, void* d_out, const void* d_a, const void* d_b, unsigned long long n, int mont_form,
                   ntt256_stream stream);

unsigned long long ntt384_launch_count(const ntt384_ctx* ctx);
const char* ntt384_last_cuda_error(const ntt384_ctx* ctx);
/* Host-only: JSON description of the pass plan for 2^log_n points (size returned, like ntt256_plan_describe). */
size_t ntt384_plan_describe(int log_n, int max_log_n, char* buf, size_t buf_len);

#ifdef __cplusplus
}
#endif
#endif
