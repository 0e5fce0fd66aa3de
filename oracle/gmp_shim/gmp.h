/* Minimal declarations of the GNU MP entry points that the reference's 256-bit program
 * (/root/reference/src/big-num.cu and the CGBN host path it includes) calls.  Test infrastructure:
 * this image ships the GMP runtime (libgmp.so.10) but not its header, so oracle/Makefile's
 * `ref256` target compiles the UNMODIFIED reference program against this file and links the
 * system library.  Only the ABI of libgmp 6.x is restated (struct layout and the __gmpz_* symbol
 * names from GMP's public documentation); nothing here is used by the product library. */
#ifndef NTT_ORACLE_GMP_SHIM_H_
#define NTT_ORACLE_GMP_SHIM_H_
#define __GMP_H__ /* the include guard of the real header: CGBN selects its GMP host path on it */
#include <stddef.h>
#include <stdio.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef unsigned long mp_limb_t;
typedef unsigned long mp_bitcnt_t;
typedef struct {
  int _mp_alloc;
  int _mp_size;
  mp_limb_t* _mp_d;
} __mpz_struct;
typedef __mpz_struct mpz_t[1];
typedef const __mpz_struct* mpz_srcptr;
typedef __mpz_struct* mpz_ptr;

#define NTT_GMP_FN(name) __gmpz_##name
#define mpz_init NTT_GMP_FN(init)
#define mpz_clear NTT_GMP_FN(clear)
#define mpz_set NTT_GMP_FN(set)
#define mpz_set_ui NTT_GMP_FN(set_ui)
#define mpz_set_si NTT_GMP_FN(set_si)
#define mpz_get_ui NTT_GMP_FN(get_ui)
#define mpz_get_si NTT_GMP_FN(get_si)
#define mpz_swap NTT_GMP_FN(swap)
#define mpz_add NTT_GMP_FN(add)
#define mpz_add_ui NTT_GMP_FN(add_ui)
#define mpz_sub NTT_GMP_FN(sub)
#define mpz_sub_ui NTT_GMP_FN(sub_ui)
#define mpz_ui_sub NTT_GMP_FN(ui_sub)
#define mpz_mul NTT_GMP_FN(mul)
#define mpz_mul_ui NTT_GMP_FN(mul_ui)
#define mpz_submul NTT_GMP_FN(submul)
#define mpz_mul_2exp NTT_GMP_FN(mul_2exp)
#define mpz_fdiv_q NTT_GMP_FN(fdiv_q)
#define mpz_fdiv_r NTT_GMP_FN(fdiv_r)
#define mpz_fdiv_qr NTT_GMP_FN(fdiv_qr)
#define mpz_fdiv_q_ui NTT_GMP_FN(fdiv_q_ui)
#define mpz_fdiv_r_ui NTT_GMP_FN(fdiv_r_ui)
#define mpz_fdiv_q_2exp NTT_GMP_FN(fdiv_q_2exp)
#define mpz_fdiv_r_2exp NTT_GMP_FN(fdiv_r_2exp)
#define mpz_powm NTT_GMP_FN(powm)
#define mpz_sqrt NTT_GMP_FN(sqrt)
#define mpz_gcd NTT_GMP_FN(gcd)
#define mpz_invert NTT_GMP_FN(invert)
#define mpz_cmp NTT_GMP_FN(cmp)
#define mpz_cmp_ui NTT_GMP_FN(cmp_ui)
#define mpz_and NTT_GMP_FN(and)
#define mpz_ior NTT_GMP_FN(ior)
#define mpz_xor NTT_GMP_FN(xor)
#define mpz_popcount NTT_GMP_FN(popcount)
#define mpz_scan1 NTT_GMP_FN(scan1)
#define mpz_setbit NTT_GMP_FN(setbit)
#define mpz_clrbit NTT_GMP_FN(clrbit)
#define mpz_tstbit NTT_GMP_FN(tstbit)
#define mpz_sizeinbase NTT_GMP_FN(sizeinbase)
#define mpz_import NTT_GMP_FN(import)
#define mpz_export NTT_GMP_FN(export)
#define mpz_out_str NTT_GMP_FN(out_str)
#define mpz_sgn(z) ((z)->_mp_size < 0 ? -1 : (z)->_mp_size > 0)

void mpz_init(mpz_ptr);
void mpz_clear(mpz_ptr);
void mpz_set(mpz_ptr, mpz_srcptr);
void mpz_set_ui(mpz_ptr, unsigned long);
void mpz_set_si(mpz_ptr, long);
unsigned long mpz_get_ui(mpz_srcptr);
long mpz_get_si(mpz_srcptr);
void mpz_swap(mpz_ptr, mpz_ptr);
void mpz_add(mpz_ptr, mpz_srcptr, mpz_srcptr);
void mpz_add_ui(mpz_ptr, mpz_srcptr, unsigned long);
void mpz_sub(mpz_ptr, mpz_srcptr, mpz_srcptr);
void mpz_sub_ui(mpz_ptr, mpz_srcptr, unsigned long);
void mpz_ui_sub(mpz_ptr, unsigned long, mpz_srcptr);
void mpz_mul(mpz_ptr, mpz_srcptr, mpz_srcptr);
void mpz_mul_ui(mpz_ptr, mpz_srcptr, unsigned long);
void mpz_submul(mpz_ptr, mpz_srcptr, mpz_srcptr);
void mpz_mul_2exp(mpz_ptr, mpz_srcptr, mp_bitcnt_t);
void mpz_fdiv_q(mpz_ptr, mpz_srcptr, mpz_srcptr);
void mpz_fdiv_r(mpz_ptr, mpz_srcptr, mpz_srcptr);
void mpz_fdiv_qr(mpz_ptr, mpz_ptr, mpz_srcptr, mpz_srcptr);
unsigned long mpz_fdiv_q_ui(mpz_ptr, mpz_srcptr, unsigned long);
unsigned long mpz_fdiv_r_ui(mpz_ptr, mpz_srcptr, unsigned long);
void mpz_fdiv_q_2exp(mpz_ptr, mpz_srcptr, mp_bitcnt_t);
void mpz_fdiv_r_2exp(mpz_ptr, mpz_srcptr, mp_bitcnt_t);
void mpz_powm(mpz_ptr, mpz_srcptr, mpz_srcptr, mpz_srcptr);
void mpz_sqrt(mpz_ptr, mpz_srcptr);
void mpz_gcd(mpz_ptr, mpz_srcptr, mpz_srcptr);
int mpz_invert(mpz_ptr, mpz_srcptr, mpz_srcptr);
int mpz_cmp(mpz_srcptr, mpz_srcptr);
int mpz_cmp_ui(mpz_srcptr, unsigned long);
void mpz_and(mpz_ptr, mpz_srcptr, mpz_srcptr);
void mpz_ior(mpz_ptr, mpz_srcptr, mpz_srcptr);
void mpz_xor(mpz_ptr, mpz_srcptr, mpz_srcptr);
mp_bitcnt_t mpz_popcount(mpz_srcptr);
mp_bitcnt_t mpz_scan1(mpz_srcptr, mp_bitcnt_t);
void mpz_setbit(mpz_ptr, mp_bitcnt_t);
void mpz_clrbit(mpz_ptr, mp_bitcnt_t);
int mpz_tstbit(mpz_srcptr, mp_bitcnt_t);
size_t mpz_sizeinbase(mpz_srcptr, int);
void mpz_import(mpz_ptr, size_t, int, size_t, int, size_t, const void*);
void* mpz_export(void*, size_t*, int, size_t, int, size_t, mpz_srcptr);
size_t mpz_out_str(FILE*, int, mpz_srcptr);

#ifdef __cplusplus
}
#endif
#endif
