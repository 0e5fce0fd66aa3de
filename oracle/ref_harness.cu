// TEST INFRASTRUCTURE ONLY.  Wraps the reference's own CPU functions -- compiled from the
// sources where they lie under /root/reference, nothing is copied into this repo -- behind a
// C ABI so tests can pin the oracle against the real thing (p = 469762049, g = 3, int64).
//   NTT()                 /root/reference/src/self-sort-in-place.cu:32-51  (== big-num.cu:37-55)
//   NTT_dif()             :53-77
//   NTT_pro1()/NTT_pro2() :79-128  (self-sort-in-place ordering)
// Built only when /root/reference exists (this container); the GPU box uses the prebuilt .so.
// Output: oracle/_ref/libref_i64.so (git-ignored, travels with gpurun).
#include <vector>
#define main ref_main_unused
#include "/root/reference/src/self-sort-in-place.cu"
#undef main

static const long long kRefP = P;
static const long long kRefRoot = root;
#undef root
#undef P

static std::vector<long long> make_reverse(unsigned log_n) {
  std::vector<long long> rev((size_t)1 << log_n, 0);
  for (size_t i = 0; i < rev.size(); ++i)
    rev[i] = log_n ? ((rev[i >> 1] >> 1) | ((long long)(i & 1) << (log_n - 1))) : 0;
  return rev;
}

extern "C" {
long long ref_modulus(void) { return kRefP; }
long long ref_generator(void) { return kRefRoot; }
void ref_ntt_i64(long long* data, unsigned log_n) {
  std::vector<long long> rev = make_reverse(log_n);
  NTT(data, rev.data(), 1ll << log_n, kRefRoot);
}
void ref_ntt_dif_i64(long long* data, unsigned log_n) {
  std::vector<long long> rev = make_reverse(log_n);
  NTT_dif(data, rev.data(), log_n, kRefRoot);
}
void ref_ssip_i64(long long* data, unsigned log_n) {
  NTT_pro1(data, log_n, kRefRoot);
  NTT_pro2(data, log_n, kRefRoot);
}
}
