// Host-only check of the index tables the pass kernels look up (ntt_b200/csrc/pass_tables.h): for plans of every
// container width's shape (2 / 4 / 8 / 16 elements per thread), whole transforms and four-step pieces, the tables
// must describe a partition of the data -- every element of the array is loaded exactly once and stored exactly
// once per pass, in place (a tile stores to the addresses it loaded) -- and carry the swizzle of the index they name.
// Test infrastructure; built and run by tests/test_capi.py without a GPU.
#include <cstdio>
#include <cstdlib>
#include <set>
#include <vector>

#include "../../ntt_b200/csrc/pass_tables.h"
#include "../../ntt_b200/csrc/plan.h"

using namespace ntt256;

static int failures = 0;
#define CHECK(c)                                                          \
  do {                                                                    \
    if (!(c)) {                                                           \
      std::printf("FAILED %s (line %d)\n", #c, __LINE__);                 \
      ++failures;                                                         \
    }                                                                     \
  } while (0)

static void check_plan(const Plan& plan, int nbits) {
  const int GB = plan.group_bits;
  for (const PassDesc& pd : plan.passes) {
    std::vector<uint32_t> thr;
    std::vector<uint2> blk;
    compute_pass_tables(pd, GB, &thr, &blk);
    const int T = pd.T;
    const uint32_t t = 1u << (T - GB);
    CHECK(pd.T + pd.nblk_bits == nbits);
    CHECK(thr.size() == (size_t)(3 + pd.nrounds) * t);
    CHECK(blk.size() == ((size_t)1 << pd.nblk_bits));
    std::vector<char> loaded((size_t)1 << nbits, 0), stored((size_t)1 << nbits, 0);
    for (size_t b = 0; b < blk.size(); ++b) {
      std::set<uint32_t> in_addr, out_addr;
      for (uint32_t tid = 0; tid < t; ++tid) {
        for (int r = 0; r < pd.nrounds; ++r) {
          const uint32_t lw = thr[(size_t)(3 + r) * t + tid];
          CHECK((lw >> 16) == pass_tables_swz(lw & 0xffffu));
          CHECK((lw & 0xffffu) < (1u << T));
        }
        const int sb_first = pd.round_sb[0], sb_last = pd.round_sb[pd.nrounds - 1];
        for (int u = 0; u < (1 << GB); ++u) {
          uint32_t oi = thr[tid], oo = thr[t + tid];
          for (int i = 0; i < GB; ++i)
            if ((u >> i) & 1) {
              oi += 1u << pd.in_pos[sb_first + i];
              oo += 1u << pd.out_pos[sb_last + i];
            }
          const uint32_t ai = blk[b].x + oi, ao = blk[b].x + oo;
          CHECK(ai < (1u << nbits) && ao < (1u << nbits));
          if (ai < (1u << nbits)) ++loaded[ai];
          if (ao < (1u << nbits)) ++stored[ao];
          in_addr.insert(ai);
          out_addr.insert(ao);
        }
      }
      CHECK(in_addr == out_addr);  // in place: a tile stores where it loaded
      CHECK(in_addr.size() == ((size_t)1 << T));
    }
    for (size_t a = 0; a < loaded.size(); ++a) {
      if (loaded[a] != 1 || stored[a] != 1) {
        CHECK(loaded[a] == 1 && stored[a] == 1);
        break;
      }
    }
  }
}

int main() {
  int plans = 0;
  for (int gb : {1, 2, 3, 4}) {
    for (int log_n : {3, 5, 8, 10, 12, 13, 16}) {
      for (int T : {4, 8, 11, 12, 13}) {
        if (T < gb || T > log_n + 2) continue;
        PlanRequest req;
        req.log_n = log_n;
        req.tile_bits = T < log_n ? T : log_n;
        req.coalesce_bits = req.tile_bits >= 6 ? 2 : 0;
        req.max_log_n = 20;
        req.group_bits = gb;
        req.max_tile_bits = kMaxTileBitsAny;
        Plan plan;
        if (req.tile_bits < gb || !make_plan(req, &plan, 12)) continue;
        check_plan(plan, log_n);
        ++plans;
      }
    }
  }
  // four-step pieces (forward and inverse, both steps) of a 2^16-point transform over 4 ranks
  for (int kind = 1; kind <= 4; ++kind)
    for (int gb : {2, 3, 4}) {
      const bool first = (kind == 1 || kind == 3);
      PlanRequest req = fourstep_request(FourStepKey{kind, 16, 8, 2, first ? 3 : 0}, 10, 2, 20);
      req.group_bits = gb;
      req.max_tile_bits = kMaxTileBitsAny;
      Plan plan;
      if (!make_plan(req, &plan, 12)) {
        CHECK(!"four-step plan");
        continue;
      }
      check_plan(plan, 14);
      ++plans;
    }
  std::printf("%d plans checked, %d failures\n", plans, failures);
  return failures ? 1 : (plans >= 40 ? 0 : 2);
}
