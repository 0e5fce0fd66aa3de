// Pass planner for the self-sorting in-place NTT (host side, no CUDA types).
//
// The transform is a radix-2 DIF factorisation executed as a few "passes".  A pass loads a
// tile of 2^T elements chosen by T index-bit positions, runs up to T butterfly stages on it
// in registers/shared memory, and stores the tile back to the SAME set of addresses with the
// index bits permuted.  Fusing the bit permutation into the store is what removes the
// reference's bit-reversal kernel (`rearrange`, /root/reference/src/big-num.cu:57-65): it is
// the generalisation of NTT_pro2's "butterfly fused with the swap of bits i-1 <-> L-i"
// (/root/reference/src/self-sort-in-place.cu:96-128) from one bit pair per stage to a whole
// tile of bits per pass.
//
// Bookkeeping: every physical index-bit position carries a label -- an unprocessed time bit
// j_b or a frequency bit k_c (processing j_b yields k_{L-1-b}).  The planner searches for the
// shortest pass sequence that processes the time bits in DIF order (L-1 .. 0), keeps `g` low
// address bits inside every tile (so each row of a tile is a run of >= 2^g contiguous
// elements = 32*2^g bytes), and ends with k_c at position c (natural order).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace ntt256 {

constexpr int kMaxTileBits = 12;   // 4096 elements = 128 KiB of shared memory (256-bit elements)
constexpr int kMaxTileBitsAny = 14;  // narrower elements may ask for more (PlanRequest::max_tile_bits; arrays hold 16)
constexpr int kMaxRounds = 12;  // T = 12 in radix-8 rounds needs 4; the two-element kernel one per stage
constexpr int kMaxIndexBits = 40;  // log2(n * batch) the descriptors can address

// One pass, as consumed by the device kernel (trivially copyable; passed by value as a
// __grid_constant__ kernel parameter).
struct PassDesc {
  int32_t T;           // log2(tile elements); the CTA has max(2^(T-3), 1) threads... see kernels
  int32_t d;           // butterfly stages in this pass (local bits [T-d, T))
  int32_t nblk_bits;   // index bits enumerated by the block index (total_bits - T)
  int32_t nrounds;     // register rounds
  int32_t is_final;    // 1: outputs are made canonical [0,p) on the way out
  int32_t has_tw;      // 1: outputs are multiplied by the pass twiddle w_{N'}^(j_low * k_hi)
  int32_t tw_lshift;   // exponent scale: E = (j_low*k_hi) << tw_lshift  (mod 2^max_log_n)
  int32_t scale_out;   // 1: multiply outputs by the plan's scale constant (inverse NTT, 1/n)
  uint32_t jlow_const; // contribution of index bits outside this device's memory (multi-GPU)
  int8_t in_pos[16];   // physical bit position of local bit i when loading
  int8_t out_pos[16];  // physical bit position of local bit i when storing
  int8_t blk_pos[kMaxIndexBits];  // physical bit position of block-index bit i
  uint32_t wt_loc[16];            // time weight (2^b or 0) of passive local bit i
  uint32_t wt_blk[kMaxIndexBits]; // time weight of block-index bit i
  int8_t round_sb[kMaxRounds];    // lowest local bit of the round's group (group_bits wide)
  int8_t round_nb[kMaxRounds];    // butterfly stages in the round (top nb bits of the group)
  int8_t tmap[kMaxRounds][16];    // local bit fed by thread-index bit i in round r
};

struct Plan {
  int log_n = 0;
  int tile_bits = 0;   // requested T
  int coalesce_bits = 0;
  int group_bits = 3;  // log2(elements per thread) the round schedule was made for (3, 2 or 1)
  std::vector<PassDesc> passes;
  unsigned long long id = 0;     // unique per created plan (keys per-plan device tables)
  std::string describe() const;  // JSON, for tests and DESIGN.md
};

// Label bookkeeping exposed for the multi-GPU driver and the tests.
struct LabelState {
  // label[p]: value b in [0,64) = unprocessed time bit j_b ; 64 + c = frequency bit k_c ;
  //           128 + x = opaque passenger bit (batch / foreign), never processed.
  std::vector<int> label;
};

struct PlanRequest {
  int log_n;          // transform length 2^log_n, time bits log_n-1 .. 0
  int tile_bits;      // max T
  int coalesce_bits;  // g: low physical bits kept in every tile
  int max_log_n;      // twiddle tables are for w_{2^max_log_n}
  // Optional generalisation used by the four-step multi-GPU driver: explicit start labels,
  // explicit targets and the range of time bits to process now.
  std::vector<int> start;   // empty: label[p] = j_p
  std::vector<int> target;  // empty: position c must end with k_c.  -1 entries = don't care
  int time_hi = -1, time_lo = 0;  // process j_time_hi .. j_time_lo (default log_n-1 .. 0)
  uint32_t jlow_const = 0;
  int total_log_n = -1;     // length of the whole (possibly distributed) transform; default log_n
  int group_bits = 3;       // 3: eight elements per thread (radix-8 rounds); 2: four (radix-4); 1: two (radix-2);
                            // 4: sixteen (radix-16 rounds; the 64-bit kernels, whose elements are two registers)
  int max_tile_bits = kMaxTileBits;  // upper limit for tile_bits (<= kMaxTileBitsAny)
};

// Which piece of work a plan is for: a whole transform (kind 0) or one local step of the four-step
// decomposition over 2^log_p ranks (SURVEY.md par. 8e): 1 / 2 = forward step 1 / 2, 3 / 4 = inverse step 1 / 2.
struct FourStepKey {
  int kind, log_n, log_r, log_p, rank;
};
// The request (start labels, targets, time range, foreign index bits) of that piece in the planner's label model;
// shared by every container width.
PlanRequest fourstep_request(const FourStepKey& key, int tile_bits, int coalesce_bits, int max_log_n);
inline bool fourstep_args_ok(int log_n, int log_r, int log_p, int rank) {
  return log_p >= 0 && log_r >= log_p && log_n - log_r >= log_p && rank >= 0 && rank < (1 << log_p) &&
         log_n - log_p >= 3 && log_r >= 1 && log_n - log_r >= 1;
}

// Returns false when no plan with <= max_passes passes exists under the constraints.
bool make_plan(const PlanRequest& req, Plan* out, int max_passes = 8);

}  // namespace ntt256
