// Pass planner: see plan.h for the model.  Pure host C++ (no CUDA), exercised on CPU by
// tests/test_plan.py through ntt256_plan_describe().
#include "plan.h"

#include <algorithm>
#include <cstring>
#include <sstream>

namespace ntt256 {
namespace {

constexpr int kFreq = 64;  // labels: [0,64) time bit j_b, [64,128) frequency bit k_c, >= 128 passenger

struct Ctx {
  int nbits;       // local physical index bits
  int Ltot;        // length (log2) of the whole transform
  int T, g;
  int time_lo;
  std::vector<int> target;  // final label per position or -1
  std::vector<int> dest_of_label;  // indexed by label value (0..255) -> position or -1
};

struct PassChoice {
  std::vector<int> S;          // tile positions
  std::vector<int> proc_pos;   // positions of processed time bits, DIF order
  std::vector<int> new_label;  // labels after the pass (full vector)
  int t_hi;                    // first processed time bit
};

inline int final_label(const Ctx& c, int label, int t_hi_all, int t_lo_all) {
  if (label < kFreq && label <= t_hi_all && label >= t_lo_all) return kFreq + (c.Ltot - 1 - label);
  return label;
}

struct Search {
  Ctx c;
  int time_hi_all;
  std::vector<std::vector<PassChoice>> solutions;
  size_t max_solutions = 4000;
  size_t max_nodes = 400000;
  size_t nodes = 0;

  bool done(const std::vector<int>& lab, int t) const {
    if (t >= c.time_lo) return false;
    for (int p = 0; p < c.nbits; ++p)
      if (c.target[p] >= 0 && lab[p] != c.target[p]) return false;
    return true;
  }

  int dest(int label) const {
    int fl = final_label(c, label, time_hi_all, c.time_lo);
    return c.dest_of_label[fl];
  }

  // Build a pass: process `dcount` bits starting at time bit t; `fill` selects the heuristic
  // used to spend the remaining tile capacity.
  bool build(const std::vector<int>& lab, int t, int dcount, int fill, PassChoice* pc) const {
    std::vector<int> pos_of(256, -1);
    for (int p = 0; p < c.nbits; ++p) pos_of[lab[p]] = p;
    std::vector<char> in(c.nbits, 0);
    std::vector<int> S;
    pc->proc_pos.clear();
    auto add = [&](int p) {
      if (!in[p]) {
        in[p] = 1;
        S.push_back(p);
      }
    };
    for (int i = 0; i < dcount; ++i) {
      int p = pos_of[t - i];
      if (p < 0) return false;  // that time bit is not in this device's memory
      pc->proc_pos.push_back(p);
      add(p);
    }
    if (c.nbits <= c.T) {
      for (int p = 0; p < c.nbits; ++p) add(p);
    } else {
      for (int p = 0; p < c.g && p < c.nbits; ++p) add(p);
    }
    if ((int)S.size() > c.T) return false;
    const int t_lo_pass = t - dcount + 1;
    auto lab_after = [&](int p) {
      int l = lab[p];
      if (l < kFreq && l <= t && l >= t_lo_pass) return kFreq + (c.Ltot - 1 - l);
      return l;
    };
    if (fill != 1 && c.nbits > c.T) {
      // spend the spare capacity on destinations so that labels can settle (cycle closure)
      bool grew = true;
      while (grew && (int)S.size() < c.T) {
        grew = false;
        std::vector<int> order;
        if (fill == 0) {
          order = pc->proc_pos;
        } else {
          order.assign(pc->proc_pos.rbegin(), pc->proc_pos.rend());
        }
        for (int p : S)
          if (std::find(order.begin(), order.end(), p) == order.end()) order.push_back(p);
        for (int p : order) {
          int dp = dest(lab[p]);
          // a label processed in a later pass cannot settle now, but its current position's
          // final owner can come in; follow the label that wants to end where we are instead
          if (dp >= 0 && !in[dp]) {
            add(dp);
            grew = true;
            break;
          }
        }
      }
    }
    // pad the tile with the lowest free positions (pure passengers): a tile below T bits means
    // a CTA with idle lanes and shorter contiguous runs, never less work
    for (int p = 0; p < c.nbits && (int)S.size() < c.T; ++p) add(p);
    std::sort(S.begin(), S.end());
    // placement
    std::vector<int> nl(lab);
    std::vector<int> slot_label(c.nbits, -1);
    std::vector<int> leftovers;
    for (int p : S) {
      int l = lab_after(p);
      int fl = final_label(c, l, time_hi_all, c.time_lo);
      int dp = c.dest_of_label[fl];
      if (dp >= 0 && in[dp]) {
        // (labels still to be processed may also wait at their destination)
        slot_label[dp] = l;
      } else {
        leftovers.push_back(p);
      }
    }
    std::vector<int> free_slots;
    for (int p : S)
      if (slot_label[p] < 0) free_slots.push_back(p);
    // stay put when possible, then ascending match
    std::vector<int> rest;
    for (int p : leftovers) {
      auto it = std::find(free_slots.begin(), free_slots.end(), p);
      if (it != free_slots.end()) {
        slot_label[p] = lab_after(p);
        free_slots.erase(it);
      } else {
        rest.push_back(p);
      }
    }
    if (rest.size() != free_slots.size()) return false;
    for (size_t i = 0; i < rest.size(); ++i) slot_label[free_slots[i]] = lab_after(rest[i]);
    for (int p : S) nl[p] = slot_label[p];
    pc->S = S;
    pc->new_label = nl;
    pc->t_hi = t;
    return true;
  }

  void dfs(const std::vector<int>& lab, int t, int depth, int limit, std::vector<PassChoice>& stack) {
    if (solutions.size() >= max_solutions || nodes >= max_nodes) return;
    ++nodes;
    if (done(lab, t)) {
      solutions.push_back(stack);
      return;
    }
    if (depth == limit) return;
    int remaining = t - c.time_lo + 1;
    if (remaining > (limit - depth) * c.T) return;
    int dmax = std::min(c.T, remaining);
    int dmin = remaining > 0 ? 1 : 0;
    for (int dcount = dmax; dcount >= dmin; --dcount) {
      // the passes still to come must be able to take the rest
      if (remaining - dcount > (limit - depth - 1) * c.T) break;
      for (int fill = 0; fill < 3; ++fill) {
        PassChoice pc;
        if (!build(lab, t, dcount, fill, &pc)) continue;
        if (dcount == 0 && pc.new_label == lab) continue;
        stack.push_back(pc);
        dfs(pc.new_label, t - dcount, depth + 1, limit, stack);
        stack.pop_back();
        if (c.nbits <= c.T) break;  // fills are identical when the tile is the whole array
      }
    }
  }
};

int rounds_of(int d, int gb) { return d == 0 ? 1 : (d + gb - 1) / gb; }

// Fill the round schedule of a pass: a short first round (d mod GB stages) followed by full
// rounds; each round owns a GB-bit group of local bits (GB = log2 of the elements a thread
// holds) and butterflies its top nb.
void schedule_rounds(PassDesc* pd, int GB) {
  const int T = pd->T, d = pd->d;
  int r = 0;
  int top = T;  // next unprocessed local bit + 1
  int first = d % GB;
  if (d == 0) {
    pd->round_sb[0] = (int8_t)std::max(T - GB, 0);
    pd->round_nb[0] = 0;
    r = 1;
  } else {
    if (first) {
      pd->round_sb[r] = (int8_t)std::max(top - GB, 0);
      pd->round_nb[r] = (int8_t)first;
      top -= first;
      ++r;
    }
    while (top > T - d) {
      pd->round_sb[r] = (int8_t)(top - GB);
      pd->round_nb[r] = (int8_t)GB;
      top -= GB;
      ++r;
    }
  }
  pd->nrounds = r;
  // thread-bit -> local-bit maps.  Lanes (thread bits 0..4) must (a) hit distinct shared
  // memory banks under the kernel's swizzle bank = (l ^ l>>5 ^ l>>10) & 31, i.e. use local
  // bits that are distinct mod 5, and (b) in the first/last round follow the lowest physical
  // address bits so global accesses coalesce.
  for (int rr = 0; rr < r; ++rr) {
    const int sb = pd->round_sb[rr];
    std::vector<int> avail;
    for (int b = 0; b < T; ++b)
      if (b < sb || b >= sb + GB) avail.push_back(b);
    std::vector<int> order;
    std::vector<char> used(T, 0);
    if (rr == 0 || rr == r - 1) {
      // coalescing first: ascending physical position of the bit (load side for round 0,
      // store side for the last round; a single-round pass prefers the load side)
      const int8_t* pos = (rr == 0) ? pd->in_pos : pd->out_pos;
      std::vector<int> byphys(avail);
      std::sort(byphys.begin(), byphys.end(), [&](int a, int b) { return pos[a] < pos[b]; });
      order = byphys;
    } else {
      // bank-conflict-free lanes: one bit per residue class mod 5, lowest first
      for (int cls = 0; cls < 5; ++cls)
        for (int b : avail)
          if (b % 5 == cls && !used[b]) {
            order.push_back(b);
            used[b] = 1;
            break;
          }
      for (int b : avail)
        if (!used[b]) {
          bool present = std::find(order.begin(), order.end(), b) != order.end();
          if (!present) order.push_back(b);
        }
    }
    for (int i = 0; i < 16; ++i) pd->tmap[rr][i] = 0;
    for (size_t i = 0; i < order.size(); ++i) pd->tmap[rr][i] = (int8_t)order[i];
  }
}

}  // namespace

bool make_plan(const PlanRequest& req, Plan* out, int max_passes) {
  Search s;
  Ctx& c = s.c;
  c.Ltot = req.total_log_n >= 0 ? req.total_log_n : req.log_n;
  c.T = std::min(req.tile_bits, std::min(req.max_tile_bits, kMaxTileBitsAny));
  std::vector<int> start = req.start;
  if (start.empty()) {
    start.resize(req.log_n);
    for (int p = 0; p < req.log_n; ++p) start[p] = p;
  }
  c.nbits = (int)start.size();
  if (c.nbits > kMaxIndexBits) return false;
  c.T = std::min(c.T, std::max(c.nbits, 3));
  c.g = std::min(req.coalesce_bits, c.T);
  s.time_hi_all = req.time_hi >= 0 ? req.time_hi : c.Ltot - 1;
  c.time_lo = req.time_lo;
  c.target = req.target;
  if (c.target.empty()) {
    c.target.resize(c.nbits);
    for (int p = 0; p < c.nbits; ++p) c.target[p] = kFreq + p;
  }
  c.dest_of_label.assign(256, -1);
  for (int p = 0; p < c.nbits; ++p)
    if (c.target[p] >= 0) c.dest_of_label[c.target[p]] = p;

  std::vector<PassChoice> best;
  bool found = false;
  for (int limit = 1; limit <= max_passes && !found; ++limit) {
    s.solutions.clear();
    s.nodes = 0;
    std::vector<PassChoice> stack;
    s.dfs(start, s.time_hi_all, 0, limit, stack);
    if (s.solutions.empty()) continue;
    found = true;
    // score: fewest register rounds, then longest contiguous runs, then balanced passes
    long best_score = -1;
    for (auto& sol : s.solutions) {
      long rounds = 0, minrun = 99, maxd = 0;
      for (auto& pc : sol) {
        int d = (int)pc.proc_pos.size();
        rounds += rounds_of(d, req.group_bits == 4 ? 4 : 3);
        int run = 0;
        while (run < (int)pc.S.size() && pc.S[run] == run) ++run;
        minrun = std::min<long>(minrun, run);
        maxd = std::max<long>(maxd, d);
      }
      long score = rounds * 10000 - minrun * 100 + maxd;
      if (best_score < 0 || score < best_score) {
        best_score = score;
        best = sol;
      }
    }
  }
  if (!found) return false;

  const int GB = (req.group_bits == 2 || req.group_bits == 1 || req.group_bits == 4) ? req.group_bits : 3;
  out->group_bits = GB;
  out->log_n = req.log_n;
  out->tile_bits = c.T;
  out->coalesce_bits = c.g;
  out->passes.clear();
  std::vector<int> lab = start;
  for (size_t pi = 0; pi < best.size(); ++pi) {
    const PassChoice& pc = best[pi];
    PassDesc pd;
    std::memset(&pd, 0, sizeof(pd));
    const int d = (int)pc.proc_pos.size();
    const int T = std::max((int)pc.S.size(), 3);
    if ((int)pc.S.size() < 3) return false;  // arrays below 8 elements use the tiny kernel
    if (d > 0 && (d + GB - 1) / GB > kMaxRounds) return false;
    pd.T = T;
    pd.d = d;
    pd.nblk_bits = c.nbits - T;
    // canonical local order: passive bits ascending by position, then processed bits with the
    // first processed (highest time bit) on top
    std::vector<int> local_pos(T, -1);
    std::vector<char> is_proc(c.nbits, 0);
    for (int p : pc.proc_pos) is_proc[p] = 1;
    int li = 0;
    for (int p : pc.S)
      if (!is_proc[p]) local_pos[li++] = p;
    for (int i = d - 1; i >= 0; --i) local_pos[li++] = pc.proc_pos[i];
    // where does each label go?
    std::vector<int> newpos_of_label(256, -1);
    for (int p = 0; p < c.nbits; ++p) newpos_of_label[pc.new_label[p]] = p;
    const int t_lo_pass = pc.t_hi - d + 1;
    for (int i = 0; i < T; ++i) {
      int p = local_pos[i];
      pd.in_pos[i] = (int8_t)p;
      int l = lab[p];
      int l2 = (l < kFreq && l <= pc.t_hi && l >= t_lo_pass && d > 0) ? kFreq + (c.Ltot - 1 - l) : l;
      pd.out_pos[i] = (int8_t)newpos_of_label[l2];
      // time weight of passive bits: unprocessed time bits contribute 2^b to j_low
      pd.wt_loc[i] = (i < T - d && l < kFreq) ? (1u << l) : 0u;
    }
    int bi = 0;
    std::vector<char> inS(c.nbits, 0);
    for (int p : pc.S) inS[p] = 1;
    for (int p = 0; p < c.nbits; ++p)
      if (!inS[p]) {
        pd.blk_pos[bi] = (int8_t)p;
        pd.wt_blk[bi] = (lab[p] < kFreq) ? (1u << lab[p]) : 0u;
        ++bi;
      }
    // pass twiddle: needed when any unprocessed time bit remains below this pass
    bool lower_left = (t_lo_pass > 0) && d > 0;
    pd.has_tw = lower_left ? 1 : 0;
    pd.tw_lshift = req.max_log_n - (pc.t_hi + 1);
    pd.jlow_const = req.jlow_const;
    pd.is_final = 0;
    pd.scale_out = 0;
    schedule_rounds(&pd, GB);
    out->passes.push_back(pd);
    lab = pc.new_label;
  }
  if (!out->passes.empty() && req.time_lo == 0) out->passes.back().is_final = 1;
  return true;
}

// Plan request of a whole transform (kind 0) or of one local step of the four-step
// decomposition (kinds 1 and 2, SURVEY.md par. 8e), expressed in the planner's label model.
PlanRequest fourstep_request(const FourStepKey& key, int tile_bits, int coalesce_bits, int max_log_n) {
  PlanRequest req;
  req.tile_bits = tile_bits;
  req.coalesce_bits = coalesce_bits;
  req.max_log_n = max_log_n;
  req.log_n = key.log_n;
  if (key.kind == 1 || key.kind == 2) {
    const int L = key.log_n, lr = key.log_r, lp = key.log_p, lc = L - lr;
    const int nloc = L - lp;
    req.total_log_n = L;
    req.start.resize(nloc);
    req.target.resize(nloc);
    if (key.kind == 1) {
      // local index = j1 * (C/P) + j2_local ; rank = top lp bits of j2
      for (int p = 0; p < lc - lp; ++p) req.start[p] = p, req.target[p] = p;
      for (int i = 0; i < lr; ++i) {
        req.start[lc - lp + i] = lc + i;
        req.target[lc - lp + i] = 64 + i;  // k1 natural above j2_local
      }
      req.time_hi = L - 1;
      req.time_lo = lc;
      req.jlow_const = (uint32_t)key.rank << (lc - lp);
    } else {
      // received layout [src rank][k1 local][j2 local]
      for (int p = 0; p < lc - lp; ++p) req.start[p] = p;
      for (int c = 0; c < lr - lp; ++c) req.start[lc - lp + c] = 64 + c;
      for (int i = 0; i < lp; ++i) req.start[lc - lp + lr - lp + i] = lc - lp + i;
      // output [k2 mod P][k1 local][k2 div P]: every field is transformed where it lies (the two
      // j2 fields are bit-reversed in place), so a C-point row NTT with C <= 2^tile_bits is ONE pass.
      // k2 bit i = k_{lr+i}; bits 0..lp-1 come from the old rank bits on top.
      for (int i = 0; i < lc - lp; ++i) req.target[i] = 64 + lr + lp + i;
      for (int c = 0; c < lr - lp; ++c) req.target[lc - lp + c] = 64 + c;
      for (int i = 0; i < lp; ++i) req.target[lc - lp + lr - lp + i] = 64 + lr + i;
      req.time_hi = lc - 1;
      req.time_lo = 0;
      req.jlow_const = 0;
    }
  }
  if (key.kind == 3 || key.kind == 4) {
    // Distributed INVERSE that consumes the forward's output shards and returns the forward's
    // input shards.  Its time index is K = k1 + R*k2 (the forward's frequency index), viewed as
    // R' x C' with C' = 2^lr columns (K2 = k1) and R' = 2^lc rows (K1 = k2): the rank is again the
    // top lp bits of the column index.  Local input layout = [k2 mod P][k1 local][k2 div P].
    const int L = key.log_n, lr = key.log_r, lp = key.log_p, lc = L - lr;
    const int nloc = L - lp;
    req.total_log_n = L;
    req.start.assign(nloc, 0);
    req.target.assign(nloc, 0);
    if (key.kind == 3) {
      for (int i = 0; i < lc - lp; ++i) req.start[i] = lr + lp + i;              // k2 div P
      for (int i = 0; i < lr - lp; ++i) req.start[lc - lp + i] = i;              // k1 local
      for (int i = 0; i < lp; ++i) req.start[lc - lp + lr - lp + i] = lr + i;    // k2 mod P
      // step 1 transforms K bits L-1 .. lr -> J bits 0 .. lc-1
      for (int i = 0; i < lc - lp; ++i) req.target[i] = 64 + i;                  // j2 local, final
      for (int i = 0; i < lr - lp; ++i) req.target[lc - lp + i] = i;             // k1 local waits
      for (int i = 0; i < lp; ++i) req.target[lc - lp + lr - lp + i] = 64 + lc - lp + i;  // destination rank
      req.time_hi = L - 1;
      req.time_lo = lr;
      req.jlow_const = (uint32_t)key.rank << (lr - lp);
    } else {
      for (int i = 0; i < lc - lp; ++i) req.start[i] = 64 + i;
      for (int i = 0; i < lr - lp; ++i) req.start[lc - lp + i] = i;
      for (int i = 0; i < lp; ++i) req.start[lc - lp + lr - lp + i] = lr - lp + i;  // the old rank bits
      for (int i = 0; i < lc - lp; ++i) req.target[i] = 64 + i;
      for (int m = 0; m < lr; ++m) req.target[lc - lp + m] = 64 + lc + m;           // j1 natural
      req.time_hi = lr - 1;
      req.time_lo = 0;
      req.jlow_const = 0;
    }
  }
  return req;
}


std::string Plan::describe() const {
  std::ostringstream o;
  o << "{\"log_n\": " << log_n << ", \"tile_bits\": " << tile_bits << ", \"group_bits\": " << group_bits
    << ", \"coalesce_bits\": " << coalesce_bits
    << ", \"passes\": [";
  for (size_t i = 0; i < passes.size(); ++i) {
    const PassDesc& p = passes[i];
    auto arr8 = [&](const int8_t* a, int n) {
      std::ostringstream s;
      s << "[";
      for (int k = 0; k < n; ++k) s << (k ? ", " : "") << (int)a[k];
      s << "]";
      return s.str();
    };
    auto arr32 = [&](const uint32_t* a, int n) {
      std::ostringstream s;
      s << "[";
      for (int k = 0; k < n; ++k) s << (k ? ", " : "") << a[k];
      s << "]";
      return s.str();
    };
    o << (i ? ", " : "") << "{\"T\": " << p.T << ", \"d\": " << p.d << ", \"nblk_bits\": " << p.nblk_bits
      << ", \"nrounds\": " << p.nrounds << ", \"is_final\": " << p.is_final << ", \"has_tw\": " << p.has_tw
      << ", \"tw_lshift\": " << p.tw_lshift << ", \"scale_out\": " << p.scale_out
      << ", \"jlow_const\": " << p.jlow_const << ", \"in_pos\": " << arr8(p.in_pos, p.T)
      << ", \"out_pos\": " << arr8(p.out_pos, p.T) << ", \"blk_pos\": " << arr8(p.blk_pos, p.nblk_bits)
      << ", \"wt_loc\": " << arr32(p.wt_loc, p.T) << ", \"wt_blk\": " << arr32(p.wt_blk, p.nblk_bits)
      << ", \"round_sb\": " << arr8(p.round_sb, p.nrounds) << ", \"round_nb\": " << arr8(p.round_nb, p.nrounds)
      << ", \"tmap\": [";
    for (int r = 0; r < p.nrounds; ++r) o << (r ? ", " : "") << arr8(p.tmap[r], p.T - group_bits);
    o << "]}";
  }
  o << "]}";
  return o.str();
}

}  // namespace ntt256
