"""CPU model of ntt_pass_kernel (ntt_b200/csrc/ntt_kernels.cuh) driven by the plan JSON that the
C++ planner emits.  Test infrastructure: it lets the planner, the descriptor semantics and the
twiddle-exponent formulas be checked against the oracle without a GPU."""
from __future__ import annotations


def _deposit(v: int, pos: list[int], n: int) -> int:
    r = 0
    for i in range(n):
        r |= ((v >> i) & 1) << pos[i]
    return r


def _brev(v: int, bits: int) -> int:
    r = 0
    for i in range(bits):
        r |= ((v >> i) & 1) << (bits - 1 - i)
    return r


class ExactArith:
    """Field operations of the model: exact residues.  tests/wl_model.py substitutes the limb-level
    wide-lazy policy of Shoup contexts (field_wl.cuh + field_shoup.cuh) to check its ranges."""

    def __init__(self, p: int):
        self.p = p

    def add(self, a, b):
        return (a + b) % self.p

    def sub(self, a, b, to_mul):
        return (a - b) % self.p

    def mul(self, x, w):
        return x * w % self.p

    def canon(self, x):
        return x % self.p


def run_pass(data: list[int], pd: dict, p: int, w_max: int, max_log_n: int, inverse: bool = False,
             scale: int | None = None, peer: dict | None = None, group_bits: int = 3, arith=None) -> None:
    """In-place simulation of one pass over one transform (len(data) = 2^(T + nblk_bits))."""
    A = arith or ExactArith(p)
    T, d = pd["T"], pd["d"]
    lowp = T - d
    GB = group_bits
    NE = 1 << GB
    nthreads = 1 << (T - GB)
    N = 1 << max_log_n

    def root(e: int) -> int:  # w_N^e with the kernel's inverse convention
        e %= N
        if inverse:
            e = (-e) % N
        return pow(w_max, e, p)

    loc_bits = min(12, max_log_n)

    for blk in range(1 << pd["nblk_bits"]):
        base = _deposit(blk, pd["blk_pos"], pd["nblk_bits"])
        jlow_blk = pd["jlow_const"] + sum(((blk >> i) & 1) * pd["wt_blk"][i] for i in range(pd["nblk_bits"]))
        tile = [None] * (1 << T)
        for r in range(pd["nrounds"]):
            sb, nb = pd["round_sb"][r], pd["round_nb"][r]
            first, last = r == 0, r == pd["nrounds"] - 1
            stores = []  # global stores happen after every thread of the CTA has loaded
            for tid in range(nthreads):
                lam0 = _deposit(tid, pd["tmap"][r], T - GB)
                lams = [lam0 | (u << sb) for u in range(NE)]
                if first:
                    x = [data[base + _deposit(l, pd["in_pos"], T)] for l in lams]
                else:
                    x = [tile[l] for l in lams]

                def bfly(i, j, to_mul):
                    a, b = x[i], x[j]
                    x[i], x[j] = A.add(a, b), A.sub(a, b, to_mul)

                # what the kernel's tail_mul says: every element with a processed bit set meets a product next
                tail_mul = (not last) or bool(pd["has_tw"]) or (scale is not None and bool(pd.get("scale_out")))

                MB = max(3, GB)  # constant twiddles are powers of w_8 (w_16 in the sixteen-element 64-bit kernels)
                w8 = lambda k: root(k * N // (1 << MB))
                for k in range(nb):  # DIF stages on the top nb bits of the group, top bit first
                    bit = GB - 1 - k
                    for u in range(NE):
                        if u & (1 << bit):
                            continue
                        t = (u >> (GB - nb)) & ((1 << (nb - 1 - k)) - 1)
                        e8 = t << (MB - (nb - k))
                        bfly(u, u | (1 << bit), e8 != 0 or (k == nb - 1 and tail_mul))
                        if e8:
                            x[u | (1 << bit)] = A.mul(x[u | (1 << bit)], w8(e8))
                if not last:
                    pb = sb + GB - nb
                    m = sb + GB - lowp
                    sh = loc_bits - m
                    tl_mask = (1 << (pb - lowp)) - 1
                    tl0 = (lam0 >> lowp) & tl_mask
                    ush = sb - lowp
                    for u in range(NE):
                        uproc = u >> (GB - nb)
                        v = _brev(uproc, nb)
                        rest = u & ((1 << (GB - nb)) - 1)
                        fold = scale is not None and pd.get("scaled_round", -1) == r
                        if uproc != 0 or fold:
                            tl = (tl0 | (rest << ush)) & tl_mask
                            e = ((tl * v) << sh) & ((1 << loc_bits) - 1)
                            w = root(e * (N >> loc_bits))
                            x[u] = A.mul(x[u], w * scale % p if fold else w)
                    for u in range(NE):
                        tile[lams[u]] = x[u]
                else:
                    if pd["has_tw"]:
                        jl0 = jlow_blk + sum(((lam0 >> i) & 1) * pd["wt_loc"][i] for i in range(lowp))
                        gw = [pd["wt_loc"][sb + i] if sb + i < lowp else 0 for i in range(GB)]
                        for u in range(NE):
                            khi = _brev(lams[u] >> lowp, d) if d else 0
                            jl = jl0 + sum(gw[i] for i in range(GB) if (u >> i) & 1)
                            e = ((jl * khi) << pd["tw_lshift"]) & (N - 1)
                            x[u] = A.mul(x[u], root(e))
                    if scale is not None and pd.get("scale_out"):
                        x = [A.mul(v, scale) for v in x]
                    if pd.get("is_final"):
                        x = [A.canon(v) for v in x]
                    for u in range(NE):
                        stores.append((base + _deposit(lams[u], pd["out_pos"], T), x[u]))
            for pos, val in stores:
                if peer is None:
                    data[pos] = val
                else:  # fused exchange: store into the destination rank's receive buffer
                    dest = pos >> peer["shift"]
                    roff = (peer["rank"] << peer["shift"]) | (pos & ((1 << peer["shift"]) - 1))
                    peer["recv"][dest][roff] = val


def run_plan(data: list[int], plan: dict, p: int, g: int, max_log_n: int, inverse: bool = False, arith=None) -> list[int]:
    out = list(data)
    w_max = pow(g, (p - 1) >> max_log_n, p)
    n = len(out)
    passes = plan["passes"]
    for i, pd in enumerate(passes):
        pd = dict(pd)
        scale = None
        if inverse and i == len(passes) - 1:
            scale = pow(n, p - 2, p)
            if pd["nrounds"] >= 2:
                pd["scaled_round"] = 0  # 1/n folded into the first round's twiddle table
            else:
                pd["scale_out"] = 1
        run_pass(out, pd, p, w_max, max_log_n, inverse, scale, group_bits=plan.get("group_bits", 3), arith=arith)
    return out
