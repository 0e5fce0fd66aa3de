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


def run_pass(data: list[int], pd: dict, p: int, w_max: int, max_log_n: int, inverse: bool = False,
             scale: int | None = None, peer: dict | None = None) -> None:
    """In-place simulation of one pass over one transform (len(data) = 2^(T + nblk_bits))."""
    T, d = pd["T"], pd["d"]
    lowp = T - d
    nthreads = 1 << (T - 3)
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
                lam0 = _deposit(tid, pd["tmap"][r], T - 3)
                lams = [lam0 | (u << sb) for u in range(8)]
                if first:
                    x = [data[base + _deposit(l, pd["in_pos"], T)] for l in lams]
                else:
                    x = [tile[l] for l in lams]

                def bfly(i, j):
                    a, b = x[i], x[j]
                    x[i], x[j] = (a + b) % p, (a - b) % p

                w8 = lambda k: root(k * (N // 8))
                if nb == 3:
                    for u in range(4):
                        bfly(u, u + 4)
                        if u:
                            x[u + 4] = x[u + 4] * w8(u) % p
                    for h in (0, 4):
                        for u in range(2):
                            bfly(h + u, h + u + 2)
                            if u:
                                x[h + u + 2] = x[h + u + 2] * w8(2) % p
                    for u in range(0, 8, 2):
                        bfly(u, u + 1)
                elif nb == 2:
                    for u in range(4):
                        bfly(u, u + 4)
                        if u & 2:
                            x[u + 4] = x[u + 4] * w8(2) % p
                    for h in (0, 4):
                        for u in range(2):
                            bfly(h + u, h + u + 2)
                elif nb == 1:
                    for u in range(4):
                        bfly(u, u + 4)
                if not last:
                    pb = sb + 3 - nb
                    m = sb + 3 - lowp
                    sh = loc_bits - m
                    tl_mask = (1 << (pb - lowp)) - 1
                    tl0 = (lam0 >> lowp) & tl_mask
                    ush = sb - lowp
                    for u in range(8):
                        uproc = u >> (3 - nb)
                        v = _brev(uproc, nb)
                        rest = u & ((1 << (3 - nb)) - 1)
                        fold = scale is not None and pd.get("scaled_round", -1) == r
                        if uproc != 0 or fold:
                            tl = (tl0 | (rest << ush)) & tl_mask
                            e = ((tl * v) << sh) & ((1 << loc_bits) - 1)
                            x[u] = x[u] * root(e * (N >> loc_bits)) % p
                            if fold:
                                x[u] = x[u] * scale % p
                    for u in range(8):
                        tile[lams[u]] = x[u]
                else:
                    if pd["has_tw"]:
                        jl0 = jlow_blk + sum(((lam0 >> i) & 1) * pd["wt_loc"][i] for i in range(lowp))
                        gw = [pd["wt_loc"][sb + i] if sb + i < lowp else 0 for i in range(3)]
                        for u in range(8):
                            khi = _brev(lams[u] >> lowp, d) if d else 0
                            jl = jl0 + sum(gw[i] for i in range(3) if (u >> i) & 1)
                            e = ((jl * khi) << pd["tw_lshift"]) & (N - 1)
                            x[u] = x[u] * root(e) % p
                    if scale is not None and pd.get("scale_out"):
                        x = [v * scale % p for v in x]
                    for u in range(8):
                        stores.append((base + _deposit(lams[u], pd["out_pos"], T), x[u]))
            for pos, val in stores:
                if peer is None:
                    data[pos] = val
                else:  # fused exchange: store into the destination rank's receive buffer
                    dest = pos >> peer["shift"]
                    roff = (peer["rank"] << peer["shift"]) | (pos & ((1 << peer["shift"]) - 1))
                    peer["recv"][dest][roff] = val


def run_plan(data: list[int], plan: dict, p: int, g: int, max_log_n: int, inverse: bool = False) -> list[int]:
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
        run_pass(out, pd, p, w_max, max_log_n, inverse, scale)
    return out
