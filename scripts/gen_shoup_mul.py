#!/usr/bin/env python
"""Generator + limb-exact model of shoup_mul() (ntt_b200/csrc/field_shoup.cuh).

shoup_mul computes  r = x*w mod p  (lazy: r in [0, 2p + 2^192))  for a TABLE CONSTANT w whose companion
w' = floor(w * 2^256 / p) is stored next to it (Shoup / Barrett with a per-constant quotient):

    q  = floor( (sum over i+j >= 6 of x_i * w'_j * 2^(32(i+j))) / 2^256 )      43 wide MACs
    r  = low 256 bits of  x*w + q*(2^256 - p)                                  56 wide MACs + 16 low MACs
    if top 64 bits of r > top 64 bits of 2p: r -= p                             (2 compares + 8 predicated adds of 2^256 - p)

q is floor(x*w'/2^256) or one less (the dropped columns sum to < 2^227), so x*w - q*p lies in
[0, p*(2 + x/2^256)) -- below 3p < 2^256 whenever p < 2^254, which is what makes the low 256 bits the
true remainder.  Against the Montgomery product it replaces (128 wide MACs + 8 low ones,
/root/reference's CGBN mont_mul: include/cgbn/core/core_mont.cu:29-77) this is 99 wide + 16 low
multiplies: on B200 a 32x32->64 multiply (IMAD.WIDE / IMAD.HI) occupies the only pipe that has it
(fmaheavy) for 4 cycles per warp whatever its carry form, a low 32x32->32 IMAD for 2
(profiles/r02_microbench_issue.md), and that pipe is what binds the NTT kernels.

The instruction list below is built ONCE and used twice: emitted as CUDA (one PTX instruction per
`asm volatile` statement, the carry flag living across statements as in field_n.cuh) and executed by
a small interpreter with Python integers (tests/test_field_model.py), so the model IS the kernel's
arithmetic, limb for limb, carry for carry.

    python scripts/gen_shoup_mul.py > ntt_b200/csrc/field_shoup.cuh
"""
from __future__ import annotations

M32 = 0xFFFFFFFF


class Prog:
    """Straight-line program over u32 SSA temporaries with one carry flag."""

    def __init__(self):
        self.ops = []  # (opcode, dst, a, b, c)   operands: temp name | ('in', array, index) | 0
        self.n = 0
        self.may_drop = set()  # results whose carry out is discarded on purpose (top of a low product)

    def tmp(self):
        self.n += 1
        return f"t{self.n}"

    def emit(self, op, a, b, c=None):
        d = self.tmp()
        self.ops.append((op, d, a, b, c))
        return d


def X(i): return ("in", "x", i)
def W(i): return ("in", "w", i)
def WQ(i): return ("in", "wq", i)
def PB(i): return ("in", "f.pbar", i)
def P(i): return ("in", "f.p", i)
def P2(i): return ("in", "f.p2", i)


def _chain_rows(pr: Prog, acc: dict, mult, vec, i: int, lo_limit: int, top: int, trunc_top: bool):
    """acc[L] (absolute limb -> temp) += mult * vec[j] * 2^(32(i+j)) for the j of both parities,
    as two carry chains.  lo_limit: only terms with i+j >= lo_limit.  top: one past the highest
    limb kept (trunc_top: limbs >= top are dropped, as in a low product; else the accumulator simply
    grows and carries ripple into fresh limbs)."""
    for parity in (0, 1):
        js = [j for j in range(8) if (i + j) % 2 == parity and i + j >= lo_limit and (not trunc_top or i + j < top)]
        if not js:
            continue
        a = acc[parity]
        first = True
        for idx, j in enumerate(js):
            L = i + j
            last = idx == len(js) - 1
            lo_only = trunc_top and L == top - 1
            # low half
            addend = a.get(L, 0)
            if lo_only:
                assert last
                op = "mad.lo" if first else "madc.lo"
                a[L] = pr.emit(op, mult, vec(j), addend)
                pr.may_drop.add(a[L])
                first = False
                continue
            op = "mad.lo.cc" if first else "madc.lo.cc"
            a[L] = pr.emit(op, mult, vec(j), addend)
            first = False
            # high half
            hi_fresh = (L + 1) not in a
            addend = a.get(L + 1, 0)
            if last and trunc_top and L + 1 == top - 1:
                a[L + 1] = pr.emit("madc.hi", mult, vec(j), addend)  # carry out of the kept range is dropped
                pr.may_drop.add(a[L + 1])
            elif last and hi_fresh:
                a[L + 1] = pr.emit("madc.hi", mult, vec(j), addend)  # hi(x*w) + carry <= 2^32 - 1: cannot carry
            else:
                a[L + 1] = pr.emit("madc.hi.cc", mult, vec(j), addend)
                if last:
                    # ripple the carry upwards until a fresh limb absorbs it
                    k = L + 2
                    while True:
                        if trunc_top and k >= top:
                            raise AssertionError("chain should have ended at the top")
                        if k not in a:
                            a[k] = pr.emit("addc", 0, 0)
                            break
                        if k == top - 1 and not trunc_top:
                            a[k] = pr.emit("addc", a[k], 0)  # top limb of a 512-bit value: cannot carry out
                            break
                        a[k] = pr.emit("addc.cc", a[k], 0)
                        k += 1


def build() -> tuple[Prog, list]:
    pr = Prog()
    # ---- step 1: q = limbs 8..15 of the truncated product x * w'
    acc = ({}, {})  # even-aligned, odd-aligned accumulators: absolute limb -> temp
    for i in range(8):
        _chain_rows(pr, acc, X(i), WQ, i, lo_limit=6, top=16, trunc_top=False)
    E, O = acc
    q = []
    carry_started = False
    for L in range(7, 16):
        e, o = E.get(L, 0), O.get(L, 0)
        if not carry_started:
            t = pr.emit("add.cc", e, o)
            carry_started = True
        elif L == 15:
            t = pr.emit("addc", e, o)
        else:
            t = pr.emit("addc.cc", e, o)
        if L >= 8:
            q.append(t)
    # ---- step 2: r = low 256 bits of x*w + q*pbar
    acc = ({}, {})
    for i in range(8):
        _chain_rows(pr, acc, X(i), W, i, lo_limit=0, top=8, trunc_top=True)
        qi = q[i]
        _chain_rows(pr, acc, qi, PB, i, lo_limit=0, top=8, trunc_top=True)
    E, O = acc
    r = [E[0]]
    for L in range(1, 8):
        op = "add.cc" if L == 1 else ("addc" if L == 7 else "addc.cc")
        r.append(pr.emit(op, E[L], O[L]))
    pr.may_drop.add(r[7])
    return pr, r


# ------------------------------------------------------------------ interpreter (the limb-exact model)
def run(pr: Prog, outs, x: int, w: int, wq: int, p: int):
    ins = {
        "x": [(x >> (32 * i)) & M32 for i in range(8)],
        "w": [(w >> (32 * i)) & M32 for i in range(8)],
        "wq": [(wq >> (32 * i)) & M32 for i in range(8)],
        "f.pbar": [(((1 << 256) - p) >> (32 * i)) & M32 for i in range(8)],
        "f.p": [(p >> (32 * i)) & M32 for i in range(8)],
        "f.p2": [((2 * p) >> (32 * i)) & M32 for i in range(8)],
    }
    v = {}
    cf = 0

    def val(o):
        if o == 0 or o is None:
            return 0
        if isinstance(o, tuple):
            return ins[o[1]][o[2]]
        return v[o]

    for op, d, a, b, c in pr.ops:
        base = op.split(".cc")[0]
        cc = op.endswith(".cc")
        if base in ("mad.lo", "madc.lo"):
            t = ((val(a) * val(b)) & M32) + val(c) + (cf if base == "madc.lo" else 0)
        elif base in ("mad.hi", "madc.hi"):
            t = ((val(a) * val(b)) >> 32) + val(c) + (cf if base == "madc.hi" else 0)
        elif base in ("add", "addc"):
            t = val(a) + val(b) + (cf if base == "addc" else 0)
        else:
            raise ValueError(op)
        if cc:
            cf = t >> 32
        elif t >> 32:
            # an instruction without .cc whose result overflowed: legal only where the carry is
            # dropped on purpose (top of a low product).  The caller decides; record it.
            if d not in pr.may_drop:
                raise OverflowError(f"{op} -> {d} lost a carry")
        v[d] = t & M32
    return sum(v[o] << (32 * i) for i, o in enumerate(outs)), v


def shoup_mul_model(x: int, w: int, p: int):
    """Returns (r, r_before_correction, q): r = x*w mod p in [0, 2p + 2^192), computed limb by limb exactly
    as the CUDA code does."""
    wq = (w << 256) // p
    pr, outs = _PROG
    r0, v = run(pr, outs, x, w, wq, p)
    r = r0 - p if (r0 >> 192) > ((2 * p) >> 192) else r0  # step 3: the comparison only looks at the top 64 bits
    return r, r0, v


_PROG = build()


# ------------------------------------------------------------------ CUDA emitter
def _operand(o):
    if o == 0:
        return "0u"
    if isinstance(o, tuple):
        return f"{o[1]}[{o[2]}]"
    return o


_FN = {
    "mad.lo.cc": "mad_lo_cc", "madc.lo.cc": "madc_lo_cc", "madc.hi.cc": "madc_hi_cc", "madc.hi": "madc_hi",
    "mad.lo": "mad_lo", "madc.lo": "madc_lo", "add.cc": "add_cc", "addc.cc": "addc_cc", "addc": "addc",
}


def emit_cuda() -> str:
    pr, outs = _PROG
    n_wide = sum(1 for o in pr.ops if o[0].startswith("madc.hi"))
    n_lo_only = sum(1 for o in pr.ops if o[0] in ("mad.lo", "madc.lo"))
    lines = []
    w = lines.append
    w("// GENERATED by scripts/gen_shoup_mul.py -- do not edit; the same instruction list is the Python model")
    w("// that tests/test_field_model.py checks against big-integer arithmetic.")
    w("//")
    w("// shoup_mul(r, x, w, wq, f):  r = x*w mod p, r in [0, 2p + 2^192), for a table constant w < p with its")
    w("// companion wq = floor(w * 2^256 / p).  x is any 256-bit value.  Needs p < 2^254 (3p < 2^256).")
    w(f"// {n_wide} wide (32x32->64) multiplies + {n_lo_only} low ones, against 128 + 8 of the Montgomery product in")
    w("// field.cuh: the wide multiply is what the fmaheavy pipe of B200 is short of (4 cycles per warp).")
    w("#pragma once")
    w('#include "field_n.cuh"')
    w("")
    w("namespace ntt256 {")
    w("namespace fn {")
    w('NTT_DEV void mad_lo(u32& r, u32 a, u32 b, u32 c) { asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); }')
    w('NTT_DEV void madc_lo(u32& r, u32 a, u32 b, u32 c) { asm volatile("madc.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); }')
    w("}  // namespace fn")
    w("")
    w("NTT_DEV void shoup_mul(u32* r, const u32* x, const u32* w, const u32* wq, const FieldParams& f) {")
    w("  u32 " + ", ".join(f"t{i}" for i in range(1, pr.n + 1)) + ";")
    for op, d, a, b, c in pr.ops:
        fnname = _FN[op]
        if op.startswith("mad"):
            w(f"  fn::{fnname}({d}, {_operand(a)}, {_operand(b)}, {_operand(c if c is not None else 0)});")
        else:
            w(f"  fn::{fnname}({d}, {_operand(a)}, {_operand(b)});")
    w("  // step 3: r in [0, 3p) -> [0, 2p + 2^192): subtract p (add 2^256 - p) when the top 64 bits of r exceed those")
    w("  // of 2p -- then r >= 2p for certain; otherwise r < 2p + 2^192 (the wide-lazy range of field_wl.cuh)")
    w("  // The quotient estimate is one short at most except in a sliver of cases, so r < 2p for nearly every lane:")
    w("  // a warp in which no lane needs the correction skips its ten instructions (predicated-off instructions")
    w("  // still take issue slots, and issue slots are what this product is short of).")
    args_out = ", ".join(f'"+r"({outs[k]})' for k in range(8))
    args_in = ", ".join(f'"r"(f.pbar[{k}])' for k in range(8))
    w(f"  if (__any_sync(__activemask(), (((unsigned long long){outs[7]} << 32) | {outs[6]}) > (((unsigned long long)f.p2[7] << 32) | f.p2[6])))")
    w('  asm volatile("{\\n\\t.reg .pred q;\\n\\t.reg .u64 a, b;\\n\\t"')
    w('      "mov.b64 a, {%6, %7};\\n\\t"')
    w('      "mov.b64 b, {%16, %17};\\n\\t"')
    w('      "setp.gt.u64 q, a, b;\\n\\t"')
    for k in range(8):
        ins = "add.cc.u32 " if k == 0 else ("addc.u32   " if k == 7 else "addc.cc.u32")
        w(f'      "@q {ins} %{k}, %{k}, %{8 + k};\\n\\t"')
    w('      "}"')
    w(f"      : {args_out}")
    w(f'      : {args_in}, "r"(f.p2[6]), "r"(f.p2[7]));')
    for k in range(8):
        w(f"  r[{k}] = {outs[k]};")
    w("}")
    w("")
    w("}  // namespace ntt256")
    return "\n".join(lines) + "\n"


if __name__ == "__main__":
    print(emit_cuda(), end="")
