"""Limb-exact Python model of mont_mul() in ntt_b200/csrc/field.cuh (even/odd CIOS on 32-bit
limbs).  Checks the value identity r = a*b*2^-256 mod p with r < 2p and, above all, the two
claims the CUDA code relies on but cannot assert at run time: the odd accumulator chain never
carries out of limb 7, and Y[7] never overflows when the even chain's carry is folded in."""
import random

import pytest

from oracle import py_oracle as po

M32 = 0xFFFFFFFF


def _limbs(v, nl=8):
    return [(v >> (32 * i)) & M32 for i in range(nl)]


def _val(l):
    return sum(x << (32 * i) for i, x in enumerate(l))


class Overflow(Exception):
    pass


def _mad_row(acc, a4, b):
    """acc[0..NL) += (a4[0..NL/2)) * b laid at limb pairs; returns carry out (as mad_row in field.cuh)."""
    c = 0
    for j in range(len(a4)):
        prod = a4[j] * b
        t = acc[2 * j] + (prod & M32) + c
        acc[2 * j] = t & M32
        c = t >> 32
        t = acc[2 * j + 1] + (prod >> 32) + c
        acc[2 * j + 1] = t & M32
        c = t >> 32
    return c


def mont_mul_model(a, b, p, np0, NL=8):
    """NL = 8: mont_mul in field.cuh; NL = 12: mont_mul_n<12> in field_n.cuh (same chains)."""
    H = NL // 2
    al, bl, pl = _limbs(a, NL), _limbs(b, NL), _limbs(p, NL)
    ae, ao = al[0::2], al[1::2]
    pe, podd = pl[0::2], pl[1::2]
    X = [0] * NL
    Y = [0] * NL
    for j in range(H):
        pr = ae[j] * bl[0]
        X[2 * j], X[2 * j + 1] = pr & M32, pr >> 32
        pr = ao[j] * bl[0]
        Y[2 * j], Y[2 * j + 1] = pr & M32, pr >> 32

    def redc():
        m = (X[0] * np0) & M32
        cx = _mad_row(X, pe, m)
        cy = _mad_row(Y, podd, m)
        if cy:
            raise Overflow("odd chain carried out")
        if Y[NL - 1] + cx > M32:
            raise Overflow("Y top limb overflow on carry fold")
        Y[NL - 1] += cx
        assert X[0] == 0

    redc()
    for i in range(1, NL):
        # shift_mad_row
        t = Y[0] + X[1]
        nX0, c = t & M32, t >> 32
        nY = X[2:NL] + [0, 0]
        # nY += a_odd * b_i with carry-in c
        cc = c
        for j in range(H):
            pr = ao[j] * bl[i]
            t = nY[2 * j] + (pr & M32) + cc
            nY[2 * j], cc = t & M32, t >> 32
            t = nY[2 * j + 1] + (pr >> 32) + cc
            nY[2 * j + 1], cc = t & M32, t >> 32
        if cc:
            raise Overflow("nY chain carried out")
        nX = [nX0] + Y[1:NL]
        cx = _mad_row(nX, ae, bl[i])
        if nY[NL - 1] + cx > M32:
            raise Overflow("nY top limb overflow")
        nY[NL - 1] += cx
        X[:] = nX
        Y[:] = nY
        redc()
    r = _val(Y) + (_val(X) >> 32)
    if r >> (32 * NL):
        raise Overflow("result exceeds the container")
    return r


FIELDS = [po.BN254_FR, po.BLS12_381_FR, po.REF29, po.Field("p255m19", (1 << 255) - 19, 2)]


@pytest.mark.parametrize("F", FIELDS, ids=lambda f: f.name)
def test_mont_mul_model_value_and_no_overflow(F):
    p = F.p
    np0 = (-pow(p, -1, 1 << 32)) % (1 << 32)
    Rinv = pow(1 << 256, -1, p)
    rng = random.Random(12345)
    two_p = 2 * p
    specials_a = [0, 1, p - 1, p - 2, (p - 1) // 2, p >> 1, 0xFFFFFFFF, p & ~M32]
    specials_b = [0, 1, p - 1, p, two_p - 1, two_p - 2, (1 << 255) - 1 if two_p > (1 << 255) else two_p - 1,
                  min(two_p - 1, (1 << 256) - 1), _val([M32] * 7 + [_limbs(two_p - 1)[7]]) % two_p]
    cases = [(a, b) for a in specials_a for b in specials_b]
    cases += [(rng.randrange(p), rng.randrange(two_p)) for _ in range(3000)]
    # adversarial: limbs all-ones where allowed
    for _ in range(500):
        a = _val([rng.choice([0, M32, rng.getrandbits(32)]) for _ in range(8)]) % p
        b = _val([rng.choice([0, M32, rng.getrandbits(32)]) for _ in range(8)]) % two_p
        cases.append((a, b))
    for a, b in cases:
        r = mont_mul_model(a, b, p, np0)  # raises Overflow if a chain carries out
        assert r < two_p
        assert r % p == a * b * Rinv % p


@pytest.mark.parametrize("F", [po.BLS12_377_FQ, po.Field("p383m187", (1 << 383) - 187, 2)], ids=lambda f: f.name)
def test_mont_mul_model_12_limbs(F):
    """field_n.cuh at NL = 12 (384-bit containers): value identity and the no-carry-out claims."""
    p, NL = F.p, 12
    np0 = (-pow(p, -1, 1 << 32)) % (1 << 32)
    Rinv = pow(1 << (32 * NL), -1, p)
    rng = random.Random(384)
    two_p = 2 * p
    cap = 1 << (32 * NL)
    cases = [(a, b) for a in (0, 1, p - 1, p - 2, p >> 1, M32) for b in (0, 1, p - 1, p, two_p - 1, min(two_p, cap) - 1)]
    cases += [(rng.randrange(p), rng.randrange(two_p)) for _ in range(1500)]
    for _ in range(300):
        a = _val([rng.choice([0, M32, rng.getrandbits(32)]) for _ in range(NL)]) % p
        b = _val([rng.choice([0, M32, rng.getrandbits(32)]) for _ in range(NL)]) % two_p
        cases.append((a, b))
    for a, b in cases:
        r = mont_mul_model(a, b, p, np0, NL)
        assert r < two_p
        assert r % p == a * b * Rinv % p


def test_lazy_add_sub_window():
    """add_lazy/sub_lazy keep values in [0, 2p); BLS12-381 Fr needs the 2^256 carry (4p > 2^256)."""
    F = po.BLS12_381_FR
    p = F.p
    assert 4 * p > 1 << 256 > 2 * p
    rng = random.Random(7)
    for _ in range(2000):
        a, b = rng.randrange(2 * p), rng.randrange(2 * p)
        t = a + b
        carry, t256 = t >> 256, t & ((1 << 256) - 1)
        d = (t256 - 2 * p) & ((1 << 256) - 1)
        borrow = t256 < 2 * p
        s = t256 if (borrow and not carry) else d
        assert s < 2 * p and s % p == (a + b) % p
        diff = (a - b) & ((1 << 256) - 1)
        if a < b:
            diff = (diff + 2 * p) & ((1 << 256) - 1)
        assert diff < 2 * p and diff % p == (a - b) % p


# --------------------------------------------------------------------------- shoup_mul (field_shoup.cuh)
def _shoup():
    import importlib.util
    import os

    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scripts", "gen_shoup_mul.py")
    spec = importlib.util.spec_from_file_location("gen_shoup_mul", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_shoup_mul_header_is_what_the_generator_emits():
    """ntt_b200/csrc/field_shoup.cuh is generated; the instruction list it was emitted from is the model
    the tests below execute, so the committed header must be the generator's current output."""
    import os

    g = _shoup()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with open(os.path.join(root, "ntt_b200", "csrc", "field_shoup.cuh")) as f:
        assert f.read() == g.emit_cuda()


def test_shoup_mul_multiply_counts():
    g = _shoup()
    pr, _ = g._PROG
    wide = sum(1 for o in pr.ops if o[0].startswith("madc.hi"))
    low = sum(1 for o in pr.ops if o[0] in ("mad.lo", "madc.lo"))
    assert (wide, low) == (99, 16)  # against 128 + 8 of the Montgomery product


@pytest.mark.parametrize("p", [po.BN254_FR.p, po.REF29.p, (1 << 254) - 127 * 2 + 1])
def test_shoup_mul_model_is_exact_and_never_loses_a_carry(p):
    """r = x*w mod p with r < 2p + 2^192 for every x < 2^256 and w < p, the uncorrected remainder below 3p, and no
    instruction without a carry-out ever overflows except where the generator drops the carry on purpose
    (the interpreter raises OverflowError otherwise)."""
    g = _shoup()
    if p % 2 == 0:
        p += 1
    rng = random.Random(20262)
    worst = 0.0
    specials_x = [0, 1, p - 1, p, 2 * p - 1, 2 * p, 3 * p, (1 << 256) - 1, (1 << 256) - p, 1 << 255, (1 << 224) - 1]
    specials_w = [0, 1, 2, p - 1, p - 2, p // 2, p // 2 + 1, (p - 1) // 3]
    cases = [(x, w) for x in specials_x for w in specials_w]
    for i in range(6000):
        kind = i % 4
        x = (rng.randrange(1 << 256), rng.randrange(2 * p), (1 << 256) - 1 - rng.randrange(1 << 64),
             rng.randrange(1 << 256) | (M32 << 224))[kind]
        cases.append((x % (1 << 256), rng.randrange(p)))
    for x, w in cases:
        r, r0, _ = g.shoup_mul_model(x, w, p)
        assert r % p == (x * w) % p
        assert r < 2 * p + (1 << 192) and r0 < 3 * p
        assert r >= p or r == r0  # a correction only ever fires on values of at least 2p
        worst = max(worst, r0 / p)
    assert worst < 3.0


def test_shoup_quotient_is_at_most_one_below_the_true_floor():
    g = _shoup()
    p = po.BN254_FR.p
    rng = random.Random(7)
    for _ in range(2000):
        x, w = rng.randrange(1 << 256), rng.randrange(p)
        wq = (w << 256) // p
        _, r0, _ = g.shoup_mul_model(x, w, p)
        q = (x * w - r0) // p
        assert (x * w - r0) % p == 0
        assert ((x * wq) >> 256) - q in (0, 1)


# --------------------------------------------------------------------------- wide-lazy add/sub (field_wl.cuh)
def _plan(log_n, max_log_n, tile_bits, coalesce_bits, group_bits=3):
    from ntt_b200.lib import plan_describe

    return plan_describe(log_n, max_log_n, tile_bits, coalesce_bits, group_bits)


def test_wide_lazy_ops_at_the_extremes_of_their_ranges():
    """add_wl / sub_wl / sub_wl_mul / canonicalize_wl on every combination of the extreme values a register can
    hold between butterflies: 0, around p, around 2p, around the comparison threshold and the largest excess the
    policy allows (2^224).  The model asserts the no-overflow / no-underflow claims; here the residues are
    checked and that the results are back in range."""
    from wl_model import WideLazyArith

    p = po.BN254_FR.p
    A = WideLazyArith(p)
    thr = ((2 * p) >> 192) << 192
    vals = sorted({0, 1, p - 1, p, p + 1, 2 * p - 1, 2 * p, 2 * p + 1, thr - 1, thr, thr + (1 << 192) - 1,
                   thr + (1 << 192), 2 * p + (1 << 192), 2 * p + (1 << 224) - 1})
    for a in vals:
        for b in vals:
            s = A.add(a, b)
            assert s % p == (a + b) % p and s < 2 * p + max(1 << 192, max(0, a - 2 * p) + max(0, b - 2 * p)) + 1
            d = A.sub(a, b, False)
            assert d % p == (a - b) % p and d < 2 * p + max(1 << 192, a - 2 * p + 1)
            m = A.sub(a, b, True)
            assert m % p == (a - b) % p and 0 < m < 5 * p + (1 << 224)
            for x in (s, d, m):
                assert A.canon(x) == x % p
    assert A.max_excess < (1 << 225)


@pytest.mark.parametrize("log_n,tile_bits,group_bits", [(6, 6, 3), (9, 6, 3), (10, 10, 3), (8, 8, 2), (7, 7, 1)])
def test_whole_transforms_through_the_wide_lazy_model(log_n, tile_bits, group_bits):
    """The kernel model (plan_sim) executes real plans with the limb-level wide-lazy arithmetic, products by the
    generated shoup_mul instruction list: the outputs are the oracle's, every intermediate stays in range (the
    model asserts it) and the excess over 2p stays tiny.  Forward and inverse, random and worst-case inputs."""
    import plan_sim
    from wl_model import WideLazyArith

    f = po.BN254_FR
    p, g = f.p, f.g
    n = 1 << log_n
    plan = _plan(log_n, 16, tile_bits, min(2, tile_bits), group_bits)
    rng = random.Random(31 * log_n + group_bits)
    for x in ([rng.randrange(p) for _ in range(n)], [p - 1] * n, [(p - 1) if i % 2 else 0 for i in range(n)]):
        A = WideLazyArith(p)
        got = plan_sim.run_plan(x, plan, p, g, 16, arith=A)
        assert got == po.ntt_ref(x, f)
        assert A.max_excess < (1 << 192) << (log_n + 1)
        assert A.max_operand < 5 * p + (1 << 224)
        B = WideLazyArith(p)
        back = plan_sim.run_plan(got, plan, p, g, 16, inverse=True, arith=B)
        assert back == x
    assert A.n_ops["sub_mul"] > 0 and A.n_ops["mul"] > 0


def test_a_pass_started_from_the_edge_of_the_lazy_range():
    """Non-final passes hand lazy values to the next launch.  One pass of a two-pass plan run on inputs that sit
    at the top of the range (2p + 2^224 - 1 and friends): residues agree with exact arithmetic and nothing
    overflows."""
    import plan_sim
    from wl_model import WideLazyArith

    f = po.BN254_FR
    p, g = f.p, f.g
    plan = _plan(10, 16, 6, 2)
    assert len(plan["passes"]) >= 2
    w_max = pow(g, (p - 1) >> 16, p)
    edge = [2 * p + (1 << 224) - 1, 2 * p + (1 << 192), 2 * p - 1, 0, p, ((2 * p) >> 192) << 192]
    rng = random.Random(5)
    data = [edge[rng.randrange(len(edge))] for _ in range(1 << 10)]
    for pd in plan["passes"][:-1]:
        a, b = list(data), [v % p for v in data]
        A = WideLazyArith(p, exact_products=False)
        plan_sim.run_pass(a, pd, p, w_max, 16, arith=A)
        plan_sim.run_pass(b, pd, p, w_max, 16)
        assert [v % p for v in a] == b
        assert A.max_excess < (1 << 226)


# ---------------------------------------------------------------- 64-bit fields (ntt_b200/csrc/field64.cuh)
GL_P = 0xFFFFFFFF00000001
M32 = 0xFFFFFFFF


def _gl_sub(a, b):
    """gl_sub: 64-bit subtract, then subtract EPS when it borrowed (canonical a, b)."""
    d = (a - b) & (2**64 - 1)
    if a < b:
        assert d >= M32  # the correction cannot borrow again
        d -= M32
    return d


def _gl_reduce96(t, x2):
    """gl_reduce96: t + x2 * EPS for any 64-bit t, +EPS on carry, one conditional subtraction of p."""
    r = t + x2 * M32
    if r >> 64:
        r = (r & (2**64 - 1)) + M32
        assert r < 2**64  # cannot carry again
    r0, r1 = r & M32, r >> 32
    if r1 == M32 and r0 != 0:
        r0, r1 = r0 - 1, 0
    return (r1 << 32) | r0


def _gl_reduce128(x0, x1, x2, x3):
    t = ((x1 << 32) | x0) - x3
    if t < 0:
        t += 2**64 - M32
        assert t >= 0
    return _gl_reduce96(t, x2)


def _gl_mul(a, b):
    x = a * b
    return _gl_reduce128(x & M32, (x >> 32) & M32, (x >> 64) & M32, x >> 96)


def _gl_mul_pow2(x, S):
    q, r = divmod(S, 32)
    assert 0 < S < 96 and r
    x0, x1 = x & M32, x >> 32
    z0, z1, z2 = (x0 << r) & M32, ((x1 << r) | (x0 >> (32 - r))) & M32, x1 >> (32 - r)
    if q == 0:
        return _gl_reduce96((z1 << 32) | z0, z2)
    if q == 1:
        return _gl_reduce128(0, z0, z1, z2)
    ze = (z0 << 32) - z0
    assert ze < GL_P and ((z2 << 32) | z1) < GL_P
    return _gl_sub(ze, (z2 << 32) | z1)


def test_goldilocks_limb_model():
    """The limb-level steps of gl_mul / gl_mul_pow2 / gl_sub, with the no-second-carry claims asserted inside, on
    extreme and random canonical operands; and the exponents the kernels use for the constant twiddles."""
    rng = random.Random(64)
    edge = [0, 1, 2, M32 - 1, M32, M32 + 1, 2**32 + 1, 2**63, GL_P - 1, GL_P - 2, GL_P - M32, GL_P - M32 - 1, 0xFFFFFFFE00000001,
            0xFFFFFFFF00000000, 0x00000000FFFFFFFF, 0xFFFFFFFEFFFFFFFF]
    vals = edge + [rng.randrange(GL_P) for _ in range(300)]
    for a in vals:
        for b in edge + [rng.randrange(GL_P) for _ in range(8)]:
            assert _gl_mul(a, b) == a * b % GL_P
            assert _gl_sub(a, b) == (a - b) % GL_P
            assert _gl_sub(a, GL_P - b) == (a + b) % GL_P  # gl_add
        for S in (12, 24, 36, 48, 60, 72, 84, 4, 31, 33, 63, 65, 95):
            assert _gl_mul_pow2(a, S) == a * pow(2, S, GL_P) % GL_P
    # w_16 = g^((p-1)/16) for g = 7 is 2^(12*13), its inverse 2^(12*3): the K16 values the kernels are compiled for
    w16 = pow(7, (GL_P - 1) // 16, GL_P)
    assert w16 == pow(2, 12 * 13, GL_P) and pow(w16, GL_P - 2, GL_P) == pow(2, 12 * 3, GL_P)
    assert pow(2, 96, GL_P) == GL_P - 1
    for K in (13, 3):
        for e in range(1, 8):
            s192 = 12 * K * e % 192
            assert s192 % 96 and (s192 % 96) % 32
