"""Parity tests of the 64-bit transforms (include/ntt64.h) on the GPU, through the C ABI, bit-exact against
(a) the outputs of the reference's own CPU NTT() over its own modulus (tests/golden/ref_i64_kat.json and, when
oracle/_ref exists, the compiled reference run live) and (b) the C / Python oracles over Goldilocks."""
import ctypes
import hashlib
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GL_P, GL_G = 0xFFFFFFFF00000001, 7
REF_P, REF_G = 469762049, 3


@pytest.fixture(scope="module")
def torch_cuda():
    import torch

    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device; there is no CPU fallback to test")
    torch.cuda.set_device(0)
    return torch


@pytest.fixture(scope="module")
def ctxs(torch_cuda):
    from ntt_b200.api64 import Ntt64

    made = {}

    def get(field, max_log_n):
        key = (field, max_log_n)
        if key not in made:
            made[key] = Ntt64(field, max_log_n=max_log_n)
        return made[key]

    yield get
    for c in made.values():
        c.close()


def _run(torch, fn, x, *args, **kw):
    d = torch.from_numpy(np.array(x, dtype=np.uint64, copy=True).view(np.int64)).cuda()
    fn(d, *args, **kw)
    torch.cuda.synchronize()
    return d.cpu().numpy().view(np.uint64)


def test_reference_golden_vectors_on_the_gpu(torch_cuda, ctxs):
    """x[j] = j over p = 469762049: the GPU output has the digest of the reference's own NTT() output."""
    with open(os.path.join(GOLD, "ref_i64_kat.json")) as f:
        g = json.load(f)
    ctx = ctxs("ref29", 20)
    for case in g["cases"]:
        n = 1 << case["log_n"]
        X = _run(torch_cuda, ctx.forward, np.arange(n, dtype=np.uint64), case["log_n"])
        assert X[:8].tolist() == case["head"][: min(8, n)]
        assert hashlib.sha256(X.astype("<i8").tobytes()).hexdigest() == case["sha256_i64le"], case["log_n"]
        assert np.array_equal(_run(torch_cuda, ctx.inverse, X, case["log_n"]), np.arange(n, dtype=np.uint64))


def test_live_against_the_compiled_reference(torch_cuda, ctxs):
    from oracle import c_oracle as co

    R = co.ref_lib()
    if R is None:
        pytest.skip("oracle/_ref not built (no /root/reference on this box)")
    i64p = ctypes.POINTER(ctypes.c_longlong)
    ctx = ctxs("ref29", 20)
    rng = np.random.default_rng(11)
    for log_n in (1, 2, 3, 6, 11, 14, 17):
        x = rng.integers(0, REF_P, size=1 << log_n, dtype=np.int64)
        a = x.copy()
        R.ref_ntt_i64(a.ctypes.data_as(i64p), log_n)
        got = _run(torch_cuda, ctx.forward, x.astype(np.uint64), log_n)
        assert np.array_equal(got.astype(np.int64), a), log_n


@pytest.mark.parametrize("field,p,g", [("goldilocks", GL_P, GL_G), ("ref29", REF_P, REF_G)])
def test_all_sizes_vs_oracle(torch_cuda, ctxs, field, p, g):
    from oracle import c_oracle as co

    ctx = ctxs(field, 20)
    for log_n in range(0, 19):
        x = co.splitmix64_elements(1 << log_n, p, 100 + log_n)
        if field == "goldilocks" and log_n >= 2:
            x[:4] = np.array([p - 1, 0, 0xFFFFFFFF, p - 0xFFFFFFFF], dtype=np.uint64)  # extreme residues
        want = co.ntt64(x, p, g, threads=8)
        got = _run(torch_cuda, ctx.forward, x, log_n)
        assert np.array_equal(got, want), (field, log_n)
        assert np.array_equal(_run(torch_cuda, ctx.inverse, want, log_n), x), (field, log_n)


@pytest.mark.parametrize("field,p,g", [("goldilocks", GL_P, GL_G), ("ref29", REF_P, REF_G)])
def test_batches(torch_cuda, ctxs, field, p, g):
    from oracle import c_oracle as co

    ctx = ctxs(field, 20)
    for log_n, batch in ((1, 33), (2, 257), (3, 1000), (5, 70000), (8, 300), (12, 64), (16, 9)):
        x = co.splitmix64_elements(batch << log_n, p, 55 + log_n)
        got = _run(torch_cuda, ctx.forward, x, log_n, batch)
        xs = x.reshape(batch, -1)
        for b in sorted({0, 1, batch // 2, batch - 1}):
            assert np.array_equal(got.reshape(batch, -1)[b], co.ntt64(xs[b], p, g)), (log_n, batch, b)
        assert np.array_equal(_run(torch_cuda, ctx.inverse, got, log_n, batch), x), (log_n, batch)


def test_large_goldilocks_full_compare(torch_cuda):
    """2^22 and 2^24 points, every output against the multi-threaded C oracle; 2^24 is also the bench size."""
    from ntt_b200.api64 import Ntt64
    from oracle import c_oracle as co

    ctx = Ntt64("goldilocks", max_log_n=24)
    for log_n in (20, 22, 24):
        x = co.splitmix64_elements(1 << log_n, GL_P, 4242 + log_n)
        want = co.ntt64(x, GL_P, GL_G, threads=co.max_threads())
        got = _run(torch_cuda, ctx.forward, x, log_n)
        assert np.array_equal(got, want), log_n
        assert np.array_equal(_run(torch_cuda, ctx.inverse, got, log_n), x), log_n
    ctx.close()


def test_tile_and_coalescing_variants_agree(torch_cuda):
    from ntt_b200.api64 import Ntt64
    from oracle import c_oracle as co

    x = co.splitmix64_elements(1 << 18, GL_P, 9)
    want = co.ntt64(x, GL_P, GL_G, threads=8)
    for T, g in ((12, 4), (14, 2), (13, 3), (10, 4), (8, 2), (12, 0), (6, 1)):
        ctx = Ntt64("goldilocks", max_log_n=20)
        ctx.set_tuning(T, g)
        assert np.array_equal(_run(torch_cuda, ctx.forward, x, 18), want), (T, g)
        assert np.array_equal(_run(torch_cuda, ctx.inverse, want, 18), x), (T, g)
        ctx.close()


def test_shift_and_multiplier_twiddles_agree(torch_cuda):
    """Goldilocks: the constant twiddles of the register rounds by shifts (default) and on the multiplier give
    the same bits, forward and inverse, for every round shape (sizes 2^3 .. 2^17) and a non-standard generator
    (whose w_16 the shift kernels are not compiled for) still transforms correctly."""
    from ntt_b200.api64 import Ntt64
    from oracle import c_oracle as co

    ctx = Ntt64("goldilocks", max_log_n=18)
    for log_n in range(3, 18):
        x = co.splitmix64_elements(1 << log_n, GL_P, 500 + log_n)
        want = co.ntt64(x, GL_P, GL_G, threads=4)
        for enable in (True, False):
            ctx.set_shift_twiddles(enable)
            assert np.array_equal(_run(torch_cuda, ctx.forward, x, log_n), want), (log_n, enable)
            assert np.array_equal(_run(torch_cuda, ctx.inverse, want, log_n), x), (log_n, enable)
    ctx.close()
    g2 = pow(GL_G, 5, GL_P)  # another generator
    ctx = Ntt64(p=GL_P, g=g2, max_log_n=14)
    x = co.splitmix64_elements(1 << 13, GL_P, 5)
    assert np.array_equal(_run(torch_cuda, ctx.forward, x, 13), co.ntt64(x, GL_P, g2))
    ctx.close()


def test_max_two_adicity_context(torch_cuda):
    """A context for 2^32 points (Goldilocks' full two-adicity) transforms small sizes with the same bits."""
    from ntt_b200.api64 import Ntt64
    from oracle import c_oracle as co

    ctx = Ntt64("goldilocks", max_log_n=32)
    for log_n in (3, 10, 16, 21):
        x = co.splitmix64_elements(1 << log_n, GL_P, log_n)
        assert np.array_equal(_run(torch_cuda, ctx.forward, x, log_n), co.ntt64(x, GL_P, GL_G, threads=8)), log_n
    ctx.close()


@pytest.mark.parametrize("field,p,g", [("goldilocks", GL_P, GL_G), ("ref29", REF_P, REF_G)])
def test_coset_and_pointwise(torch_cuda, ctxs, field, p, g):
    """coset_forward = NTT of x[j] * s^j; coset_inverse undoes it; vec_mul is the plain product; together they
    multiply two polynomials (cyclic convolution checked against Python integers)."""
    from oracle import c_oracle as co

    ctx = ctxs(field, 20)
    s = 7 if field == "ref29" else 0x123456789ABCDEF
    for log_n, batch in ((1, 1), (2, 3), (4, 2), (10, 1), (13, 3), (17, 1)):
        n = 1 << log_n
        x = co.splitmix64_elements(batch << log_n, p, 300 + log_n)
        pre = np.array([(int(v) * pow(s, j % n, p)) % p for j, v in enumerate(x)], dtype=np.uint64) if n <= (1 << 13) else None
        got = _run(torch_cuda, ctx.coset_forward, x, log_n, s, batch)
        if pre is not None:
            for b in range(batch):
                assert np.array_equal(got.reshape(batch, -1)[b], co.ntt64(pre.reshape(batch, -1)[b], p, g)), (log_n, b)
        assert np.array_equal(_run(torch_cuda, ctx.coset_inverse, got, log_n, s, batch), x), (log_n, batch)
    # polynomial product of degree < n/2 factors through transforms of length n
    log_n = 10
    n = 1 << log_n
    a = co.splitmix64_elements(n, p, 1)
    b = co.splitmix64_elements(n, p, 2)
    a[n // 2:] = 0
    b[n // 2:] = 0
    torch = torch_cuda
    da = torch.from_numpy(a.view(np.int64)).cuda()
    db = torch.from_numpy(b.view(np.int64)).cuda()
    ctx.forward(da, log_n)
    ctx.forward(db, log_n)
    ctx.vec_mul(da, db)
    ctx.inverse(da, log_n)
    torch.cuda.synchronize()
    got = da.cpu().numpy().view(np.uint64)
    want = [0] * n
    ai, bi = [int(v) for v in a[: n // 2]], [int(v) for v in b[: n // 2]]
    for i, u in enumerate(ai):
        for j, v in enumerate(bi):
            want[i + j] = (want[i + j] + u * v) % p
    assert [int(v) for v in got] == want


def test_coset_with_factored_tables(torch_cuda):
    """Above the table limit the coset shift comes from two factored tables and a separate element-wise pass
    (the inverse's low table carries 1/n): same bits as the oracle on pre-scaled inputs, and the round trip."""
    from ntt_b200.api64 import Ntt64
    from oracle import c_oracle as co

    for field, p, g in (("goldilocks", GL_P, GL_G), ("ref29", REF_P, REF_G)):
        ctx = Ntt64(field, max_log_n=16)
        ctx.set_coset_table_limit(1024)  # 128 elements: everything below uses the factored tables
        s = 0xABCDEF12345 % p
        for log_n, batch in ((8, 3), (11, 1), (13, 2)):
            n = 1 << log_n
            x = co.splitmix64_elements(batch << log_n, p, 640 + log_n)
            pre = np.array([(int(v) * pow(s, j % n, p)) % p for j, v in enumerate(x)], dtype=np.uint64)
            got = _run(torch_cuda, ctx.coset_forward, x, log_n, s, batch)
            for b in range(batch):
                assert np.array_equal(got.reshape(batch, -1)[b], co.ntt64(pre.reshape(batch, -1)[b], p, g)), (field, log_n, b)
            assert np.array_equal(_run(torch_cuda, ctx.coset_inverse, got, log_n, s, batch), x), (field, log_n)
        ctx.close()


def test_host_buffers_and_pipelined_batches(torch_cuda, ctxs):
    from oracle import c_oracle as co

    ctx = ctxs("goldilocks", 20)
    for log_n, batch in ((10, 1), (16, 3), (12, 4096), (20, 1)):
        x = co.splitmix64_elements(batch << log_n, GL_P, 77 + log_n)
        got = ctx.forward_host(x, log_n, batch)
        xs = x.reshape(batch, -1)
        for b in sorted({0, batch // 3, batch - 1}):
            assert np.array_equal(got.reshape(batch, -1)[b], co.ntt64(xs[b], GL_P, GL_G, threads=4)), (log_n, b)
        back = ctx.inverse_host(got, log_n, batch, out=got)  # in place
        assert np.array_equal(back, x)
    assert ctx.launch_count > 0


def test_linearity_and_checksum_at_2_26(torch_cuda):
    """Size-independent properties at 2^26 points (512 MiB): X[0] = sum x, NTT(a + b) = NTT(a) + NTT(b) on
    samples, round trip."""
    from ntt_b200.api64 import Ntt64
    from oracle import c_oracle as co

    torch = torch_cuda
    log_n = 26
    ctx = Ntt64("goldilocks", max_log_n=log_n)
    a = co.splitmix64_elements(1 << log_n, GL_P, 1)
    b = co.splitmix64_elements(1 << log_n, GL_P, 2)
    s = np.array([(int(u) + int(v)) % GL_P for u, v in zip(a[:1024], b[:1024])], dtype=np.uint64)
    A = _run(torch, ctx.forward, a, log_n)
    # X[0] through 32-bit halves (exact in Python integers without an object array)
    lo = int((a & np.uint64(0xFFFFFFFF)).sum(dtype=np.uint64))
    hi = int((a >> np.uint64(32)).sum(dtype=np.uint64))
    assert int(A[0]) == (lo + (hi << 32)) % GL_P
    B = _run(torch, ctx.forward, b, log_n)
    ab = a.copy()
    # a + b mod p elementwise without overflow: through Python for the sampled prefix only
    ab[:1024] = s
    ab[1024:] = 0
    a0 = a.copy(); a0[1024:] = 0
    b0 = b.copy(); b0[1024:] = 0
    S = _run(torch, ctx.forward, ab, log_n)
    A0 = _run(torch, ctx.forward, a0, log_n)
    B0 = _run(torch, ctx.forward, b0, log_n)
    idx = np.array([0, 1, 2, 12345, (1 << 25) + 7, (1 << 26) - 1])
    for k in idx:
        assert int(S[k]) == (int(A0[k]) + int(B0[k])) % GL_P
    assert np.array_equal(_run(torch, ctx.inverse, A, log_n), a)
    del B
    ctx.close()


@pytest.mark.parametrize("field,p,g", [("goldilocks", GL_P, GL_G), ("ref29", REF_P, REF_G)])
@pytest.mark.parametrize("log_n,world", [(12, 1), (14, 2), (16, 8), (18, 4), (21, 8), (22, 2)])
def test_fourstep_ranks_emulated_on_one_gpu(torch_cuda, ctxs, field, p, g, log_n, world):
    """The 64-bit four-step pieces: scatter -> step 1 per rank -> the all-to-all by its host definition -> step 2 per
    rank -> gather equals the single-transform oracle; the inverse pair takes the output shards back."""
    from ntt_b200 import dist as nd
    from oracle import c_oracle as co

    torch = torch_cuda
    ctx = ctxs(field, 22)
    log_p, log_r = world.bit_length() - 1, nd.default_log_r(log_n)
    x = co.splitmix64_elements(1 << log_n, p, 40 + log_n)
    want = co.ntt64(x, p, g, threads=8)

    def each(fn, shards):
        out = []
        for r, sh in enumerate(shards):
            d = torch.from_numpy(np.ascontiguousarray(sh).reshape(-1).view(np.int64)).cuda()
            fn(d, log_n, log_r, log_p, r)
            torch.cuda.synchronize()
            out.append(d.cpu().numpy().view(np.uint64).reshape(-1, 1))
        return out

    ins = nd.scatter_input(x.reshape(-1, 1), log_n, world, log_r, limbs=1)
    s2 = each(ctx.fourstep_step2, nd.exchange_reference(each(ctx.fourstep_step1, ins), world))
    assert np.array_equal(nd.gather_output(s2, log_n, world, log_r, limbs=1).reshape(-1), want)
    back = each(ctx.fourstep_inverse_step2, nd.exchange_reference(each(ctx.fourstep_inverse_step1, s2), world))
    for r in range(world):
        assert np.array_equal(back[r], ins[r]), r


@pytest.mark.parametrize("log_n,world", [(14, 2), (18, 8)])
def test_fused_peer_store_on_one_gpu(torch_cuda, ctxs, log_n, world):
    """ntt64_fourstep_step1_peer with every rank's receive buffer on this GPU: the store phase of the butterfly
    kernel delivers exactly what step 1 + all-to-all delivers."""
    from ntt_b200 import dist as nd
    from oracle import c_oracle as co

    torch = torch_cuda
    ctx = ctxs("goldilocks", 22)
    lp, lr = world.bit_length() - 1, nd.default_log_r(log_n)
    x = co.splitmix64_elements(1 << log_n, GL_P, 99 + log_n)
    shards = nd.scatter_input(x.reshape(-1, 1), log_n, world, limbs=1)
    recv = [torch.zeros(1 << (log_n - lp), dtype=torch.int64, device="cuda") for _ in range(world)]
    for r in range(world):
        d = torch.from_numpy(np.ascontiguousarray(shards[r]).reshape(-1).view(np.int64)).cuda()
        ctx.fourstep_step1_peer(d, [t.data_ptr() for t in recv], log_n, lr, lp, r)
    torch.cuda.synchronize()
    outs = []
    for r in range(world):
        ctx.fourstep_step2(recv[r], log_n, lr, lp, r)
        torch.cuda.synchronize()
        outs.append(recv[r].cpu().numpy().view(np.uint64).reshape(-1, 1))
    assert np.array_equal(nd.gather_output(outs, log_n, world, lr, limbs=1).reshape(-1), co.ntt64(x, GL_P, GL_G, threads=8))
