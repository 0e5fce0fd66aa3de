"""384-bit containers (include/ntt384.h, SURVEY.md par. 8f row 4): the CUDA path through the C ABI
against the Python-int oracle over BLS12-377 Fq (377 bits) -- bit-exact -- plus round trips and
linearity at sizes the Python oracle does not reach, and the outputs of the reference's own GZKP
driver instantiated at 512-bit containers over the same field (tests/golden/ref512_gpu_kat.json)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env():
    import torch

    from ntt_b200.api import FIELDS384, Ntt384

    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device; there is no CPU fallback to test")
    torch.cuda.set_device(0)
    ctx = Ntt384("bls12_377_fq", max_log_n=22)
    yield torch, ctx, FIELDS384["bls12_377_fq"]
    ctx.close()


def _to_limbs(vals):
    a = np.zeros((len(vals), 12), dtype=np.uint32)
    for i, v in enumerate(vals):
        for k in range(12):
            a[i, k] = (v >> (32 * k)) & 0xFFFFFFFF
    return a


def _to_ints(a):
    return [sum(int(a[i, k]) << (32 * k) for k in range(12)) for i in range(a.shape[0])]


def _run(torch, fn, x, *args):
    d = torch.from_numpy(np.array(x, dtype=np.uint32, copy=True).view(np.int32)).cuda()
    fn(d, *args)
    torch.cuda.synchronize()
    return d.cpu().numpy().view(np.uint32).reshape(-1, 12)


def test_all_sizes_vs_python_oracle(env):
    from oracle import py_oracle as po

    torch, ctx, (p, g) = env
    F = po.BLS12_377_FQ
    rng = np.random.default_rng(384)
    for log_n in range(0, 14):
        n = 1 << log_n
        vals = [int.from_bytes(rng.bytes(48), "little") % p for _ in range(n)]
        if n >= 4:
            vals[0], vals[1], vals[2], vals[3] = 0, p - 1, 1, p - 2
        x = _to_limbs(vals)
        got = _run(torch, ctx.forward, x, log_n)
        assert _to_ints(got) == po.ntt_ref(vals, F), f"forward 2^{log_n}"
        back = _run(torch, ctx.inverse, got, log_n)
        assert np.array_equal(back, x), f"inverse 2^{log_n}"
        assert _to_ints(_run(torch, ctx.inverse, x, log_n)) == po.ntt_ref(vals, F, inverse=True)


def test_batches(env):
    from oracle import py_oracle as po

    torch, ctx, (p, g) = env
    F = po.BLS12_377_FQ
    rng = np.random.default_rng(7)
    for log_n, batch in ((1, 5), (2, 7), (6, 3), (10, 4)):
        n = 1 << log_n
        vals = [int.from_bytes(rng.bytes(48), "little") % p for _ in range(n * batch)]
        got = _to_ints(_run(torch, ctx.forward, _to_limbs(vals), log_n, batch))
        for b in range(batch):
            assert got[b * n:(b + 1) * n] == po.ntt_ref(vals[b * n:(b + 1) * n], F), (log_n, b)


def test_vec_mul(env):
    torch, ctx, (p, g) = env
    rng = np.random.default_rng(9)
    n = 1000
    a = [int.from_bytes(rng.bytes(48), "little") % p for _ in range(n)]
    b = [int.from_bytes(rng.bytes(48), "little") % p for _ in range(n)]
    a[:3], b[:3] = [0, p - 1, p - 1], [5, p - 1, 1]
    da = torch.from_numpy(_to_limbs(a).view(np.int32)).cuda()
    db = torch.from_numpy(_to_limbs(b).view(np.int32)).cuda()
    out = torch.empty_like(da)
    ctx.vec_mul(da, db, out=out)
    torch.cuda.synchronize()
    assert _to_ints(out.cpu().numpy().view(np.uint32).reshape(-1, 12)) == [x * y % p for x, y in zip(a, b)]
    ctx.vec_mul(da, db, out=out, mont_form=True)
    torch.cuda.synchronize()
    rinv = pow(1 << 384, -1, p)
    assert _to_ints(out.cpu().numpy().view(np.uint32).reshape(-1, 12)) == [x * y * rinv % p for x, y in zip(a, b)]


@pytest.mark.parametrize("log_n", [16, 20])
def test_large_sizes_closed_form_linearity_round_trip(env, log_n):
    """Beyond the Python oracle's reach: a geometric input x[j] = c*h^j has the closed form
    X[k] = c*(h^n - 1)/(h*w^k - 1); linearity and inverse(forward(x)) == x on random data."""
    torch, ctx, (p, g) = env
    n = 1 << log_n
    c, h = 0xC0FFEE1234567890ABCDEF % p, 3
    # geometric sequence built with the library's own element-wise product: x = c * h^j by doubling
    seq = np.zeros((n, 12), dtype=np.uint32)
    vals = [c]
    for _ in range(1, 1 << 10):
        vals.append(vals[-1] * h % p)
    block = _to_limbs(vals)
    seq[: 1 << 10] = block
    d = torch.from_numpy(seq.view(np.int32)).cuda().view(n, 12)
    filled = 1 << 10
    while filled < n:  # x[filled + j] = x[j] * h^filled
        hk = torch.from_numpy(np.repeat(_to_limbs([pow(h, filled, p)]), filled, axis=0).view(np.int32)).cuda()
        ctx.vec_mul(d[:filled].contiguous(), hk, out=hk)
        d[filled:2 * filled] = hk
        filled *= 2
    x_dev = d.contiguous().view(-1)
    X = x_dev.clone()
    ctx.forward(X, log_n)
    torch.cuda.synchronize()
    w = pow(g, (p - 1) >> log_n, p)
    num = c * (pow(h, n, p) - 1) % p
    Xh = X.view(n, 12)
    for k in (0, 1, 2, n // 2, n - 1, 12345 % n):
        got = sum(int(v) << (32 * i) for i, v in enumerate(Xh[k].cpu().numpy().view(np.uint32)))
        assert got == num * pow((h * pow(w, k, p) - 1) % p, p - 2, p) % p, k
    ctx.inverse(X, log_n)
    torch.cuda.synchronize()
    assert torch.equal(X, x_dev)
    # linearity on random canonical data (top limb small enough to stay below p)
    rng = torch.Generator(device="cuda").manual_seed(1)
    a = torch.randint(0, 2**31 - 1, (n, 12), dtype=torch.int32, device="cuda", generator=rng)
    b = torch.randint(0, 2**31 - 1, (n, 12), dtype=torch.int32, device="cuda", generator=rng)
    a[:, 11] &= 0x00FFFFFF
    b[:, 11] &= 0x00FFFFFF  # a, b < 2^376 < p and a + b < p? no: keep the sum below p by halving the range
    a[:, 11] >>= 1
    b[:, 11] >>= 1
    s64 = a.to(torch.int64) + b.to(torch.int64)  # limb-wise sum with manual carry propagation
    carry = torch.zeros(n, dtype=torch.int64, device="cuda")
    s = torch.empty_like(a)
    for k in range(12):
        t = s64[:, k] + carry
        s[:, k] = (t & 0xFFFFFFFF).to(torch.int64).where(t & 0xFFFFFFFF < 2**31, (t & 0xFFFFFFFF) - 2**32).to(torch.int32)
        carry = t >> 32
    A, B, S = a.clone().view(-1), b.clone().view(-1), s.clone().view(-1)
    for t in (A, B, S):
        ctx.forward(t, log_n)
    torch.cuda.synchronize()
    # NTT(a) + NTT(b) mod p must equal NTT(a + b): compare a sample of positions with Python ints
    Ah, Bh, Sh = (t.view(n, 12) for t in (A, B, S))
    for k in (0, 1, n // 3, n - 1):
        ia, ib, is_ = (sum(int(v) << (32 * i) for i, v in enumerate(t[k].cpu().numpy().view(np.uint32))) for t in (Ah, Bh, Sh))
        assert (ia + ib) % p == is_, k


def test_library_matches_the_reference_gzkp_golden_vectors(env):
    """tests/golden/ref512_gpu_kat.json: outputs of the reference's own NTT_GZKP (512-bit containers)
    over BLS12-377 Fq on a B200; ntt384_forward must give the same 48-byte images."""
    import hashlib
    import json
    import os

    torch, ctx, (p, g) = env
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref512_gpu_kat.json")) as f:
        kat = json.load(f)
    for case in kat["cases"]:
        n = 1 << case["log_n"]
        if case["input"] == "iota":
            vals = list(range(n))
        else:
            rng = np.random.default_rng(case["seed"])
            vals = [int.from_bytes(rng.bytes(48), "little") % p for _ in range(n)]
        got = _run(torch, ctx.forward, _to_limbs(vals), case["log_n"])
        assert hashlib.sha256(got.tobytes()).hexdigest() == case["sha256"], (case["log_n"], case["input"])


def test_library_against_the_reference_gzkp_driver_live(env):
    """When oracle/_ref/libref512.so travelled to the box: fresh random inputs through the reference's
    NTT_GZKP<16,512> and through ntt384_forward, compared as integers."""
    import ctypes
    import os

    torch, ctx, (p, g) = env
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "libref512.so")
    if not os.path.exists(path):
        pytest.skip("oracle/_ref/libref512.so not present (built only where /root/reference exists)")
    R = ctypes.CDLL(path)
    u32p = ctypes.POINTER(ctypes.c_uint32)
    R.ref512_ntt_gzkp.argtypes = [u32p, ctypes.c_uint, u32p, u32p]
    rng = np.random.default_rng()
    pl = np.array([(p >> (32 * k)) & 0xFFFFFFFF for k in range(16)], dtype=np.uint32)
    gl = np.array([(g >> (32 * k)) & 0xFFFFFFFF for k in range(16)], dtype=np.uint32)
    for log_n in (5, 9, 13):
        vals = [int.from_bytes(rng.bytes(48), "little") % p for _ in range(1 << log_n)]
        a = np.zeros((1 << log_n, 16), dtype=np.uint32)
        a[:, :12] = _to_limbs(vals)
        assert R.ref512_ntt_gzkp(a.ctypes.data_as(u32p), log_n, pl.ctypes.data_as(u32p), gl.ctypes.data_as(u32p)) == 0
        ours = _run(torch, ctx.forward, _to_limbs(vals), log_n)
        assert not a[:, 12:].any() and np.array_equal(a[:, :12], ours), log_n


def test_full_c_oracle_2p18(env):
    """Every output of a 2^18-point transform against the C oracle (oracle/ntt_oracle_n.c, pinned to the
    reference's GZKP digests), forward and inverse, plus a batch of 8 x 2^15."""
    from oracle import c_oracle as co

    torch, ctx, (p, g) = env
    rng = np.random.default_rng(2018)
    log_n = 18
    x = rng.integers(0, 1 << 32, size=(1 << log_n, 12), dtype=np.uint64).astype(np.uint32)
    x[:, 11] &= 0x00FFFFFF  # below 2^376 < p
    want = co.ntt384(x, p, g, threads=8)
    got = _run(torch, ctx.forward, x, log_n)
    assert np.array_equal(got, want)
    assert np.array_equal(_run(torch, ctx.inverse, got, log_n), x)
    n = 1 << 15
    gotb = _run(torch, ctx.forward, x, 15, 8)
    for b in range(8):
        assert np.array_equal(gotb[b * n:(b + 1) * n], co.ntt384(x[b * n:(b + 1) * n], p, g, threads=8)), b


def _rand384(seed, n):
    rng = np.random.default_rng(seed)
    x = rng.integers(0, 1 << 32, size=(n, 12), dtype=np.uint64).astype(np.uint32)
    x[:, 11] &= 0x00FFFFFF  # below 2^376 < p
    return x


@pytest.mark.parametrize("log_n,world", [(12, 1), (14, 2), (16, 8), (18, 4)])
def test_fourstep_ranks_emulated_on_one_gpu(env, log_n, world):
    """The 384-bit four-step pieces (same pass body as the 256-bit ones): scatter -> step 1 per rank -> the
    all-to-all by its host definition -> step 2 per rank -> gather equals the single-transform C oracle; the
    inverse pair takes the output shards back to the input shards."""
    from ntt_b200 import dist as nd
    from oracle import c_oracle as co

    torch, ctx, (p, g) = env
    log_p, log_r = world.bit_length() - 1, nd.default_log_r(log_n)
    x = _rand384(40 + log_n, 1 << log_n)
    want = co.ntt384(x, p, g, threads=8)

    def each(fn, shards):
        out = []
        for r, sh in enumerate(shards):
            d = torch.from_numpy(np.ascontiguousarray(sh).view(np.int32)).cuda().view(-1)
            fn(d, log_n, log_r, log_p, r)
            torch.cuda.synchronize()
            out.append(d.cpu().numpy().view(np.uint32).reshape(-1, 12))
        return out

    ins = nd.scatter_input(x, log_n, world, log_r, limbs=12)
    s2 = each(ctx.fourstep_step2, nd.exchange_reference(each(ctx.fourstep_step1, ins), world))
    assert np.array_equal(nd.gather_output(s2, log_n, world, log_r, limbs=12), want)
    back = each(ctx.fourstep_inverse_step2, nd.exchange_reference(each(ctx.fourstep_inverse_step1, s2), world))
    for r in range(world):
        assert np.array_equal(back[r], ins[r]), r


@pytest.mark.parametrize("log_n,world", [(14, 2), (17, 8)])
def test_fused_peer_store_on_one_gpu(env, log_n, world):
    """ntt384_fourstep_step1_peer with every rank's receive buffer on this GPU: the store phase of the butterfly
    kernel delivers exactly what step 1 + all-to-all delivers."""
    from ntt_b200 import dist as nd
    from oracle import c_oracle as co

    torch, ctx, (p, g) = env
    lp, lr = world.bit_length() - 1, nd.default_log_r(log_n)
    x = _rand384(99 + log_n, 1 << log_n)
    shards = nd.scatter_input(x, log_n, world, limbs=12)
    recv = [torch.zeros(12 << (log_n - lp), dtype=torch.int32, device="cuda") for _ in range(world)]
    for r in range(world):
        d = torch.from_numpy(shards[r].view(np.int32)).cuda().view(-1)
        ctx.fourstep_step1_peer(d, [t.data_ptr() for t in recv], log_n, lr, lp, r)
    torch.cuda.synchronize()
    outs = []
    for r in range(world):
        ctx.fourstep_step2(recv[r], log_n, lr, lp, r)
        torch.cuda.synchronize()
        outs.append(recv[r].cpu().numpy().view(np.uint32).reshape(-1, 12))
    assert np.array_equal(nd.gather_output(outs, log_n, world, lr, limbs=12), co.ntt384(x, p, g, threads=8))


def test_host_buffers_pinned_and_pageable(env):
    from oracle import c_oracle as co

    torch, ctx, (p, g) = env
    for log_n, batch in ((5, 3), (12, 2), (16, 1)):
        x = _rand384(7 + log_n, batch << log_n)
        want = np.concatenate([co.ntt384(x[b << log_n:(b + 1) << log_n], p, g, threads=4) for b in range(batch)])
        assert np.array_equal(ctx.forward_host(x, log_n, batch).reshape(-1, 12), want)  # pageable: staged copies
        h_in = torch.from_numpy(x.view(np.int32).copy()).pin_memory()
        h_out = torch.zeros_like(h_in).pin_memory()
        got = ctx.forward_host(h_in.numpy().view(np.uint32), log_n, batch, out=h_out.numpy().view(np.uint32))
        assert np.array_equal(got.reshape(-1, 12), want)  # pinned: zero-copy end passes
        back = ctx.inverse_host(h_out.numpy().view(np.uint32), log_n, batch, out=h_out.numpy().view(np.uint32))
        assert np.array_equal(back.reshape(-1, 12), x)


def test_coset_transforms_and_fused_pointwise_product(env):
    """coset_forward = NTT of x[j] * s^j; coset_inverse undoes it; forward_mul multiplies the outputs by a
    Montgomery-form vector inside the last pass."""
    from oracle import c_oracle as co

    torch, ctx, (p, g) = env
    s = 0x1234567 * (1 << 300) + 12345
    R = 1 << 384
    for log_n, batch in ((3, 2), (10, 1), (13, 2)):
        n = 1 << log_n
        x = _rand384(300 + log_n, batch << log_n)
        xi = _to_ints(x)
        pre = _to_limbs([v * pow(s, j % n, p) % p for j, v in enumerate(xi)])
        want = np.concatenate([co.ntt384(pre[b << log_n:(b + 1) << log_n], p, g, threads=4) for b in range(batch)])
        got = _run(torch, ctx.coset_forward, x, log_n, s, batch)
        assert np.array_equal(got, want), log_n
        assert np.array_equal(_run(torch, ctx.coset_inverse, got, log_n, s, batch), x), log_n
        # fused pointwise product: outputs times m (given in Montgomery form)
        m = _to_ints(_rand384(900 + log_n, batch << log_n))
        d = torch.from_numpy(x.view(np.int32).copy()).cuda()
        dm = torch.from_numpy(_to_limbs([v * R % p for v in m]).view(np.int32)).cuda()
        ctx.forward_mul(d, dm, log_n, batch)
        torch.cuda.synchronize()
        plain = np.concatenate([co.ntt384(x[b << log_n:(b + 1) << log_n], p, g, threads=4) for b in range(batch)])
        want_mul = _to_limbs([a * b % p for a, b in zip(_to_ints(plain), m)])
        assert np.array_equal(d.cpu().numpy().view(np.uint32).reshape(-1, 12), want_mul), log_n
