"""Parity tests proper: the CUDA path, called through the C ABI (ctypes -> libntt256.so),
against the CPU oracle on the same seeded inputs -- bit-exact -- plus golden vectors and, at
BASELINE.json's full sizes, size-independent properties (round trip, sampled outputs by
Horner, linearity).  Run on the B200 box:  python -m pytest tests -m gpu -x -q"""
import hashlib
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def torch_cuda():
    import torch

    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device; there is no CPU fallback to test")
    torch.cuda.set_device(0)
    return torch


@pytest.fixture(scope="module")
def ctxs(torch_cuda):
    from ntt_b200.api import Ntt256

    made = {}

    def get(field, max_log_n):
        key = (field, max_log_n)
        if key not in made:
            made[key] = Ntt256(field, max_log_n=max_log_n)
        return made[key]

    yield get
    for c in made.values():
        c.close()


def _gpu_forward(torch, ctx, x, log_n, batch=1, inverse=False):
    d = torch.from_numpy(np.array(x, dtype=np.uint32, copy=True).view(np.int32)).cuda()
    (ctx.inverse if inverse else ctx.forward)(d, log_n, batch)
    torch.cuda.synchronize()
    return d.cpu().numpy().view(np.uint32).reshape(-1, 8)


# --------------------------------------------------------------------------- config 1
@pytest.mark.parametrize("field", ["bn254_fr", "bls12_381_fr"])
def test_all_sizes_vs_oracle(torch_cuda, ctxs, field):
    """2^0 .. 2^16 points (config 1 of BASELINE.json is the 2^16 BN254 case), forward and inverse."""
    from ntt_b200.api import FIELDS
    from oracle import c_oracle as co

    p, g = FIELDS[field]
    ctx = ctxs(field, 20)
    for log_n in range(0, 17):
        x = co.splitmix_elements(1 << log_n, p, 0x4E5454 + log_n)
        want = co.ntt(x, p, g, threads=4)
        got = _gpu_forward(torch_cuda, ctx, x, log_n)
        assert np.array_equal(got, want), f"forward 2^{log_n}"
        back = _gpu_forward(torch_cuda, ctx, got, log_n, inverse=True)
        assert np.array_equal(back, x), f"inverse 2^{log_n}"
        wanti = co.ntt(x, p, g, inverse=True, threads=4)
        assert np.array_equal(_gpu_forward(torch_cuda, ctx, x, log_n, inverse=True), wanti)


@pytest.mark.parametrize("tile_bits,coalesce_bits", [(6, 1), (8, 3), (10, 2), (11, 2), (12, 2), (12, 0)])
def test_every_tiling_gives_the_same_bits(torch_cuda, tile_bits, coalesce_bits):
    """Plan selection must not change results: sweep tile sizes / coalescing widths."""
    from ntt_b200.api import FIELDS, Ntt256
    from oracle import c_oracle as co

    p, g = FIELDS["bls12_381_fr"]
    ctx = Ntt256("bls12_381_fr", max_log_n=18)
    ctx.set_tuning(tile_bits, coalesce_bits)
    for log_n in (7, 12, 13, 15, 18):
        x = co.splitmix_elements(1 << log_n, p, 31 + log_n)
        assert np.array_equal(_gpu_forward(torch_cuda, ctx, x, log_n), co.ntt(x, p, g, threads=4))
    ctx.close()


# --------------------------------------------------------------------------- reference KATs
def test_reference_field_kats(torch_cuda):
    """Created with the reference's own field (p = 469762049, g = 3, big-num.cu:9-10) and its own
    input x[j] = j (big-num.cu:399-402), the library must reproduce the outputs of the
    reference's NTT(): tests/golden/ref_i64_kat.json was generated from the reference's code."""
    from ntt_b200.api import Ntt256

    with open(os.path.join(GOLD, "ref_i64_kat.json")) as f:
        gold = json.load(f)
    ctx = Ntt256(p=gold["p"], g=gold["g"], max_log_n=20)
    for case in gold["cases"]:
        log_n = case["log_n"]
        n = 1 << log_n
        x = np.zeros((n, 8), dtype=np.uint32)
        x[:, 0] = np.arange(n, dtype=np.uint32)  # int64_to_cgbn(i), big-num.cu:355-362
        got = _gpu_forward(torch_cuda, ctx, x, log_n)
        assert not got[:, 2:].any()
        lo64 = got[:, 0].astype(np.int64) | (got[:, 1].astype(np.int64) << 32)  # cgbn_to_int64, :364-367
        assert lo64[:8].tolist() == case["head"][: min(8, n)]
        assert hashlib.sha256(lo64.tobytes()).hexdigest() == case["sha256_i64le"]
    ctx.close()


def test_bigfield_golden(torch_cuda, ctxs):
    from oracle import py_oracle as po

    with open(os.path.join(GOLD, "bigfield_kat.json")) as f:
        gold = json.load(f)
    for case in gold["cases"]:
        F = po.FIELDS[case["field"]]
        n = 1 << case["log_n"]
        vals = list(range(n)) if case["input"] == "iota" else po.splitmix_elements(n, F, case["seed"])
        x = np.frombuffer(po.to_bytes(vals), dtype=np.uint32).reshape(n, 8)
        got = _gpu_forward(torch_cuda, ctxs(case["field"], 20), x, case["log_n"])
        assert hashlib.sha256(got.tobytes()).hexdigest() == case["sha256_image"]


def test_reference_named_entry_points(torch_cuda):
    """NTT_GZKP / NTT_GPU_Naive keep the reference's argument meaning (device data in place,
    host prime, omega = generator) and both equal the CPU NTT (big-num.cu:439-467)."""
    from ntt_b200.api import FIELDS, NTT_GPU_Naive, NTT_GZKP
    from oracle import c_oracle as co

    torch = torch_cuda
    p, g = FIELDS["bn254_fr"]
    for kk in (5, 9, 14):  # the reference sweeps kk = 5..24 (big-num.cu:370)
        length = 1 << kk
        x = co.splitmix_elements(length, p, kk)
        want = co.ntt(x, p, g, threads=4)
        d1 = torch.from_numpy(x.view(np.int32)).cuda()
        NTT_GPU_Naive(d1, length, None, 0, prime=p, omega=g)
        d2 = torch.from_numpy(x.view(np.int32)).cuda()
        NTT_GZKP(d2, length, None, 0, prime=co.limbs(p), omega=co.limbs(g), B=5, G=8)
        torch.cuda.synchronize()
        assert np.array_equal(d1.cpu().numpy().view(np.uint32), want)
        assert np.array_equal(d2.cpu().numpy().view(np.uint32), want)
    with pytest.raises(ValueError):
        NTT_GZKP(d1, 48, prime=p, omega=g)


# --------------------------------------------------------------------------- config 2
def test_config2_bls_2p20_forward_inverse(torch_cuda, ctxs):
    from ntt_b200.api import FIELDS
    from oracle import c_oracle as co

    p, g = FIELDS["bls12_381_fr"]
    ctx = ctxs("bls12_381_fr", 20)
    x = co.splitmix_elements(1 << 20, p, 0x4E5454 + 2)
    fwd = _gpu_forward(torch_cuda, ctx, x, 20)
    assert np.array_equal(fwd, co.ntt(x, p, g, threads=8))
    assert np.array_equal(_gpu_forward(torch_cuda, ctx, fwd, 20, inverse=True), x)


# --------------------------------------------------------------------------- config 3 (one GPU's share)
def test_batches(torch_cuda, ctxs):
    from ntt_b200.api import FIELDS
    from oracle import c_oracle as co

    p, g = FIELDS["bn254_fr"]
    ctx = ctxs("bn254_fr", 20)
    for log_n, batch in ((1, 5), (2, 7), (4, 3), (10, 9), (12, 32), (16, 4)):
        n = 1 << log_n
        x = co.splitmix_elements(n * batch, p, 500 + log_n)
        got = _gpu_forward(torch_cuda, ctx, x, log_n, batch=batch)
        for b in range(batch):
            assert np.array_equal(got[b * n:(b + 1) * n], co.ntt(x[b * n:(b + 1) * n], p, g)), (log_n, b)
        back = _gpu_forward(torch_cuda, ctx, got, log_n, batch=batch, inverse=True)
        assert np.array_equal(back, x)


def test_config3_batch256_of_2p16_properties(torch_cuda, ctxs):
    """256 x 2^16 (512 MiB): every transform must equal the single-transform result of the same
    data (checked bit-exactly on 3 of them) and the batch must round-trip."""
    from ntt_b200.api import FIELDS
    from oracle import c_oracle as co

    torch = torch_cuda
    p, g = FIELDS["bn254_fr"]
    ctx = ctxs("bn254_fr", 20)
    n, batch = 1 << 16, 256
    x = co.splitmix_elements(n * batch, p, 0x4E5454 + 3)
    d = torch.from_numpy(x.view(np.int32)).cuda().view(-1)
    ctx.forward(d, 16, batch)
    torch.cuda.synchronize()
    for b in (0, 101, 255):
        got = d[b * n * 8:(b + 1) * n * 8].cpu().numpy().view(np.uint32).reshape(n, 8)
        assert np.array_equal(got, co.ntt(x[b * n:(b + 1) * n], p, g, threads=8))
    ctx.inverse(d, 16, batch)
    torch.cuda.synchronize()
    assert np.array_equal(d.cpu().numpy().view(np.uint32).reshape(-1, 8), x)


# --------------------------------------------------------------------------- configs 4/5 on one GPU
@pytest.mark.parametrize("field,log_n", [("bn254_fr", 24), ("bls12_381_fr", 22)])
def test_large_transform_properties(torch_cuda, ctxs, field, log_n):
    """Full-size checks that need no full second oracle run: sampled outputs by Horner on the
    CPU, linearity NTT(x+y) = NTT(x)+NTT(y), and inverse(forward(x)) == x."""
    from ntt_b200.api import FIELDS
    from oracle import c_oracle as co

    torch = torch_cuda
    p, g = FIELDS[field]
    ctx = ctxs(field, 24)
    n = 1 << log_n
    x = co.splitmix_elements(n, p, 0x4E5454 + 4)
    y = co.splitmix_elements(n, p, 0x4E5454 + 5)
    X = _gpu_forward(torch, ctx, x, log_n)
    for k in (0, 1, n // 2 + 7, n - 1):
        assert co.dft_sample(x, p, g, k) == co.array_to_ints(X[k:k + 1])[0], k
    Y = _gpu_forward(torch, ctx, y, log_n)
    S = _gpu_forward(torch, ctx, co.vec_add(x, y, p), log_n)
    assert np.array_equal(S, co.vec_add(X, Y, p))
    assert np.array_equal(_gpu_forward(torch, ctx, X, log_n, inverse=True), x)
    assert (co.array_to_ints(X[:64]) < np.array([p] * 64, dtype=object)).all()  # canonical outputs


def test_full_oracle_2p22(torch_cuda, ctxs):
    """Full bit-exact comparison at 2^22 (a three-pass plan; 2^24 below is the four-pass one)."""
    from ntt_b200.api import FIELDS
    from oracle import c_oracle as co

    p, g = FIELDS["bn254_fr"]
    x = co.splitmix_elements(1 << 22, p, 99)
    assert np.array_equal(_gpu_forward(torch_cuda, ctxs("bn254_fr", 24), x, 22), co.ntt(x, p, g, threads=16))


@pytest.mark.parametrize("field", ["bn254_fr", "bls12_381_fr"])
def test_config4_full_2p24_bit_exact(torch_cuda, ctxs, field):
    """The headline size, every element: the reference's own harness compares all 2^24 outputs with its
    CPU NTT (/root/reference/src/big-num.cu:446-467); so does this test, forward and inverse, against
    the multi-threaded C oracle (a few seconds on the box's host cores)."""
    from ntt_b200.api import FIELDS
    from oracle import c_oracle as co

    p, g = FIELDS[field]
    n = 1 << 24
    x = co.splitmix_elements(n, p, 0x4E5454 + 24)
    threads = max(1, min(co.max_threads(), os.cpu_count() or 1))
    want = co.ntt(x, p, g, threads=threads)
    got = _gpu_forward(torch_cuda, ctxs(field, 24), x, 24)
    assert np.array_equal(got, want)
    assert np.array_equal(_gpu_forward(torch_cuda, ctxs(field, 24), want, 24, inverse=True), x)


def test_shoup_and_montgomery_twiddles_give_the_same_bits(torch_cuda):
    """BN254 Fr (p < 2^254) runs on (w, floor(w 2^256/p)) tables and Shoup products by default; the
    Montgomery path (the only one for BLS12-381 Fr) must produce identical bytes: every plan shape, forward
    and inverse (1/n folded into a scaled round table), batches, streamed and two-lookup pass twiddles, the
    coset entry points.  A 255-bit modulus refuses the Shoup mode."""
    from ntt_b200.api import FIELDS, Ntt256
    from ntt_b200.lib import ERR_FIELD, Ntt256Error
    from oracle import c_oracle as co

    torch = torch_cuda
    p, g = FIELDS["bn254_fr"]
    ctx = Ntt256("bn254_fr", max_log_n=22)
    assert ctx.twiddle_mode == "shoup"
    cases = [(3, 5), (7, 3), (12, 2), (16, 1), (19, 1), (21, 1)]
    data = {c: co.splitmix_elements(c[1] << c[0], p, 600 + c[0]) for c in cases}
    results = {}
    for mode in (True, False, True):
        ctx.set_twiddle_mode(mode)
        assert ctx.twiddle_mode == ("shoup" if mode else "montgomery")
        for limit in (1 << 40, 0):
            ctx.set_stream_twiddle_limit(limit)
            for (log_n, batch), x in data.items():
                X = _gpu_forward(torch, ctx, x, log_n, batch=batch)
                key = (log_n, batch)
                if key not in results:
                    n = 1 << log_n
                    for b in range(batch):
                        assert np.array_equal(X[b * n:(b + 1) * n], co.ntt(x[b * n:(b + 1) * n], p, g, threads=8)), key
                    results[key] = X
                assert np.array_equal(X, results[key]), (mode, limit, key)
                assert np.array_equal(_gpu_forward(torch, ctx, X, log_n, batch=batch, inverse=True), x), (mode, limit, key)
        d = torch.from_numpy(data[(16, 1)].view(np.int32).copy()).cuda()
        ctx.coset_forward(d, 16, 7)
        torch.cuda.synchronize()
        got = d.cpu().numpy().copy()
        assert np.array_equal(results.setdefault("coset", got), got), mode
        ctx.coset_inverse(d, 16, 7)
        torch.cuda.synchronize()
        assert np.array_equal(d.cpu().numpy().view(np.uint32).reshape(-1, 8), data[(16, 1)])
    ctx.close()
    bls = Ntt256("bls12_381_fr", max_log_n=12)
    assert bls.twiddle_mode == "montgomery"
    with pytest.raises(Ntt256Error) as e:
        bls.set_twiddle_mode(True)
    assert e.value.status == ERR_FIELD
    bls.close()


# --------------------------------------------------------------------------- edge cases
def test_edge_values(torch_cuda, ctxs):
    """0, 1, p-1 and constant inputs: maximal carries through the lazy-reduction paths."""
    from ntt_b200.api import FIELDS
    from oracle import c_oracle as co

    for field in ("bn254_fr", "bls12_381_fr"):
        p, g = FIELDS[field]
        ctx = ctxs(field, 20)
        for log_n in (3, 10, 13):
            n = 1 << log_n
            for vals in ([p - 1] * n, [0] * n, [1] + [0] * (n - 1), [(p - 1) if i % 2 else 0 for i in range(n)],
                         [p - 1 - i for i in range(n)]):
                x = co.ints_to_array(vals)
                assert np.array_equal(_gpu_forward(torch_cuda, ctx, x, log_n), co.ntt(x, p, g, threads=4))


def test_host_entry_point_and_errors(torch_cuda, ctxs):
    from ntt_b200.api import FIELDS
    from ntt_b200.lib import ERR_ARG, Ntt256Error
    from oracle import c_oracle as co

    torch = torch_cuda
    p, g = FIELDS["bn254_fr"]
    ctx = ctxs("bn254_fr", 20)
    x = co.splitmix_elements(3 << 12, p, 12)
    got = ctx.forward_host(x, 12, batch=3)
    for b in range(3):
        assert np.array_equal(got[b << 12:(b + 1) << 12], co.ntt(x[b << 12:(b + 1) << 12], p, g))
    assert np.array_equal(ctx.inverse_host(got, 12, batch=3), x)
    d = torch.zeros(8 << 21, dtype=torch.int32, device="cuda")
    with pytest.raises(Ntt256Error) as e:
        ctx.forward(d, 21)  # beyond max_log_n of this context
    assert e.value.status == ERR_ARG
    with pytest.raises(ValueError):
        ctx.forward(d, 20)  # buffer size does not match
    with pytest.raises(TypeError):
        ctx.forward(x, 12)  # host array passed to a device entry point
    before = ctx.launch_count
    ctx.forward(torch.zeros(8 << 10, dtype=torch.int32, device="cuda"), 10)
    assert ctx.launch_count > before


def test_streams_are_respected(torch_cuda, ctxs):
    from ntt_b200.api import FIELDS
    from oracle import c_oracle as co

    torch = torch_cuda
    p, g = FIELDS["bn254_fr"]
    ctx = ctxs("bn254_fr", 20)
    x = co.splitmix_elements(1 << 14, p, 3)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        d = torch.from_numpy(x.view(np.int32)).cuda(non_blocking=False)
        ctx.forward(d, 14)
        ctx.inverse(d, 14)
        ctx.forward(d, 14)
    s.synchronize()
    assert np.array_equal(d.cpu().numpy().view(np.uint32).reshape(-1, 8), co.ntt(x, p, g, threads=4))


# --------------------------------------------------------------------------- configs 4/5: four-step pieces
@pytest.mark.parametrize("log_n,world", [(12, 1), (16, 2), (16, 8), (20, 4), (21, 8)])
def test_fourstep_ranks_emulated_on_one_gpu(torch_cuda, ctxs, log_n, world):
    """Every rank's step 1 and step 2 run on this GPU one after the other, the all-to-all is
    replaced by its host definition (ntt_b200.dist.exchange_reference): scatter -> step1 ->
    exchange -> step2 -> gather must equal the single-transform oracle bit for bit."""
    from ntt_b200 import dist as nd
    from ntt_b200.api import FIELDS
    from oracle import c_oracle as co

    torch = torch_cuda
    p, g = FIELDS["bn254_fr"]
    ctx = ctxs("bn254_fr", 24)
    log_p = world.bit_length() - 1
    log_r = nd.default_log_r(log_n)
    x = co.splitmix_elements(1 << log_n, p, 0x4E5454 + 40 + log_n)
    want = co.ntt(x, p, g, threads=8)
    s1 = []
    for r, sh in enumerate(nd.scatter_input(x, log_n, world, log_r)):
        d = torch.from_numpy(sh.view(np.int32)).cuda().view(-1)
        ctx.fourstep_step1(d, log_n, log_r, log_p, r)
        torch.cuda.synchronize()
        s1.append(d.cpu().numpy().view(np.uint32).reshape(-1, 8))
    s2 = []
    for r, sh in enumerate(nd.exchange_reference(s1, world)):
        d = torch.from_numpy(np.ascontiguousarray(sh).view(np.int32)).cuda().view(-1)
        ctx.fourstep_step2(d, log_n, log_r, log_p, r)
        torch.cuda.synchronize()
        s2.append(d.cpu().numpy().view(np.uint32).reshape(-1, 8))
    assert np.array_equal(nd.gather_output(s2, log_n, world, log_r), want)


@pytest.mark.parametrize("log_n,world", [(14, 2), (16, 4), (20, 8)])
def test_fused_peer_store_on_one_gpu(torch_cuda, ctxs, log_n, world):
    """ntt256_fourstep_step1_peer with every rank's receive buffer on this GPU: the store phase of
    the butterfly kernel must deliver exactly what step 1 + all-to-all delivers."""
    from ntt_b200 import dist as nd
    from ntt_b200.api import FIELDS
    from oracle import c_oracle as co

    torch = torch_cuda
    p, g = FIELDS["bls12_381_fr"]
    ctx = ctxs("bls12_381_fr", 24)
    lp, lr = world.bit_length() - 1, nd.default_log_r(log_n)
    x = co.splitmix_elements(1 << log_n, p, 99 + log_n)
    shards = nd.scatter_input(x, log_n, world)
    plain = []
    for r in range(world):
        d = torch.from_numpy(shards[r].view(np.int32)).cuda().view(-1)
        ctx.fourstep_step1(d, log_n, lr, lp, r)
        torch.cuda.synchronize()
        plain.append(d.cpu().numpy().view(np.uint32).reshape(-1, 8))
    want = nd.exchange_reference(plain, world)
    recv = [torch.zeros(8 << (log_n - lp), dtype=torch.int32, device="cuda") for _ in range(world)]
    for r in range(world):
        d = torch.from_numpy(shards[r].view(np.int32)).cuda().view(-1)
        ctx.fourstep_step1_peer(d, [t.data_ptr() for t in recv], log_n, lr, lp, r)
    torch.cuda.synchronize()
    outs = []
    for r in range(world):
        assert np.array_equal(recv[r].cpu().numpy().view(np.uint32).reshape(-1, 8), want[r])
        ctx.fourstep_step2(recv[r], log_n, lr, lp, r)
        torch.cuda.synchronize()
        outs.append(recv[r].cpu().numpy().view(np.uint32).reshape(-1, 8))
    assert np.array_equal(nd.gather_output(outs, log_n, world, lr), co.ntt(x, p, g, threads=8))


def test_small_contexts_and_huge_batches(torch_cuda):
    """Contexts whose tables are smaller than the defaults (max_log_n < 12: short round-twiddle
    table) and batches beyond one grid dimension (65535) must work."""
    from ntt_b200.api import FIELDS, Ntt256
    from oracle import c_oracle as co

    torch = torch_cuda
    p, g = FIELDS["bn254_fr"]
    for max_log_n in (1, 2, 3, 5, 8, 11):
        ctx = Ntt256("bn254_fr", max_log_n=max_log_n)
        for log_n in range(0, max_log_n + 1):
            x = co.splitmix_elements(3 << log_n, p, 17 * max_log_n + log_n)
            got = _gpu_forward(torch, ctx, x, log_n, batch=3)
            for b in range(3):
                assert np.array_equal(got[b << log_n:(b + 1) << log_n], co.ntt(x[b << log_n:(b + 1) << log_n], p, g)), (max_log_n, log_n)
            assert np.array_equal(_gpu_forward(torch, ctx, got, log_n, batch=3, inverse=True), x)
        ctx.close()
    ctx = Ntt256("bn254_fr", max_log_n=8)
    batch, log_n = 70000, 3
    x = co.splitmix_elements(batch << log_n, p, 5)
    got = _gpu_forward(torch, ctx, x, log_n, batch=batch)
    for b in (0, 1, 65534, 65535, 65536, 69999):
        assert np.array_equal(got[b << log_n:(b + 1) << log_n], co.ntt(x[b << log_n:(b + 1) << log_n], p, g)), b
    assert np.array_equal(_gpu_forward(torch, ctx, got, log_n, batch=batch, inverse=True), x)
    ctx.close()


def test_montgomery_form_data_is_transparent(torch_cuda, ctxs):
    """The transform is linear, so data kept in Montgomery form (x*R mod p, what most prover
    stacks store) goes through the same entry points unchanged: NTT(xR) = R*NTT(x)."""
    from ntt_b200.api import FIELDS
    from oracle import c_oracle as co

    p, g = FIELDS["bls12_381_fr"]
    R = (1 << 256) % p
    ctx = ctxs("bls12_381_fr", 20)
    log_n = 13
    vals = co.array_to_ints(co.splitmix_elements(1 << log_n, p, 77))
    xm = co.ints_to_array([v * R % p for v in vals])
    got = co.array_to_ints(_gpu_forward(torch_cuda, ctx, xm, log_n))
    want = co.array_to_ints(co.ntt(co.ints_to_array(vals), p, g, threads=4))
    assert got == [v * R % p for v in want]


# --------------------------------------------------------------------------- SURVEY par. 8f rows
def test_vec_mul_and_montgomery_conversions(torch_cuda, ctxs):
    from ntt_b200.api import FIELDS
    from oracle import c_oracle as co

    torch = torch_cuda
    for field in ("bn254_fr", "bls12_381_fr"):
        p, g = FIELDS[field]
        R = (1 << 256) % p
        Rinv = pow(R, -1, p)
        ctx = ctxs(field, 20)
        n = 5000  # not a multiple of the block size
        a = co.array_to_ints(co.splitmix_elements(n, p, 1)) 
        b = co.array_to_ints(co.splitmix_elements(n, p, 2))
        a[0], b[0], a[1], b[1], a[2], b[2] = p - 1, p - 1, 0, 5, 1, p - 1
        da = torch.from_numpy(co.ints_to_array(a).view(np.int32)).cuda().view(-1)
        db = torch.from_numpy(co.ints_to_array(b).view(np.int32)).cuda().view(-1)
        out = torch.empty_like(da)
        ctx.vec_mul(da, db, out=out)
        torch.cuda.synchronize()
        assert co.array_to_ints(out.cpu().numpy().view(np.uint32).reshape(-1, 8)) == [x * y % p for x, y in zip(a, b)]
        ctx.vec_mul(da, db, out=out, mont_form=True)
        torch.cuda.synchronize()
        assert co.array_to_ints(out.cpu().numpy().view(np.uint32).reshape(-1, 8)) == [x * y * Rinv % p for x, y in zip(a, b)]
        m = da.clone()
        ctx.to_montgomery(m)
        torch.cuda.synchronize()
        assert co.array_to_ints(m.cpu().numpy().view(np.uint32).reshape(-1, 8)) == [x * R % p for x in a]
        ctx.from_montgomery(m)
        torch.cuda.synchronize()
        assert torch.equal(m, da)
        # in-place product (out aliases a)
        ctx.vec_mul(m, db)
        torch.cuda.synchronize()
        assert co.array_to_ints(m.cpu().numpy().view(np.uint32).reshape(-1, 8)) == [x * y % p for x, y in zip(a, b)]


@pytest.mark.parametrize("log_n,batch", [(0, 1), (2, 3), (9, 1), (13, 2), (16, 1)])
def test_coset_transforms(torch_cuda, ctxs, log_n, batch):
    """X[k] = sum_j shift^j x[j] w^(jk): oracle = the reference NTT of the pre-scaled input."""
    from ntt_b200.api import FIELDS
    from oracle import c_oracle as co

    torch = torch_cuda
    p, g = FIELDS["bn254_fr"]
    ctx = ctxs("bn254_fr", 20)
    n = 1 << log_n
    for shift in (g, 0x1234567890ABCDEF1234567890ABCDEF % p):
        x = co.splitmix_elements(n * batch, p, 1000 + log_n)
        xi = co.array_to_ints(x)
        pw = [1] * n
        for j in range(1, n):
            pw[j] = pw[j - 1] * shift % p
        scaled = co.ints_to_array([xi[b * n + j] * pw[j] % p for b in range(batch) for j in range(n)])
        want = np.concatenate([co.ntt(scaled[b * n:(b + 1) * n], p, g, threads=4) for b in range(batch)])
        d = torch.from_numpy(x.view(np.int32)).cuda().view(-1)
        ctx.coset_forward(d, log_n, shift, batch=batch)
        torch.cuda.synchronize()
        assert np.array_equal(d.cpu().numpy().view(np.uint32).reshape(-1, 8), want)
        ctx.coset_inverse(d, log_n, shift, batch=batch)
        torch.cuda.synchronize()
        assert np.array_equal(d.cpu().numpy().view(np.uint32).reshape(-1, 8), x)


@pytest.mark.parametrize("field", ["bn254_fr", "bls12_381_fr"])
def test_fused_and_separate_companions_agree_and_match_the_oracle(torch_cuda, field):
    """Coset shift inside the first pass / s^-k inside the last pass / pointwise product inside the last pass
    (ntt256_ctx_set_fused_companions) against the separate element-wise kernels and against Python integers;
    sizes that run on the 2-, 4- and 8-elements-per-thread kernels, batches, one- and multi-pass plans."""
    from ntt_b200.api import FIELDS, Ntt256
    from oracle import c_oracle as co

    torch = torch_cuda
    p, g = FIELDS[field]
    R = (1 << 256) % p
    ctx = Ntt256(field, max_log_n=22)
    try:
        for log_n, batch in ((3, 5), (8, 3), (12, 2), (16, 1), (17, 2), (18, 1), (21, 1)):
            n = 1 << log_n
            shift = pow(g, 3, p) if log_n % 2 else 0x1234567890ABCDEF1234567890ABCDEF % p
            x = co.splitmix_elements(n * batch, p, 4000 + log_n)
            b = co.splitmix_elements(n * batch, p, 5000 + log_n)
            d0 = torch.from_numpy(x.view(np.int32)).cuda().view(-1)
            bm = torch.from_numpy(b.view(np.int32)).cuda().view(-1)
            ctx.to_montgomery(bm)
            outs = {}
            ctx.coset_forward(d0.clone(), log_n, shift, batch=batch)  # builds the shift tables (counted launches)
            ctx.coset_inverse(d0.clone(), log_n, shift, batch=batch)
            for fused in (True, False):
                ctx.set_fused_companions(fused)
                before = ctx.launch_count
                d = d0.clone()
                ctx.coset_forward(d, log_n, shift, batch=batch)
                cf = d.clone()
                ctx.coset_inverse(d, log_n, shift, batch=batch)
                torch.cuda.synchronize()
                assert torch.equal(d, d0), (field, log_n, fused)
                d = d0.clone()
                ctx.forward_mul(d, bm, log_n, batch=batch)
                fm = d.clone()
                d = d0.clone()
                ctx.inverse_mul(d, bm, log_n, batch=batch)
                torch.cuda.synchronize()
                outs[fused] = (cf.cpu().numpy(), fm.cpu().numpy(), d.cpu().numpy(), ctx.launch_count - before)
            for k in range(3):
                assert np.array_equal(outs[True][k], outs[False][k]), (field, log_n, k)
            assert outs[True][3] < outs[False][3]  # the fused path really launches fewer kernels
            if log_n <= 16:
                xi, bi = co.array_to_ints(x), co.array_to_ints(b)
                pw = [1] * n
                for j in range(1, n):
                    pw[j] = pw[j - 1] * shift % p
                scaled = co.ints_to_array([xi[t * n + j] * pw[j] % p for t in range(batch) for j in range(n)])
                want_cf = np.concatenate([co.ntt(scaled[t * n:(t + 1) * n], p, g, threads=4) for t in range(batch)])
                assert np.array_equal(outs[True][0].view(np.uint32).reshape(-1, 8), want_cf)
                fw = np.concatenate([co.ntt(x[t * n:(t + 1) * n], p, g, threads=4) for t in range(batch)])
                want_fm = [v * w % p for v, w in zip(co.array_to_ints(fw), bi)]
                assert co.array_to_ints(outs[True][1].view(np.uint32).reshape(-1, 8)) == want_fm
    finally:
        ctx.close()


def test_fused_coset_with_the_montgomery_twiddle_mode(torch_cuda):
    """BN254 contexts default to Shoup twiddles; the fused companions must also work with 32-byte tables."""
    from ntt_b200.api import FIELDS, Ntt256
    from oracle import c_oracle as co

    torch = torch_cuda
    p, g = FIELDS["bn254_fr"]
    ctx = Ntt256("bn254_fr", max_log_n=18)
    try:
        x = co.splitmix_elements(1 << 14, p, 99)
        res = []
        for mode in (True, False):
            ctx.set_twiddle_mode(mode)
            d = torch.from_numpy(x.view(np.int32)).cuda().view(-1)
            ctx.coset_forward(d, 14, g)
            torch.cuda.synchronize()
            res.append(d.cpu().numpy())
            ctx.coset_inverse(d, 14, g)
            torch.cuda.synchronize()
            assert np.array_equal(d.cpu().numpy().view(np.uint32).reshape(-1, 8), x)
        assert np.array_equal(res[0], res[1])
    finally:
        ctx.close()


def test_pipelined_host_batches(torch_cuda, ctxs):
    """Host-buffer entry point with a batch large enough to be streamed in chunks (3-slot ring,
    uplink / kernels / downlink on separate streams): same bits as transform-by-transform."""
    from ntt_b200.api import FIELDS
    from oracle import c_oracle as co

    torch = torch_cuda
    p, g = FIELDS["bn254_fr"]
    ctx = ctxs("bn254_fr", 20)
    log_n, batch = 12, 1100  # 128 KiB per transform -> chunks of 256 transforms, last one partial
    x = co.splitmix_elements(batch << log_n, p, 31337)
    h_in = torch.from_numpy(x.view(np.int32)).pin_memory()
    h_out = torch.empty_like(h_in).pin_memory()
    got = ctx.forward_host(h_in.numpy().view(np.uint32), log_n, batch=batch, out=h_out.numpy().view(np.uint32)).reshape(-1, 8)
    for b in list(range(0, batch, 97)) + [255, 256, 767, 768, 1023, 1024, batch - 1]:
        assert np.array_equal(got[b << log_n:(b + 1) << log_n], co.ntt(x[b << log_n:(b + 1) << log_n], p, g)), b
    back = ctx.inverse_host(got, log_n, batch=batch).reshape(-1, 8)   # pageable memory path
    assert np.array_equal(back, x)


def test_streamed_and_two_lookup_twiddles_agree(torch_cuda):
    """The first pass of a transform longer than 2^18 takes its twiddles either from a streamed
    n-entry array (default) or from two table lookups and an extra product (limit 0): same bits,
    forward and inverse, also after switching back and forth."""
    from ntt_b200.api import FIELDS, Ntt256
    from oracle import c_oracle as co

    p, g = FIELDS["bls12_381_fr"]
    ctx = Ntt256("bls12_381_fr", max_log_n=22)
    x = co.splitmix_elements(1 << 21, p, 2121)
    want = co.ntt(x, p, g, threads=8)
    for limit in (0, 1 << 30, 0, 1 << 26):
        ctx.set_stream_twiddle_limit(limit)
        got = _gpu_forward(torch_cuda, ctx, x, 21)
        assert np.array_equal(got, want), limit
        assert np.array_equal(_gpu_forward(torch_cuda, ctx, got, 21, inverse=True), x), limit
    ctx.close()


def test_streamed_twiddle_arrays_stay_within_their_budget(torch_cuda):
    """The streamed first-pass arrays are as large as the transform (per size and direction): with a budget of
    one 2^20 array the least recently used ones are dropped and rebuilt on demand, the bits never change, and
    the context reports what it holds."""
    from ntt_b200.api import FIELDS, Ntt256
    from oracle import c_oracle as co

    p, g = FIELDS["bls12_381_fr"]
    ctx = Ntt256("bls12_381_fr", max_log_n=21)
    one = 32 << 20
    ctx.set_stream_twiddle_budget(one + (one >> 1))  # room for 2^20 alone, or for 2^19 + one more 2^19
    xs = {L: co.splitmix_elements(1 << L, p, 900 + L) for L in (19, 20)}
    want = {L: co.ntt(xs[L], p, g, threads=8) for L in (19, 20)}
    seen = set()
    for L, inverse in ((19, False), (19, True), (20, False), (19, False), (20, True), (20, False), (19, True)):
        if inverse:
            assert np.array_equal(_gpu_forward(torch_cuda, ctx, want[L], L, inverse=True), xs[L]), (L, inverse)
        else:
            assert np.array_equal(_gpu_forward(torch_cuda, ctx, xs[L], L), want[L]), (L, inverse)
        total, streamed = ctx.table_bytes()
        assert 0 < streamed <= one + (one >> 1), (L, inverse, streamed)
        assert total > streamed
        seen.add(streamed)
    assert len(seen) > 1  # arrays really came and went
    ctx.set_stream_twiddle_budget(0)
    assert ctx.table_bytes()[1] == 0
    assert np.array_equal(_gpu_forward(torch_cuda, ctx, xs[19], 19), want[19])  # two-lookup path when nothing fits
    ctx.close()


def test_dependent_launch_on_and_off_give_the_same_bits(torch_cuda):
    """The pass kernels are launched with programmatic stream serialization (a pass becomes resident while its
    predecessor drains, and waits before it touches data): same bits as plain launches for back-to-back
    transforms of every kernel shape, forward and inverse, without host synchronisation in between."""
    from ntt_b200.api import FIELDS, Ntt256
    from oracle import c_oracle as co

    torch = torch_cuda
    p, g = FIELDS["bn254_fr"]
    ctx = Ntt256("bn254_fr", max_log_n=22)
    for log_n, batch in ((10, 1), (14, 1), (16, 1), (18, 1), (20, 1), (22, 1), (12, 50)):
        x = co.splitmix_elements(batch << log_n, p, 31 + log_n)
        want = co.ntt(x[: 1 << log_n], p, g, threads=8)
        for enable in (True, False):
            ctx.set_dependent_launch(enable)
            d = torch.from_numpy(x.view(np.int32).copy()).cuda()
            for _ in range(3):  # forward / inverse pairs back to back on the stream
                ctx.forward(d, log_n, batch)
                ctx.inverse(d, log_n, batch)
            ctx.forward(d, log_n, batch)
            torch.cuda.synchronize()
            got = d.cpu().numpy().view(np.uint32).reshape(-1, 8)
            assert np.array_equal(got[: 1 << log_n], want), (log_n, batch, enable)
    ctx.close()


def test_constant_field_kernel_matches_the_generic_instance(torch_cuda):
    """BN254 Fr contexts run the throughput kernel in an instance whose field constants are compile-time immediates:
    same bits as the generic instance and the oracle (forward, inverse, batches, a four-step shard), and a context
    over another modulus never selects it."""
    from ntt_b200 import dist as nd
    from ntt_b200.api import FIELDS, Ntt256
    from oracle import c_oracle as co

    torch = torch_cuda
    p, g = FIELDS["bn254_fr"]
    ctx = Ntt256("bn254_fr", max_log_n=22)
    for log_n, batch in ((21, 1), (22, 1), (16, 20), (12, 300)):
        x = co.splitmix_elements(batch << log_n, p, 901 + log_n)
        want = co.ntt(x[: 1 << log_n], p, g, threads=8)
        outs = []
        for enable in (True, False):
            ctx.set_constant_field_kernel(enable)
            got = _gpu_forward(torch, ctx, x, log_n, batch)
            assert np.array_equal(got[: 1 << log_n], want), (log_n, batch, enable)
            assert np.array_equal(_gpu_forward(torch, ctx, got, log_n, batch, inverse=True), x), (log_n, batch, enable)
            outs.append(got)
        assert np.array_equal(outs[0], outs[1])
    log_n, lr, lp = 22, nd.default_log_r(22), 2
    sh = co.splitmix_elements(1 << (log_n - lp), p, 77)
    outs = []
    for enable in (True, False):
        ctx.set_constant_field_kernel(enable)
        d = torch.from_numpy(sh.view(np.int32).copy()).cuda().view(-1)
        ctx.fourstep_step1(d, log_n, lr, lp, 3)
        ctx.fourstep_step2(d, log_n, lr, lp, 3)
        torch.cuda.synchronize()
        outs.append(d.cpu().numpy())
    assert np.array_equal(outs[0], outs[1])
    ctx.close()


def test_zero_copy_host_buffers(torch_cuda, ctxs):
    """Pinned host buffers: the first pass reads host memory and the last pass writes it directly
    (no staging copies).  Distinct and aliased in/out buffers, one pass and several passes."""
    from ntt_b200.api import FIELDS
    from oracle import c_oracle as co

    torch = torch_cuda
    p, g = FIELDS["bn254_fr"]
    ctx = ctxs("bn254_fr", 24)
    for log_n, batch in ((5, 1), (10, 3), (14, 1), (20, 1), (21, 2)):
        x = co.splitmix_elements(batch << log_n, p, 4000 + log_n)
        h_in = torch.from_numpy(x.view(np.int32).copy()).pin_memory()
        h_out = torch.zeros_like(h_in).pin_memory()
        got = ctx.forward_host(h_in.numpy().view(np.uint32), log_n, batch=batch, out=h_out.numpy().view(np.uint32)).reshape(-1, 8)
        assert np.array_equal(h_in.numpy().view(np.uint32).reshape(-1, 8), x), "input buffer must be untouched"
        n = 1 << log_n
        for b in range(batch):
            assert np.array_equal(got[b * n:(b + 1) * n], co.ntt(x[b * n:(b + 1) * n], p, g, threads=8)), (log_n, b)
        # aliased: transform the pinned buffer in place, then invert it in place
        buf = h_in.numpy().view(np.uint32)
        ctx.forward_host(buf, log_n, batch=batch, out=buf)
        assert np.array_equal(buf.reshape(-1, 8), got)
        ctx.inverse_host(buf, log_n, batch=batch, out=buf)
        assert np.array_equal(buf.reshape(-1, 8), x)


def test_zero_copy_host_buffers_with_more_transforms_than_one_grid(torch_cuda, ctxs):
    """Batches above 65535 transforms are issued in slices (grid.y limit); the pinned zero-copy ends must
    follow every slice (round-1 advisor finding: they were dropped and the call returned uninitialised
    staging memory with status OK)."""
    from ntt_b200.api import FIELDS
    from oracle import c_oracle as co

    torch = torch_cuda
    p, g = FIELDS["bn254_fr"]
    ctx = ctxs("bn254_fr", 24)
    log_n, batch = 3, 70000
    x = co.splitmix_elements(batch << log_n, p, 4321)
    h_in = torch.from_numpy(x.view(np.int32).copy()).pin_memory()
    h_out = torch.zeros_like(h_in).pin_memory()
    got = ctx.forward_host(h_in.numpy().view(np.uint32), log_n, batch=batch, out=h_out.numpy().view(np.uint32)).reshape(-1, 8)
    n = 1 << log_n
    for b in (0, 1, 65534, 65535, 65536, 69999):
        assert np.array_equal(got[b * n:(b + 1) * n], co.ntt(x[b * n:(b + 1) * n], p, g)), b
    # the whole batch against the device-buffer entry point
    assert np.array_equal(got, _gpu_forward(torch, ctx, x, log_n, batch=batch))
    buf = h_out.numpy().view(np.uint32)
    ctx.inverse_host(buf, log_n, batch=batch, out=buf)
    assert np.array_equal(buf.reshape(-1, 8), x)


def test_cpp_program_mirroring_the_reference_main(torch_cuda):
    """tests/cpp/ref_flow_main.cpp re-creates the reference's self-checking main()
    (big-num.cu:369-488: x[i] = i, CPU NTT, NTT_GPU_Naive, compare, NTT_GZKP, compare, sizes
    2^5..) in C++ against include/ntt256_compat.hpp, which keeps the reference's driver
    signatures.  Full 256-bit comparison, three fields, exit code 0 only without mismatches."""
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "tests", "cpp", "ref_flow_main")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-C", os.path.join(root, "tests", "cpp")])
    out = subprocess.run([exe, "20"], capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "correct (0 mismatches)" in out.stdout


@pytest.mark.parametrize("log_n,world", [(12, 1), (16, 2), (20, 8), (21, 4)])
def test_distributed_inverse_ranks_emulated_on_one_gpu(torch_cuda, ctxs, log_n, world):
    """Inverse four-step on the forward's output shards must return the forward's input shards
    (values = oracle inverse NTT), every rank emulated on this GPU, exchange by its host definition."""
    from ntt_b200 import dist as nd
    from ntt_b200.api import FIELDS
    from oracle import c_oracle as co

    torch = torch_cuda
    p, g = FIELDS["bn254_fr"]
    ctx = ctxs("bn254_fr", 24)
    log_p, log_r = world.bit_length() - 1, nd.default_log_r(log_n)
    x = co.splitmix_elements(1 << log_n, p, 0x4E5454 + 70 + log_n)
    X = co.ntt(x, p, g, threads=8)
    s1 = []
    for r, sh in enumerate(nd.scatter_output_layout(X, log_n, world, log_r)):
        d = torch.from_numpy(sh.view(np.int32)).cuda().view(-1)
        ctx.fourstep_inverse_step1(d, log_n, log_r, log_p, r)
        torch.cuda.synchronize()
        s1.append(d.cpu().numpy().view(np.uint32).reshape(-1, 8))
    outs = []
    for r, sh in enumerate(nd.exchange_reference(s1, world)):
        d = torch.from_numpy(np.ascontiguousarray(sh).view(np.int32)).cuda().view(-1)
        ctx.fourstep_inverse_step2(d, log_n, log_r, log_p, r)
        torch.cuda.synchronize()
        outs.append(d.cpu().numpy().view(np.uint32).reshape(-1, 8))
    want = nd.scatter_input(x, log_n, world, log_r)
    for r in range(world):
        assert np.array_equal(outs[r], want[r]), r
    assert np.array_equal(nd.gather_input_layout(outs, log_n, world, log_r), x)


def test_repeated_runs_are_bitwise_identical(torch_cuda, ctxs):
    """Race smoke test (compute-sanitizer is closed on this pool): many repetitions under several
    tilings, streams in flight concurrently, must reproduce the oracle every time."""
    from ntt_b200.api import FIELDS, Ntt256
    from oracle import c_oracle as co

    torch = torch_cuda
    p, g = FIELDS["bn254_fr"]
    ctx = Ntt256("bn254_fr", max_log_n=20)
    for tile_bits, coalesce_bits in ((11, 2), (9, 1), (12, 3)):
        ctx.set_tuning(tile_bits, coalesce_bits)
        for log_n in (13, 17, 20):
            x = co.splitmix_elements(1 << log_n, p, 8080 + log_n)
            want = torch.from_numpy(co.ntt(x, p, g, threads=8).view(np.int32)).cuda().view(-1)
            src = torch.from_numpy(x.view(np.int32)).cuda().view(-1)
            streams = [torch.cuda.Stream() for _ in range(3)]
            bufs = [torch.empty_like(src) for _ in range(3)]
            bad = 0
            for rep in range(12):
                for s, b in zip(streams, bufs):
                    s.wait_stream(torch.cuda.current_stream())
                    with torch.cuda.stream(s):
                        b.copy_(src)
                        ctx.forward(b, log_n)
                        bad += int(not torch.equal(b, want))
            torch.cuda.synchronize()
            assert bad == 0, (tile_bits, log_n, bad)
    ctx.close()


def test_2p26_properties(torch_cuda):
    """2 GiB of data, 5 passes, a 2 GiB streamed first-pass twiddle array (the two-lookup alternative
    is covered by test_streamed_and_two_lookup_twiddles_agree): one output sampled by Horner on the
    CPU, and the round trip on the device."""
    from ntt_b200.api import FIELDS, Ntt256
    from oracle import c_oracle as co

    torch = torch_cuda
    p, g = FIELDS["bn254_fr"]
    ctx = Ntt256("bn254_fr", max_log_n=26)
    log_n = 26
    x = co.splitmix_elements(1 << log_n, p, 2626)
    src = torch.from_numpy(x.view(np.int32)).cuda().view(-1)
    d = src.clone()
    ctx.forward(d, log_n)
    torch.cuda.synchronize()
    k = (1 << 25) + 12345
    got = d[8 * k: 8 * k + 8].cpu().numpy().view(np.uint32)
    assert co.dft_sample(x, p, g, k) == sum(int(got[i]) << (32 * i) for i in range(8))
    ctx.inverse(d, log_n)
    torch.cuda.synchronize()
    assert torch.equal(d, src)
    ctx.close()


def test_maximum_size_2p28_closed_form(torch_cuda):
    """BASELINE config 5's size on ONE GPU, the largest transform BN254 Fr admits (two-adicity 28;
    8 GiB of data).  Dense input with a closed form: x[j] = c*h^j (made by the coset entry point from
    a constant array), so X[k] = c*(h^n - 1)/(h*w_n^k - 1) -- every output is checked-able exactly
    without a second 2^28-point transform; sampled k, then the inverse restores the constant array."""
    from ntt_b200.api import FIELDS, Ntt256

    torch = torch_cuda
    free_b, _ = torch.cuda.mem_get_info()
    if free_b < (40 << 30):
        pytest.skip("needs ~34 GiB of device memory")
    p, g = FIELDS["bn254_fr"]
    log_n = 28
    n = 1 << log_n
    c, h = 0x1234567890ABCDEF1122334455667788 % p, 7
    ctx = Ntt256("bn254_fr", max_log_n=log_n)
    limbs = torch.tensor([(c >> (32 * k)) & 0xFFFFFFFF for k in range(8)], dtype=torch.int64).to(torch.int32)
    src = limbs.cuda().repeat(n)  # 8 GiB
    d = src.clone()
    ctx.coset_forward(d, log_n, h)
    torch.cuda.synchronize()
    w = pow(g, (p - 1) >> log_n, p)
    num = c * (pow(h, n, p) - 1) % p
    for k in (0, 1, 2, n // 3, n // 2, n - 1, 0x0ABCDEF, 0xFEDCBA9):
        got = d[8 * k: 8 * k + 8].cpu().numpy().view(np.uint32)
        want = num * pow((h * pow(w, k, p) - 1) % p, p - 2, p) % p
        assert sum(int(got[i]) << (32 * i) for i in range(8)) == want, k
    ctx.coset_inverse(d, log_n, h)
    torch.cuda.synchronize()
    assert torch.equal(d, src)
    ctx.close()


# --------------------------------------------------------------------------- small-transform kernel
@pytest.mark.parametrize("field", ["bn254_fr", "ref29"])
def test_small_transform_kernel_matches_throughput_kernel_and_oracle(torch_cuda, field):
    """Calls of <= 2^16 / 2^17 elements run on ntt_pass2_kernel / ntt_pass4_kernel (two / four elements
    per thread); the same calls with that path disabled run on ntt_pass_kernel.  All must give the
    oracle's bits."""
    from ntt_b200.api import FIELDS, Ntt256
    from oracle import c_oracle as co

    p, g = FIELDS[field]
    small = Ntt256(field, max_log_n=18)
    big = Ntt256(field, max_log_n=18)
    big.set_small_transform_limit(0)
    try:
        for log_n, batch in ((4, 1), (5, 3), (7, 1), (9, 5), (11, 1), (12, 2), (13, 1), (14, 16), (16, 1), (16, 2), (13, 16), (17, 1),
                             (18, 1), (10, 256)):
            n = 1 << log_n
            x = co.splitmix_elements(n * batch, p, 0x5A11 + log_n)
            want = np.concatenate([co.ntt(x[b * n:(b + 1) * n], p, g, threads=4) for b in range(batch)])
            for ctx in (small, big):
                before = ctx.launch_count
                got = _gpu_forward(torch_cuda, ctx, x, log_n, batch)
                assert np.array_equal(got, want), (log_n, batch, ctx is small)
                assert np.array_equal(_gpu_forward(torch_cuda, ctx, got, log_n, batch, inverse=True), x)
                assert ctx.launch_count > before
    finally:
        small.close()
        big.close()


def test_transform_is_capturable_in_a_cuda_graph(torch_cuda, ctxs):
    """After one warm-up call (lazy tables are built outside capture) a forward+inverse pair can be
    captured in a CUDA graph and replayed: launch-bound callers of small transforms replay graphs."""
    from ntt_b200.api import FIELDS
    from oracle import c_oracle as co

    torch = torch_cuda
    p, g = FIELDS["bn254_fr"]
    ctx = ctxs("bn254_fr", 20)
    log_n, batch = 12, 3
    n = 1 << log_n
    x = co.splitmix_elements(n * batch, p, 0x6A4)
    want = np.concatenate([co.ntt(x[b * n:(b + 1) * n], p, g, threads=4) for b in range(batch)])
    d = torch.from_numpy(np.array(x, dtype=np.uint32, copy=True).view(np.int32)).cuda()
    ctx.forward(d, log_n, batch)
    ctx.inverse(d, log_n, batch)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        ctx.forward(d, log_n, batch)
    for _ in range(3):  # forward three times = replay thrice; undo two of them
        graph.replay()
    ctx.inverse(d, log_n, batch)
    ctx.inverse(d, log_n, batch)
    torch.cuda.synchronize()
    assert np.array_equal(d.cpu().numpy().view(np.uint32).reshape(-1, 8), want)


def test_u64_limb_view_is_the_same_bytes(torch_cuda, ctxs):
    """prover stacks hold field elements as 4 x u64 little-endian limbs; on a
    little-endian host that is byte-for-byte the 8 x u32 layout, so the same buffer goes in."""
    from ntt_b200.api import FIELDS
    from oracle import py_oracle as po

    torch = torch_cuda
    p, g = FIELDS["bls12_381_fr"]
    ctx = ctxs("bls12_381_fr", 20)
    log_n = 9
    rng = np.random.default_rng(11)
    vals = [int.from_bytes(rng.bytes(32), "little") % p for _ in range(1 << log_n)]
    x64 = np.array([[(v >> (64 * k)) & (2**64 - 1) for k in range(4)] for v in vals], dtype=np.uint64)
    d = torch.from_numpy(x64.view(np.int64)).cuda()          # int64[n, 4] on the device
    ctx.forward(d, log_n)
    torch.cuda.synchronize()
    got64 = d.cpu().numpy().view(np.uint64)
    got = [sum(int(got64[i, k]) << (64 * k) for k in range(4)) for i in range(1 << log_n)]
    assert got == po.ntt_ref(vals, po.BLS12_381_FR)


# --------------------------------------------------------------------------- the reference's GPU drivers
def test_library_matches_the_reference_gpu_golden_vectors(torch_cuda, ctxs):
    """tests/golden/ref256_gpu_kat.json (the reference's own NTT_GPU_Naive / NTT_GZKP run on a B200
    for p = 469762049, BN254 Fr, BLS12-381 Fr): the CUDA path must give the same bytes."""
    from ntt_b200.api import FIELDS
    from oracle import c_oracle as co

    with open(os.path.join(GOLD, "ref256_gpu_kat.json")) as f:
        kat = json.load(f)
    for case in kat["cases"]:
        p, g = FIELDS[case["field"]]
        n = 1 << case["log_n"]
        if case["input"] == "iota":
            x = np.zeros((n, 8), dtype=np.uint32)
            x[:, 0] = np.arange(n, dtype=np.uint32)
        else:
            x = co.splitmix_elements(n, p, case["seed"])
        got = _gpu_forward(torch_cuda, ctxs(case["field"], 20), x, case["log_n"])
        assert hashlib.sha256(got.tobytes()).hexdigest() == case["sha256_gzkp"], (case["field"], case["log_n"], case["input"])


def test_library_against_the_reference_gpu_drivers_live(torch_cuda, ctxs):
    """When oracle/_ref/libref256.so travelled to the box (built where /root/reference exists): run the
    reference's NTT_GPU_Naive and NTT_GZKP and this library side by side on fresh random inputs."""
    from ntt_b200.api import FIELDS
    from oracle import c_oracle as co

    if co.ref256_lib() is None:
        pytest.skip("oracle/_ref/libref256.so not present (built only where /root/reference exists)")
    rng = np.random.default_rng()
    for field in ("bn254_fr", "bls12_381_fr"):
        p, g = FIELDS[field]
        for log_n in (5, 9, 13, 17):
            x = co.splitmix_elements(1 << log_n, p, int(rng.integers(1, 2**31)))
            ours = _gpu_forward(torch_cuda, ctxs(field, 20), x, log_n)
            for driver in (0, 1):
                assert np.array_equal(co.ref256_ntt(x, log_n, p, g, driver), ours), (field, log_n, driver)
