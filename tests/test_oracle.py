"""The oracle against every vector that pins it: the reference's own compiled CPU code
(oracle/_ref, when present), golden vectors generated from it, the survey-time anchors, the
DFT definition and the second (self-sort-in-place) statement of the algorithm."""
import ctypes
import hashlib
import json
import os

import numpy as np
import pytest

from oracle import c_oracle as co
from oracle import py_oracle as po

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load(name):
    with open(os.path.join(GOLD, name)) as f:
        return json.load(f)


def test_ref29_survey_kats(built):
    a = _load("anchors.json")["ref29"]
    F = po.REF29
    assert po.ntt_ref(list(range(8)), F) == a["8"]
    assert po.ntt_ref(list(range(16)), F)[:8] == a["16_head"]
    assert po.ntt_ref(list(range(1024)), F)[:8] == a["1024_head"]
    assert F.omega(256) == a["omega_256"]  # also hard-coded in /root/reference/src/twiddlecheck.py:11


@pytest.mark.parametrize("impl", ["py", "c", "c_mt", "c_ssip", "py_ssip"])
def test_reference_golden_vectors(built, impl):
    """Outputs of the reference's own NTT() (p = 469762049, x[j] = j), tests/golden/ref_i64_kat.json."""
    g = _load("ref_i64_kat.json")
    F = po.Field("ref", g["p"], g["g"])
    for case in g["cases"]:
        log_n = case["log_n"]
        n = 1 << log_n
        if impl.startswith("py") and log_n > 12:
            continue
        x = list(range(n))
        if impl == "py":
            X = po.ntt_ref(x, F)
        elif impl == "py_ssip":
            X = po.ntt_ssip(x, F)
        elif impl == "c":
            X = co.array_to_ints(co.ntt(co.ints_to_array(x), F.p, F.g))
        elif impl == "c_mt":
            X = co.array_to_ints(co.ntt(co.ints_to_array(x), F.p, F.g, threads=4))
        else:
            X = co.array_to_ints(co.ntt_ssip(co.ints_to_array(x), F.p, F.g))
        assert X[:8] == case["head"][: min(8, n)]
        assert X[-4:] == case["tail"][-min(4, n):]
        assert hashlib.sha256(np.array(X, dtype=np.int64).tobytes()).hexdigest() == case["sha256_i64le"]
        if "full" in case:
            assert X == case["full"]


def test_against_compiled_reference(built):
    """Live comparison with the reference's compiled CPU functions (only where oracle/_ref exists)."""
    R = co.ref_lib()
    if R is None:
        pytest.skip("oracle/_ref not built (no /root/reference on this box)")
    i64p = ctypes.POINTER(ctypes.c_longlong)
    rng = np.random.default_rng(5)
    P = int(R.ref_modulus())
    for log_n in (1, 2, 3, 7, 10, 13):
        n = 1 << log_n
        # NOTE the reference's harness input is x[j] = j; random inputs are run through the same code
        x = rng.integers(0, P, size=n, dtype=np.int64)
        a = x.copy()
        R.ref_ntt_i64(a.ctypes.data_as(i64p), log_n)
        b = x.copy()
        R.ref_ssip_i64(b.ctypes.data_as(i64p), log_n)
        c = co.array_to_ints(co.ntt(co.ints_to_array(x.tolist()), P, int(R.ref_generator())))
        assert a.tolist() == c
        assert b.tolist() == c


def test_oracle64_reproduces_the_reference_golden_vectors(built):
    """ntt_oracle64.c (the checker of the 64-bit GPU path) on the outputs of the reference's own NTT()
    (p = 469762049, x[j] = j), single- and multi-threaded, and its inverse."""
    g = _load("ref_i64_kat.json")
    for case in g["cases"]:
        n = 1 << case["log_n"]
        x = np.arange(n, dtype=np.uint64)
        for threads in (1, 4):
            X = co.ntt64(x, g["p"], g["g"], threads=threads)
            assert hashlib.sha256(X.astype("<i8").tobytes()).hexdigest() == case["sha256_i64le"]
            assert X[:8].tolist() == case["head"][: min(8, n)]
        assert np.array_equal(co.ntt64(X, g["p"], g["g"], inverse=True), x)


def test_oracle64_against_compiled_reference_and_python(built):
    """Random inputs: the reference's compiled NTT() (when oracle/_ref exists), the Python oracle and the C
    64-bit oracle agree over the reference's modulus; the Python and C oracles agree over Goldilocks."""
    rng = np.random.default_rng(64)
    R = co.ref_lib()
    i64p = ctypes.POINTER(ctypes.c_longlong)
    P, G = 469762049, 3
    for log_n in (1, 2, 5, 9, 12):
        x = rng.integers(0, P, size=1 << log_n, dtype=np.int64)
        c = co.ntt64(x.astype(np.uint64), P, G, threads=2)
        assert c.tolist() == po.ntt_ref(x.tolist(), po.REF29)
        if R is not None:
            a = x.copy()
            R.ref_ntt_i64(a.ctypes.data_as(i64p), log_n)
            assert a.tolist() == c.tolist()
    GL = po.Field("goldilocks", 0xFFFFFFFF00000001, 7)
    for log_n in (0, 1, 3, 8, 11):
        x = co.splitmix64_elements(1 << log_n, GL.p, 7 + log_n)
        # extreme residues too
        x[: min(4, x.size)] = np.array([GL.p - 1, 0, 0xFFFFFFFF, GL.p - 0xFFFFFFFF], dtype=np.uint64)[: min(4, x.size)]
        c = co.ntt64(x, GL.p, GL.g, threads=3)
        assert [int(v) for v in c] == po.ntt_ref([int(v) for v in x], GL)
        assert np.array_equal(co.ntt64(c, GL.p, GL.g, inverse=True, threads=2), x)


def test_oracle64_reproduces_the_goldilocks_known_answers(built):
    """tests/golden/goldilocks_kat.json (made by the Python oracle, scripts/gen_golden64.py): the C 64-bit oracle --
    the checker of the Goldilocks GPU path -- gives the same digests, single- and multi-threaded."""
    g = _load("goldilocks_kat.json")
    for case in g["cases"]:
        n = 1 << case["log_n"]
        x = np.arange(n, dtype=np.uint64) if case["input"] == "iota" else co.splitmix64_elements(n, g["p"], case["seed"])
        for threads in (1, 3):
            X = co.ntt64(x, g["p"], g["g"], threads=threads)
            assert hashlib.sha256(X.astype("<u8").tobytes()).hexdigest() == case["sha256_u64le"], case
            assert [int(v) for v in X[:4]] == case["head"][: min(4, n)]
            if "full" in case:
                assert [int(v) for v in X] == case["full"]


def test_bigfield_golden(built):
    g = _load("bigfield_kat.json")
    for case in g["cases"]:
        F = po.FIELDS[case["field"]]
        n = 1 << case["log_n"]
        x = list(range(n)) if case["input"] == "iota" else po.splitmix_elements(n, F, case["seed"])
        X = co.ntt(co.ints_to_array(x), F.p, F.g)
        assert hashlib.sha256(X.tobytes()).hexdigest() == case["sha256_image"]
        assert [hex(v) for v in co.array_to_ints(X[:4])] == case["head"]
        if case["log_n"] <= 10:
            assert po.sha256_image(po.ntt_ref(x, F)) == case["sha256_image"]


def test_survey_anchors(built):
    a = _load("anchors.json")
    for name in ("bn254_fr", "bls12_381_fr"):
        F = po.FIELDS[name]
        X8 = po.ntt_ref(list(range(8)), F)
        assert X8[0] == 0x1C and X8[4] == F.p - 4
        assert X8[1] == int(a[name]["n8_X1"], 16)
        X10 = co.ntt(co.ints_to_array(range(1 << 10)), F.p, F.g)
        assert hashlib.sha256(X10.tobytes()).hexdigest() == a[name]["sha256_2^10"]
        X16 = co.ntt(co.ints_to_array(range(1 << 16)), F.p, F.g, threads=4)
        assert hashlib.sha256(X16.tobytes()).hexdigest() == a[name]["sha256_2^16"]
        assert co.array_to_ints(X16[1:2])[0] == int(a[name]["2^16_X1"], 16)


@pytest.mark.parametrize("name", ["ref29", "bn254_fr", "bls12_381_fr"])
def test_definition_and_statements_agree(built, name):
    F = po.FIELDS[name]
    for log_n in range(0, 7):
        x = po.splitmix_elements(1 << log_n, F, 77 + log_n)
        want = po.dft_direct(x, F)
        assert po.ntt_ref(x, F) == want
        if log_n:
            assert po.ntt_ssip(x, F) == want
        arr = co.ints_to_array(x)
        assert co.array_to_ints(co.ntt(arr, F.p, F.g)) == want
        if log_n:
            assert co.array_to_ints(co.ntt_ssip(arr, F.p, F.g)) == want
        assert co.array_to_ints(co.ntt(co.ntt(arr, F.p, F.g), F.p, F.g, inverse=True)) == x
        assert po.ntt_ref(want, F, inverse=True) == x


def test_c_oracle_helpers(built):
    F = po.BLS12_381_FR
    x = co.splitmix_elements(1 << 11, F.p, 9)
    assert co.array_to_ints(x) == po.splitmix_elements(1 << 11, F, 9)
    assert all(v < F.p for v in co.array_to_ints(x))
    X = co.ntt(x, F.p, F.g, threads=3)
    assert np.array_equal(X, co.ntt(x, F.p, F.g, threads=1))
    for k in (0, 1, 1000, 2047):
        assert co.dft_sample(x, F.p, F.g, k) == co.array_to_ints(X[k : k + 1])[0]
    y = co.splitmix_elements(1 << 11, F.p, 10)
    s = co.vec_add(x, y, F.p)
    assert co.array_to_ints(s) == [(a + b) % F.p for a, b in zip(co.array_to_ints(x), co.array_to_ints(y))]
    with pytest.raises(ValueError):
        co.ntt(co.ints_to_array(range(8)), 469762049 * 2 + 1, 3)  # 2^3 does not divide p-1 = 2*469762049


def _ref256_input(case, p):
    n = 1 << case["log_n"]
    if case["input"] == "iota":
        x = np.zeros((n, 8), dtype=np.uint32)
        x[:, 0] = np.arange(n, dtype=np.uint32)
        return x
    return co.splitmix_elements(n, p, case["seed"])


def test_oracle_matches_the_reference_gpu_drivers_on_the_big_fields(built):
    """tests/golden/ref256_gpu_kat.json: outputs of the reference's OWN 256-bit GPU drivers
    (NTT_GPU_Naive / NTT_GZKP of big-num.cu, compiled unmodified, run on a B200 by
    scripts/gen_golden_ref256.py) on p = 469762049, BN254 Fr and BLS12-381 Fr, x[j] = j and random
    inputs, 2^5 .. 2^16 points.  The oracle must reproduce every digest: this is what pins parity on
    the >64-bit fields to the reference's code rather than to a restatement of it."""
    kat = _load("ref256_gpu_kat.json")
    assert len(kat["cases"]) >= 40
    fields = {"ref29": po.REF29, "bn254_fr": po.BN254_FR, "bls12_381_fr": po.BLS12_381_FR}
    for case in kat["cases"]:
        assert case["drivers_agree"] and case["sha256_naive"] == case["sha256_gzkp"]
        F = fields[case["field"]]
        x = _ref256_input(case, F.p)
        got = co.ntt(x, F.p, F.g, threads=4)
        assert hashlib.sha256(got.tobytes()).hexdigest() == case["sha256_naive"], (case["field"], case["log_n"], case["input"])
        assert [hex(sum(int(got[i][k]) << (32 * k) for k in range(8))) for i in range(4)] == case["first"]
        if case["log_n"] <= 8:  # the Python-int statement as well
            xi = [sum(int(x[i][k]) << (32 * k) for k in range(8)) for i in range(len(x))]
            want = po.ntt_ref(xi, F)
            assert [sum(int(got[i][k]) << (32 * k) for k in range(8)) for i in range(len(x))] == want


def _ref512_vals(case, p):
    n = 1 << case["log_n"]
    if case["input"] == "iota":
        return list(range(n))
    rng = np.random.default_rng(case["seed"])
    return [int.from_bytes(rng.bytes(48), "little") % p for _ in range(n)]


def test_oracle_matches_the_reference_gzkp_driver_on_a_377_bit_field(built):
    """tests/golden/ref512_gpu_kat.json: the reference's NTT_GZKP instantiated at 512-bit containers
    (compiled unmodified, oracle/ref512_harness.cu) run on a B200 over BLS12-377 Fq.  The Python
    oracle must reproduce the digests (sizes it reaches in seconds): pins the 384-bit path's oracle."""
    kat = _load("ref512_gpu_kat.json")
    F = po.BLS12_377_FQ
    checked = 0
    for case in kat["cases"]:
        assert case["ours_equals_reference"]
        if case["log_n"] > 10:
            continue
        got = po.ntt_ref(_ref512_vals(case, F.p), F)
        img = b"".join(v.to_bytes(48, "little") for v in got)
        assert hashlib.sha256(img).hexdigest() == case["sha256"], (case["log_n"], case["input"])
        assert [hex(v) for v in got[:3]] == case["first"]
        checked += 1
    assert checked >= 8


def _arr12(vals):
    a = np.zeros((len(vals), 12), dtype=np.uint32)
    for i, v in enumerate(vals):
        for k in range(12):
            a[i, k] = (v >> (32 * k)) & 0xFFFFFFFF
    return a


def test_c_oracle_384_equals_python_and_the_reference_gzkp_digests(built):
    """oracle/ntt_oracle_n.c (W = 6): equal to py_oracle on random data (forward and inverse,
    2^0 .. 2^10), and to EVERY digest of tests/golden/ref512_gpu_kat.json -- including the 2^12 and
    2^14 cases the Python oracle is too slow for."""
    import random

    F = po.BLS12_377_FQ
    rng = random.Random(384)
    for log_n in range(0, 11):
        vals = [rng.randrange(F.p) for _ in range(1 << log_n)]
        fwd = co.ntt384(_arr12(vals), F.p, F.g, threads=1 + log_n % 3)
        assert np.array_equal(fwd, _arr12(po.ntt_ref(vals, F))), log_n
        inv = co.ntt384(_arr12(vals), F.p, F.g, inverse=True, threads=2)
        assert np.array_equal(inv, _arr12(po.ntt_ref(vals, F, inverse=True))), log_n
        assert np.array_equal(co.ntt384(fwd, F.p, F.g, inverse=True), _arr12(vals))
    kat = _load("ref512_gpu_kat.json")
    for case in kat["cases"]:
        got = co.ntt384(_arr12(_ref512_vals(case, F.p)), F.p, F.g, threads=4)
        assert hashlib.sha256(got.tobytes()).hexdigest() == case["sha256"], (case["log_n"], case["input"])
