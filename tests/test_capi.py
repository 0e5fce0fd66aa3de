"""The C-ABI library loads on a CPU-only box, exports every symbol include/ntt256.h declares,
validates arguments before touching CUDA, and fails loudly (no CPU fallback) when no device
can run it."""
import ctypes
import os
import re

import numpy as np
import pytest

from ntt_b200 import lib as L_mod  # noqa: F401  (package re-exports the loader)
from ntt_b200.lib import EXPORTS, lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_exports_match_header(built):
    with open(os.path.join(ROOT, "include", "ntt256.h")) as f:
        header = f.read()
    declared = sorted(set(re.findall(r"\b(ntt256_[a-z0-9_]+)\s*\(", header)))
    assert declared == sorted(EXPORTS)
    L = lib()
    for name in declared:
        assert hasattr(L, name), name


def test_exports_match_header_384(built):
    from ntt_b200.lib import EXPORTS384

    with open(os.path.join(ROOT, "include", "ntt384.h")) as f:
        header = f.read()
    declared = sorted(set(re.findall(r"\b(ntt384_[a-z0-9_]+)\s*\(", header)))
    assert declared == sorted(EXPORTS384)
    L = lib()
    for name in declared:
        assert hasattr(L, name), name


def test_argument_and_field_validation_384(built):
    from ntt_b200.api import FIELDS384
    from ntt_b200.lib import ERR_ARG, ERR_FIELD

    L = lib()
    u32p = ctypes.POINTER(ctypes.c_uint32)

    def create(p, g, max_log_n):
        h = ctypes.c_void_p()
        pl = np.array([(p >> (32 * i)) & 0xFFFFFFFF for i in range(12)], dtype=np.uint32)
        gl = np.array([(g >> (32 * i)) & 0xFFFFFFFF for i in range(12)], dtype=np.uint32)
        return L.ntt384_ctx_create(ctypes.byref(h), 0, pl.ctypes.data_as(u32p), gl.ctypes.data_as(u32p), max_log_n)

    q, g = FIELDS384["bls12_377_fq"]
    assert create(q, g, 0) == ERR_ARG
    assert create(q, g, 31) == ERR_ARG
    assert create(q - 1, g, 10) == ERR_FIELD            # even
    assert create((1 << 383) + 1, g, 10) == ERR_FIELD    # no lazy-reduction headroom
    assert create(q, 4, 10) == ERR_FIELD                 # a square is not a generator
    assert create(q, g, 47) == ERR_ARG                   # above the API's 2^30 limit (2-adicity is 46)
    assert L.ntt384_forward(None, None, 3, None) == ERR_ARG


def test_exports_match_header_64(built):
    from ntt_b200.lib import EXPORTS64

    with open(os.path.join(ROOT, "include", "ntt64.h")) as f:
        header = f.read()
    declared = sorted(set(re.findall(r"\b(ntt64_[a-z0-9_]+)\s*\(", header)))
    assert declared == sorted(EXPORTS64)
    L = lib()
    for name in declared:
        assert hasattr(L, name), name


def test_argument_and_field_validation_64(built):
    from ntt_b200.lib import ERR_ARG, ERR_FIELD

    L = lib()
    GL = 0xFFFFFFFF00000001

    def create(p, g, max_log_n):
        h = ctypes.c_void_p()
        return L.ntt64_ctx_create(ctypes.byref(h), 0, p, g, max_log_n)

    assert create(GL, 7, 0) == ERR_ARG
    assert create(GL, 7, 33) == ERR_ARG                  # two-adicity of Goldilocks is 32
    assert create(GL - 1, 7, 10) == ERR_FIELD            # even
    assert create((1 << 63) + 1025, 3, 10) == ERR_FIELD  # the Montgomery instance needs p < 2^63
    assert create(469762049, 3, 27) == ERR_FIELD         # 2^26 | p - 1 only
    assert create(469762049, 4, 10) == ERR_FIELD         # a square is not a generator
    assert create(GL, 0, 10) == ERR_FIELD
    assert L.ntt64_forward(None, None, 3, None) == ERR_ARG
    assert L.ntt64_vec_mul(None, None, None, None, 4, None) == ERR_ARG


def test_strings(built):
    L = lib()
    assert L.ntt256_strerror(0) == b"ok"
    assert b"argument" in L.ntt256_strerror(1)
    assert b"sm_100a" in L.ntt256_version()


def _limbs(v):
    return np.array([(v >> (32 * i)) & 0xFFFFFFFF for i in range(8)], dtype=np.uint32)


def _create(p, g, max_log_n):
    L = lib()
    h = ctypes.c_void_p()
    u32p = ctypes.POINTER(ctypes.c_uint32)
    rc = L.ntt256_ctx_create(ctypes.byref(h), 0, _limbs(p).ctypes.data_as(u32p), _limbs(g).ctypes.data_as(u32p), max_log_n)
    return rc, h


def test_argument_and_field_validation(built):
    from ntt_b200.lib import ERR_ARG, ERR_FIELD

    bn = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
    assert _create(bn, 5, 0)[0] == ERR_ARG
    assert _create(bn, 5, 33)[0] == ERR_ARG
    assert _create(bn - 1, 5, 10)[0] == ERR_FIELD            # even modulus
    assert _create((1 << 255) + 1, 5, 10)[0] == ERR_FIELD     # p >= 2^255: no lazy-reduction headroom
    assert _create(bn, 5, 29)[0] == ERR_FIELD                 # two-adicity of BN254 Fr is 28
    assert _create(bn, 0, 10)[0] == ERR_FIELD
    assert _create(bn, 4, 10)[0] == ERR_FIELD                 # 4 is a square: w_N^(N/2) != -1
    L = lib()
    assert L.ntt256_forward(None, None, 3, None) == ERR_ARG
    assert L.ntt256_ctx_create_named(ctypes.byref(ctypes.c_void_p()), 0, 7, 10) == ERR_ARG


def test_no_cpu_fallback(built):
    """Without a CUDA device context creation must fail with ERR_CUDA, never compute on the host."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from ntt_b200.api import Ntt256
    from ntt_b200.lib import ERR_CUDA, Ntt256Error

    with pytest.raises(Ntt256Error) as e:
        Ntt256("bn254_fr", max_log_n=12)
    assert e.value.status == ERR_CUDA


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under ntt_b200/ may reference it."""
    pkg = os.path.join(ROOT, "ntt_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                with open(os.path.join(dirpath, fn)) as f:
                    src = f.read()
                assert "oracle" not in src.replace("CPU oracle NTT()", ""), f"{fn} mentions the oracle"


def test_host_field_arithmetic_cpp(built):
    """tests/cpp/hostfield_test.cpp: the host-side constant derivation (R, R^2, -p^-1, powers,
    inverses) of both widths -- the generic W-limb field equals the 256-bit one at W = 4 and satisfies
    the field identities over BLS12-377 Fq at W = 6."""
    import subprocess

    d = os.path.join(ROOT, "tests", "cpp")
    subprocess.check_call(["make", "-C", d, "hostfield_test"], stdout=subprocess.DEVNULL)
    out = subprocess.run([os.path.join(d, "hostfield_test")], capture_output=True, text=True)
    assert out.returncode == 0 and "HOSTFIELD_OK" in out.stdout, out.stdout


def test_pass_index_tables_cpp(built):
    """tests/cpp/pass_tables_test.cpp: the per-plan index tables the pass kernels look up instead of depositing
    bits (csrc/pass_tables.h) describe, for every kernel shape and for the four-step pieces, a partition of the
    data -- each element loaded once and stored once per pass, in place -- and carry the swizzle of the index they
    name.  Host arithmetic only."""
    import subprocess

    d = os.path.join(ROOT, "tests", "cpp")
    subprocess.check_call(["make", "-C", d, "pass_tables_test"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    out = subprocess.run([os.path.join(d, "pass_tables_test")], capture_output=True, text=True)
    assert out.returncode == 0 and "0 failures" in out.stdout, out.stdout[-2000:]


def test_headers_are_plain_c():
    """The drop-in boundary is a C ABI: every public header compiles as C99 on its own (no C++ or torch types in a
    signature), and the 64-bit shim of INTEGRATION.md compiles against it."""
    import subprocess
    import tempfile

    inc = os.path.join(ROOT, "include")
    for h in ("ntt256.h", "ntt384.h", "ntt64.h"):
        subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-x", "c", os.path.join(inc, h)])
    shim = r'''
#include <stdlib.h>
#include "ntt64.h"
void NTT_gpu(long long data[], long long reverse[], long long len, long long omega) {
  static ntt64_ctx* ctx = 0;
  int log_n = 0;
  (void)reverse;
  if (!ctx && ntt64_ctx_create(&ctx, 0, 469762049ull, (uint64_t)omega, 26)) abort();
  while ((1ll << log_n) < len) ++log_n;
  if (ntt64_forward_host(ctx, data, data, log_n, 1)) abort();
}
'''
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "shim.c")
        with open(src, "w") as f:
            f.write(shim)
        subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", inc, "-c", src, "-o", os.path.join(d, "shim.o")])
