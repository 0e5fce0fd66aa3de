"""TEST INFRASTRUCTURE ONLY -- ctypes loader for the C oracle (oracle/liboracle.so) and, when
present, the reference's own CPU functions (oracle/_ref/libref_i64.so).  Never imported by
ntt_b200/."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_REF = None


def build(force: bool = False) -> None:
    """Compile liboracle.so (and oracle/_ref when /root/reference exists)."""
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("ntt_oracle.c", "ntt_oracle_n.c", "ntt_oracle64.c")]
    stale = os.path.exists(so) and any(os.path.exists(f) and os.path.getmtime(f) > os.path.getmtime(so) for f in srcs)
    if force or stale or not os.path.exists(so):
        subprocess.check_call(["make", "-C", _HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    if os.path.isdir("/root/reference/src") and (
        force or not os.path.exists(os.path.join(_HERE, "_ref", "libref_i64.so"))
    ):
        subprocess.check_call(["make", "-C", _HERE, "ref"], stdout=subprocess.DEVNULL)


def lib() -> ctypes.CDLL:
    global _LIB
    if _LIB is None:
        build()
        L = ctypes.CDLL(os.path.join(_HERE, "liboracle.so"))
        u32p = ctypes.POINTER(ctypes.c_uint32)
        L.oracle_ntt256.argtypes = [u32p, ctypes.c_int, u32p, u32p, ctypes.c_int, ctypes.c_int]
        L.oracle_ntt256.restype = ctypes.c_int
        L.oracle_ntt256_ssip.argtypes = [u32p, ctypes.c_int, u32p, u32p]
        L.oracle_ntt256_ssip.restype = ctypes.c_int
        L.oracle_dft_sample.argtypes = [u32p, u32p, ctypes.c_int, u32p, u32p, ctypes.c_uint64]
        L.oracle_dft_sample.restype = ctypes.c_int
        L.oracle_vec_add.argtypes = [u32p, u32p, u32p, ctypes.c_uint64, u32p]
        L.oracle_splitmix_elements.argtypes = [u32p, ctypes.c_uint64, u32p, ctypes.c_uint64]
        L.oracle_max_threads.restype = ctypes.c_int
        _LIB = L
    return _LIB


def ref_lib():
    """The reference's own compiled CPU NTT (int64, p = 469762049) or None."""
    global _REF
    if _REF is None:
        path = os.path.join(_HERE, "_ref", "libref_i64.so")
        if not os.path.exists(path):
            try:
                build()
            except Exception:
                return None
        if not os.path.exists(path):
            return None
        R = ctypes.CDLL(path)
        i64p = ctypes.POINTER(ctypes.c_longlong)
        for name in ("ref_ntt_i64", "ref_ntt_dif_i64", "ref_ssip_i64"):
            getattr(R, name).argtypes = [i64p, ctypes.c_uint]
        R.ref_modulus.restype = ctypes.c_longlong
        R.ref_generator.restype = ctypes.c_longlong
        _REF = R
    return _REF


def _p(a: np.ndarray):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32))


def limbs(v: int) -> np.ndarray:
    return np.array([(v >> (32 * i)) & 0xFFFFFFFF for i in range(8)], dtype=np.uint32)


def ints_to_array(values) -> np.ndarray:
    buf = b"".join(int(v).to_bytes(32, "little") for v in values)
    return np.frombuffer(buf, dtype=np.uint32).reshape(-1, 8).copy()


def array_to_ints(a: np.ndarray):
    b = np.ascontiguousarray(a, dtype=np.uint32).tobytes()
    return [int.from_bytes(b[i : i + 32], "little") for i in range(0, len(b), 32)]


def ntt(data: np.ndarray, p: int, g: int, inverse: bool = False, threads: int = 1) -> np.ndarray:
    """Out-of-place wrapper around oracle_ntt256 on an [n, 8] uint32 array."""
    out = np.ascontiguousarray(data, dtype=np.uint32).copy()
    n = out.shape[0]
    log_n = n.bit_length() - 1
    assert n == 1 << log_n and out.shape[1] == 8
    rc = lib().oracle_ntt256(_p(out), log_n, _p(limbs(p)), _p(limbs(g)), int(inverse), threads)
    if rc:
        raise ValueError(f"oracle_ntt256 rc={rc}")
    return out


def ntt384(data: np.ndarray, p: int, g: int, inverse: bool = False, threads: int = 1) -> np.ndarray:
    """Out-of-place wrapper around oracle_ntt384 (ntt_oracle_n.c, W = 6) on an [n, 12] uint32 array."""
    out = np.ascontiguousarray(data, dtype=np.uint32).copy()
    n = out.shape[0]
    log_n = n.bit_length() - 1
    assert n == 1 << log_n and out.shape[1] == 12
    L = lib()
    u32p = ctypes.POINTER(ctypes.c_uint32)
    L.oracle_ntt384.argtypes = [u32p, ctypes.c_int, u32p, u32p, ctypes.c_int, ctypes.c_int]
    pl = np.array([(p >> (32 * k)) & 0xFFFFFFFF for k in range(12)], dtype=np.uint32)
    gl = np.array([(g >> (32 * k)) & 0xFFFFFFFF for k in range(12)], dtype=np.uint32)
    rc = L.oracle_ntt384(_p(out), log_n, _p(pl), _p(gl), int(inverse), threads)
    if rc:
        raise ValueError(f"oracle_ntt384 rc={rc}")
    return out


def ntt64(data: np.ndarray, p: int, g: int, inverse: bool = False, threads: int = 1) -> np.ndarray:
    """Out-of-place wrapper around oracle_ntt64 (ntt_oracle64.c) on a uint64 array of 2^log_n elements."""
    L = lib()
    out = np.array(data, dtype=np.uint64, copy=True).reshape(-1)
    log_n = int(out.size).bit_length() - 1
    assert out.size == 1 << log_n
    u64p = ctypes.POINTER(ctypes.c_uint64)
    L.oracle_ntt64.argtypes = [u64p, ctypes.c_int, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_int, ctypes.c_int]
    L.oracle_ntt64.restype = ctypes.c_int
    rc = L.oracle_ntt64(out.ctypes.data_as(u64p), log_n, p, g, int(inverse), threads)
    if rc:
        raise ValueError(f"oracle_ntt64 rc={rc}")
    return out


def splitmix64_elements(n: int, p: int, seed: int) -> np.ndarray:
    """n seeded canonical residues below p (64-bit fields)."""
    L = lib()
    out = np.empty(n, dtype=np.uint64)
    u64p = ctypes.POINTER(ctypes.c_uint64)
    L.oracle_splitmix64_elements.argtypes = [u64p, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint64]
    L.oracle_splitmix64_elements.restype = None
    L.oracle_splitmix64_elements(out.ctypes.data_as(u64p), n, p, seed)
    return out


def ntt_ssip(data: np.ndarray, p: int, g: int) -> np.ndarray:
    out = np.ascontiguousarray(data, dtype=np.uint32).copy()
    log_n = out.shape[0].bit_length() - 1
    rc = lib().oracle_ntt256_ssip(_p(out), log_n, _p(limbs(p)), _p(limbs(g)))
    if rc:
        raise ValueError(f"oracle_ntt256_ssip rc={rc}")
    return out


def dft_sample(data: np.ndarray, p: int, g: int, k: int) -> int:
    data = np.ascontiguousarray(data, dtype=np.uint32)
    out = np.zeros(8, dtype=np.uint32)
    log_n = data.shape[0].bit_length() - 1
    rc = lib().oracle_dft_sample(_p(out), _p(data), log_n, _p(limbs(p)), _p(limbs(g)), k)
    if rc:
        raise ValueError("oracle_dft_sample failed")
    return sum(int(out[i]) << (32 * i) for i in range(8))


def vec_add(a: np.ndarray, b: np.ndarray, p: int) -> np.ndarray:
    r = np.empty_like(a)
    lib().oracle_vec_add(_p(r), _p(a), _p(b), a.shape[0], _p(limbs(p)))
    return r


def splitmix_elements(n: int, p: int, seed: int) -> np.ndarray:
    out = np.empty((n, 8), dtype=np.uint32)
    lib().oracle_splitmix_elements(_p(out), n, _p(limbs(p)), seed)
    return out


def max_threads() -> int:
    return lib().oracle_max_threads()


_REF256 = None


def ref256_lib():
    """The reference's own 256-bit GPU drivers (NTT_GPU_Naive / NTT_GZKP of big-num.cu, compiled
    unmodified: oracle/ref256_harness.cu) or None when oracle/_ref/libref256.so is absent.  Needs a GPU."""
    global _REF256
    if _REF256 is None:
        path = os.path.join(_HERE, "_ref", "libref256.so")
        if not os.path.exists(path):
            return None
        R = ctypes.CDLL(path)
        u32p = ctypes.POINTER(ctypes.c_uint32)
        R.ref256_ntt.argtypes = [u32p, ctypes.c_uint, u32p, u32p, ctypes.c_int]
        _REF256 = R
    return _REF256


def ref256_ntt(x: np.ndarray, log_n: int, p: int, g: int, driver: int = 1) -> np.ndarray:
    """Forward NTT of uint32[n, 8] through the reference driver (0 = NTT_GPU_Naive, 1 = NTT_GZKP)."""
    a = np.array(x, dtype=np.uint32, copy=True)
    rc = ref256_lib().ref256_ntt(_p(a), log_n, _p(limbs(p)), _p(limbs(g)), driver)
    if rc:
        raise RuntimeError(f"reference driver failed: CUDA error {rc}")
    return a
