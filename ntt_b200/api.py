"""Host-side mirror of the reference's 256-bit NTT entry points, over libntt256.so.

The reference's boundary is a pair of C++ host drivers called from main()
(/root/reference/src/big-num.cu:106-107, 260-261):

    NTT_GPU_Naive<tpi,bits>(data_d, len, reverse_d, reverse_len, prime, omega)
    NTT_GZKP<tpi,bits>(data_d, len, reverse_d, reverse_len, prime, omega, B, G)

with `data_d` a device array of cgbn_mem_t<256> transformed in place (natural order in and
out), `prime` / `omega` host values and `omega` the GENERATOR (the callee derives w_n).
`NTT_GZKP` / `NTT_GPU_Naive` below keep those names and argument meanings (the bit-reversal
list and the B/G knobs are accepted and ignored: this implementation has no permutation pass
and picks its own tiling), and `Ntt256` is the context object the C ABI is built around.

torch is used for device memory and streams only; all arithmetic happens in the CUDA library.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Sequence

import numpy as np

from .lib import Ntt256Error, lib as _load_lib

FIELDS = {
    "bn254_fr": (0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001, 5),
    "bls12_381_fr": (0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001, 7),
    "ref29": (469762049, 3),  # the only field the reference itself runs (big-num.cu:9-10)
}


def _limbs(v) -> np.ndarray:
    if isinstance(v, (int, np.integer)):
        v = int(v)
        return np.array([(v >> (32 * i)) & 0xFFFFFFFF for i in range(8)], dtype=np.uint32)
    a = np.ascontiguousarray(v, dtype=np.uint32).reshape(-1)
    if a.size != 8:
        raise ValueError("a field element is 8 little-endian u32 limbs")
    return a


def _u32p(a: np.ndarray):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32))


def _stream_handle(stream) -> Optional[int]:
    if stream is None:
        import torch

        return torch.cuda.current_stream().cuda_stream
    if isinstance(stream, int):
        return stream
    return stream.cuda_stream


def _device_ptr(data, n_elements: int) -> int:
    """data: torch CUDA tensor holding n_elements * 32 contiguous bytes."""
    if not getattr(data, "is_cuda", False):
        raise TypeError("device entry points take a torch CUDA tensor (use *_host for numpy arrays)")
    if not data.is_contiguous():
        raise ValueError("data must be contiguous")
    if data.numel() * data.element_size() != n_elements * 32:
        raise ValueError(f"expected {n_elements} elements of 32 bytes, got {data.numel() * data.element_size()} bytes")
    return data.data_ptr()


class Ntt256:
    """ntt256_ctx: field constants, device twiddle tables and plans for sizes up to 2^max_log_n."""

    def __init__(self, field="bn254_fr", max_log_n: int = 24, device: int = 0, p=None, g=None):
        L = _load_lib()
        self._L = L
        self._h = ctypes.c_void_p()
        if p is None:
            p, g = FIELDS[field]
        self.p, self.g = int(p), int(g)
        self.max_log_n = max_log_n
        self.device = device
        pl, gl = _limbs(self.p), _limbs(self.g)
        rc = L.ntt256_ctx_create(ctypes.byref(self._h), device, _u32p(pl), _u32p(gl), max_log_n)
        if rc:
            raise Ntt256Error(rc)

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._L.ntt256_ctx_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int):
        if rc:
            raise Ntt256Error(rc, self._L.ntt256_last_cuda_error(self._h).decode())

    def set_tuning(self, tile_bits: int, coalesce_bits: int):
        self._check(self._L.ntt256_ctx_set_tuning(self._h, tile_bits, coalesce_bits))

    def set_small_transform_limit(self, max_elements: int):
        """Calls of at most this many elements use the four-elements-per-thread kernel (0: never)."""
        self._check(self._L.ntt256_ctx_set_small_transform_limit(self._h, max_elements))

    def set_elements_per_thread(self, log2_elements: int):
        """Force the pass kernel with 2^log2_elements (1..3) elements per thread; 0 = automatic."""
        self._check(self._L.ntt256_ctx_set_elements_per_thread(self._h, log2_elements))

    def set_dependent_launch(self, enable: bool):
        """Programmatic dependent launch between the pass kernels (default on)."""
        self._check(self._L.ntt256_ctx_set_dependent_launch(self._h, int(bool(enable))))

    def set_constant_field_kernel(self, enable: bool):
        """BN254 Fr: throughput kernel with compile-time field constants (default on) or the generic instance."""
        self._check(self._L.ntt256_ctx_set_constant_field_kernel(self._h, int(bool(enable))))

    def set_twiddle_mode(self, shoup: bool):
        """True: (w, floor(w 2^256/p)) tables and Shoup products (fields with p < 2^254); False: Montgomery."""
        self._check(self._L.ntt256_ctx_set_twiddle_mode(self._h, int(bool(shoup))))

    @property
    def twiddle_mode(self) -> str:
        return "shoup" if self._L.ntt256_ctx_get_twiddle_mode(self._h) else "montgomery"

    def set_peer_sync(self, flag_ptrs=None, rank: int = 0, timeout_s: float = 0.0):
        """Arm (flag_ptrs = every rank's flag buffer as mapped here) or disarm (None) the rank
        synchronisation folded into the four-step kernels (ntt256_ctx_set_peer_sync)."""
        if flag_ptrs is None:
            self._check(self._L.ntt256_ctx_set_peer_sync(self._h, None, 0, 0, 0.0))
        else:
            arr = (ctypes.c_void_p * len(flag_ptrs))(*[int(v) for v in flag_ptrs])
            self._check(self._L.ntt256_ctx_set_peer_sync(self._h, arr, len(flag_ptrs), rank, float(timeout_s)))

    def set_stream_twiddle_limit(self, max_bytes: int):
        """Largest streamed first-pass twiddle array built automatically (0: always two lookups)."""
        self._check(self._L.ntt256_ctx_set_stream_twiddle_limit(self._h, max_bytes))

    def set_stream_twiddle_budget(self, total_bytes: int):
        """All streamed twiddle arrays together stay below this; least recently used ones are dropped first."""
        self._check(self._L.ntt256_ctx_set_stream_twiddle_budget(self._h, total_bytes))

    def table_bytes(self) -> tuple:
        """(device bytes held in twiddle / coset tables, of which streamed first-pass arrays)."""
        st = ctypes.c_size_t(0)
        tot = int(self._L.ntt256_ctx_table_bytes(self._h, ctypes.byref(st)))
        return tot, int(st.value)

    @property
    def launch_count(self) -> int:
        return int(self._L.ntt256_launch_count(self._h))

    # ---- device entry points (in place, asynchronous on `stream`)
    def forward(self, data, log_n: int, batch: int = 1, stream=None):
        ptr = _device_ptr(data, batch << log_n)
        self._check(self._L.ntt256_forward_batch(self._h, ptr, log_n, batch, _stream_handle(stream)))
        return data

    def inverse(self, data, log_n: int, batch: int = 1, stream=None):
        ptr = _device_ptr(data, batch << log_n)
        self._check(self._L.ntt256_inverse_batch(self._h, ptr, log_n, batch, _stream_handle(stream)))
        return data

    # ---- companions (SURVEY.md par. 8f)
    def vec_mul(self, a, b, out=None, mont_form: bool = False, stream=None):
        """out = a*b element-wise (canonical data) or the Montgomery product (mont_form=True)."""
        n = a.numel() * a.element_size() // 32
        out = a if out is None else out
        pa, pb, po = _device_ptr(a, n), _device_ptr(b, n), _device_ptr(out, n)
        self._check(self._L.ntt256_vec_mul(self._h, po, pa, pb, n, int(mont_form), _stream_handle(stream)))
        return out

    def to_montgomery(self, data, stream=None):
        n = data.numel() * data.element_size() // 32
        self._check(self._L.ntt256_to_montgomery(self._h, _device_ptr(data, n), n, _stream_handle(stream)))
        return data

    def from_montgomery(self, data, stream=None):
        n = data.numel() * data.element_size() // 32
        self._check(self._L.ntt256_from_montgomery(self._h, _device_ptr(data, n), n, _stream_handle(stream)))
        return data

    def coset_forward(self, data, log_n: int, shift: int, batch: int = 1, stream=None):
        ptr = _device_ptr(data, batch << log_n)
        self._check(self._L.ntt256_coset_forward(self._h, ptr, log_n, batch, _u32p(_limbs(shift)), _stream_handle(stream)))
        return data

    def coset_inverse(self, data, log_n: int, shift: int, batch: int = 1, stream=None):
        ptr = _device_ptr(data, batch << log_n)
        self._check(self._L.ntt256_coset_inverse(self._h, ptr, log_n, batch, _u32p(_limbs(shift)), _stream_handle(stream)))
        return data

    def forward_mul(self, data, mul_mont, log_n: int, batch: int = 1, stream=None):
        """data <- NTT(data) * mul (element-wise); `mul_mont` holds batch * 2^log_n elements in Montgomery form."""
        n = batch << log_n
        self._check(self._L.ntt256_forward_mul(self._h, _device_ptr(data, n), _device_ptr(mul_mont, n), log_n, batch,
                                               _stream_handle(stream)))
        return data

    def inverse_mul(self, data, mul_mont, log_n: int, batch: int = 1, stream=None):
        n = batch << log_n
        self._check(self._L.ntt256_inverse_mul(self._h, _device_ptr(data, n), _device_ptr(mul_mont, n), log_n, batch,
                                               _stream_handle(stream)))
        return data

    def set_dense_pass4(self, enable: bool):
        self._check(self._L.ntt256_ctx_set_dense_pass4(self._h, int(bool(enable))))

    def set_fused_companions(self, enable):
        """True / False: companions inside the end passes / as separate kernels; None: by call size (default)."""
        self._check(self._L.ntt256_ctx_set_fused_companions(self._h, -1 if enable is None else int(bool(enable))))

    def fourstep_step1(self, shard, log_n: int, log_r: int, log_p: int, rank: int, stream=None):
        ptr = _device_ptr(shard, 1 << (log_n - log_p))
        self._check(self._L.ntt256_fourstep_step1(self._h, ptr, log_n, log_r, log_p, rank, _stream_handle(stream)))
        return shard

    def fourstep_step1_peer(self, shard, peer_ptrs: Sequence[int], log_n: int, log_r: int, log_p: int, rank: int,
                            stream=None):
        """step 1 whose last pass stores straight into the peers' receive buffers (device pointers
        valid in this process, one per rank)."""
        ptr = _device_ptr(shard, 1 << (log_n - log_p))
        arr = (ctypes.c_void_p * len(peer_ptrs))(*[int(v) for v in peer_ptrs])
        self._check(self._L.ntt256_fourstep_step1_peer(self._h, ptr, arr, log_n, log_r, log_p, rank,
                                                       _stream_handle(stream)))
        return shard

    def fourstep_step2(self, shard, log_n: int, log_r: int, log_p: int, rank: int, stream=None):
        ptr = _device_ptr(shard, 1 << (log_n - log_p))
        self._check(self._L.ntt256_fourstep_step2(self._h, ptr, log_n, log_r, log_p, rank, _stream_handle(stream)))
        return shard

    def fourstep_inverse_step1(self, shard, log_n: int, log_r: int, log_p: int, rank: int, stream=None, peer_ptrs=None):
        ptr = _device_ptr(shard, 1 << (log_n - log_p))
        if peer_ptrs is None:
            self._check(self._L.ntt256_fourstep_inverse_step1(self._h, ptr, log_n, log_r, log_p, rank, _stream_handle(stream)))
        else:
            arr = (ctypes.c_void_p * len(peer_ptrs))(*[int(v) for v in peer_ptrs])
            self._check(self._L.ntt256_fourstep_inverse_step1_peer(self._h, ptr, arr, log_n, log_r, log_p, rank,
                                                                   _stream_handle(stream)))
        return shard

    def fourstep_inverse_step2(self, shard, log_n: int, log_r: int, log_p: int, rank: int, stream=None):
        ptr = _device_ptr(shard, 1 << (log_n - log_p))
        self._check(self._L.ntt256_fourstep_inverse_step2(self._h, ptr, log_n, log_r, log_p, rank, _stream_handle(stream)))
        return shard

    # ---- host entry points (numpy in, numpy out, synchronous; H2D + transform + D2H)
    def forward_host(self, x: np.ndarray, log_n: int, batch: int = 1, out: Optional[np.ndarray] = None):
        return self._host(self._L.ntt256_forward_host, x, log_n, batch, out)

    def inverse_host(self, x: np.ndarray, log_n: int, batch: int = 1, out: Optional[np.ndarray] = None):
        return self._host(self._L.ntt256_inverse_host, x, log_n, batch, out)

    def _host(self, fn, x, log_n, batch, out):
        x = np.ascontiguousarray(x, dtype=np.uint32)
        if x.size != (batch << log_n) * 8:
            raise ValueError("host array must hold batch * 2^log_n elements of 8 u32 limbs")
        if out is None:
            out = np.empty_like(x)
        self._check(fn(self._h, x.ctypes.data, out.ctypes.data, log_n, batch))
        return out


_CTX_CACHE: dict = {}


def _cached_ctx(device: int, p: int, g: int, log_n: int) -> Ntt256:
    two_adicity = ((p - 1) & -(p - 1)).bit_length() - 1
    max_log_n = min(max(log_n, 12), two_adicity, 32)
    key = (device, p, g)
    ctx = _CTX_CACHE.get(key)
    if ctx is None or ctx.max_log_n < log_n:
        if ctx is not None:
            ctx.close()
            max_log_n = min(max(max_log_n, ctx.max_log_n), two_adicity)
        ctx = Ntt256(max_log_n=max_log_n, device=device, p=p, g=g)
        _CTX_CACHE[key] = ctx
    return ctx


def _from_limbs(v) -> int:
    if isinstance(v, (int, np.integer)):
        return int(v)
    return sum(int(l) << (32 * i) for i, l in enumerate(np.asarray(v).reshape(-1)))


def NTT_GZKP(data, length: int, reverse=None, reverse_len: int = 0, prime=None, omega=None, B: int = 0, G: int = 0,
             stream=None):
    """Drop-in for NTT_GZKP<8,256> (big-num.cu:260-353): in-place forward NTT of `length` points
    on device memory; `prime`, `omega` (= generator) are ints or 8-limb arrays.  `reverse`,
    `reverse_len`, `B`, `G` are accepted for signature compatibility and ignored.  Unlike the
    reference it does not synchronise or print; errors raise Ntt256Error."""
    if length <= 0 or length & (length - 1):
        raise ValueError("length must be a power of two (the reference asserts it, big-num.cu:385)")
    log_n = length.bit_length() - 1
    p, g = _from_limbs(prime), _from_limbs(omega)
    ctx = _cached_ctx(data.device.index or 0, p, g, log_n)
    return ctx.forward(data, log_n, 1, stream)


def NTT_GPU_Naive(data, length: int, reverse=None, reverse_len: int = 0, prime=None, omega=None, stream=None):
    """Drop-in for NTT_GPU_Naive<8,256> (big-num.cu:106-170): same contract and same results as
    NTT_GZKP (the reference checks both against the same CPU output, big-num.cu:439-467)."""
    return NTT_GZKP(data, length, reverse, reverse_len, prime, omega, 0, 0, stream)


# 384-bit fields (include/ntt384.h): name -> (p, generator)
FIELDS384 = {
    # BLS12-377 base field = BW6-761 scalar field: 377 bits, 2^46 | p - 1, generator 15 (arkworks)
    "bls12_377_fq": (0x01AE3A4617C510EAC63B05C06CA1493B1A22D9F300F5138F1EF3622FBA094800170B5D44300000008508C00000000001, 15),
}


class Ntt384:
    """ntt384_ctx: forward / inverse NTT over a prime field in 384-bit containers (12 LE u32 limbs,
    48 bytes per element, canonical residues), one GPU; the reference's `bits` template parameter
    (big-num.cu:57,67,172,260) at 384."""

    def __init__(self, field="bls12_377_fq", max_log_n: int = 20, device: int = 0, p=None, g=None):
        L = _load_lib()
        self._L = L
        self._h = ctypes.c_void_p()
        if p is None:
            p, g = FIELDS384[field]
        self.p, self.g = int(p), int(g)
        self.max_log_n = max_log_n
        pl = np.array([(self.p >> (32 * i)) & 0xFFFFFFFF for i in range(12)], dtype=np.uint32)
        gl = np.array([(self.g >> (32 * i)) & 0xFFFFFFFF for i in range(12)], dtype=np.uint32)
        rc = L.ntt384_ctx_create(ctypes.byref(self._h), device, _u32p(pl), _u32p(gl), max_log_n)
        if rc:
            raise Ntt256Error(rc)

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._L.ntt384_ctx_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int):
        if rc:
            raise Ntt256Error(rc, self._L.ntt384_last_cuda_error(self._h).decode())

    def set_tuning(self, tile_bits: int, coalesce_bits: int):
        self._check(self._L.ntt384_ctx_set_tuning(self._h, tile_bits, coalesce_bits))

    @staticmethod
    def _ptr(data, n_elements: int) -> int:
        if data.numel() * data.element_size() != n_elements * 48 or not data.is_cuda or not data.is_contiguous():
            raise ValueError(f"expected a contiguous CUDA tensor of {n_elements} elements of 48 bytes")
        return data.data_ptr()

    @property
    def launch_count(self) -> int:
        return int(self._L.ntt384_launch_count(self._h))

    def forward(self, data, log_n: int, batch: int = 1, stream=None):
        self._check(self._L.ntt384_forward_batch(self._h, self._ptr(data, batch << log_n), log_n, batch, _stream_handle(stream)))
        return data

    def inverse(self, data, log_n: int, batch: int = 1, stream=None):
        self._check(self._L.ntt384_inverse_batch(self._h, self._ptr(data, batch << log_n), log_n, batch, _stream_handle(stream)))
        return data

    # ---- host buffers (numpy uint32 [.., 12]; synchronous)
    def _host(self, fn, x, log_n, batch, out):
        x = np.ascontiguousarray(x, dtype=np.uint32)
        if x.size != (batch << log_n) * 12:
            raise ValueError("host array must hold batch * 2^log_n elements of 12 u32 limbs")
        if out is None:
            out = np.empty_like(x)
        self._check(fn(self._h, x.ctypes.data, out.ctypes.data, log_n, batch))
        return out

    def forward_host(self, x, log_n: int, batch: int = 1, out=None):
        return self._host(self._L.ntt384_forward_host, x, log_n, batch, out)

    def inverse_host(self, x, log_n: int, batch: int = 1, out=None):
        return self._host(self._L.ntt384_inverse_host, x, log_n, batch, out)

    # ---- companions fused into the end passes
    def _shift(self, shift: int):
        return _u32p(np.array([(int(shift) >> (32 * i)) & 0xFFFFFFFF for i in range(12)], dtype=np.uint32))

    def coset_forward(self, data, log_n: int, shift: int, batch: int = 1, stream=None):
        self._check(self._L.ntt384_coset_forward(self._h, self._ptr(data, batch << log_n), log_n, batch, self._shift(shift),
                                                 _stream_handle(stream)))
        return data

    def coset_inverse(self, data, log_n: int, shift: int, batch: int = 1, stream=None):
        self._check(self._L.ntt384_coset_inverse(self._h, self._ptr(data, batch << log_n), log_n, batch, self._shift(shift),
                                                 _stream_handle(stream)))
        return data

    def forward_mul(self, data, mul_mont, log_n: int, batch: int = 1, stream=None):
        n = batch << log_n
        self._check(self._L.ntt384_forward_mul(self._h, self._ptr(data, n), self._ptr(mul_mont, n), log_n, batch, _stream_handle(stream)))
        return data

    def inverse_mul(self, data, mul_mont, log_n: int, batch: int = 1, stream=None):
        n = batch << log_n
        self._check(self._L.ntt384_inverse_mul(self._h, self._ptr(data, n), self._ptr(mul_mont, n), log_n, batch, _stream_handle(stream)))
        return data

    # ---- multi-GPU four-step pieces (same names and meaning as Ntt256's: ntt_b200.dist.FourStep drives either)
    def set_peer_sync(self, flag_ptrs=None, rank: int = 0, timeout_s: float = 0.0):
        if flag_ptrs is None:
            self._check(self._L.ntt384_ctx_set_peer_sync(self._h, None, 0, 0, 0.0))
        else:
            arr = (ctypes.c_void_p * len(flag_ptrs))(*[int(v) for v in flag_ptrs])
            self._check(self._L.ntt384_ctx_set_peer_sync(self._h, arr, len(flag_ptrs), rank, float(timeout_s)))

    def fourstep_step1(self, shard, log_n: int, log_r: int, log_p: int, rank: int, stream=None):
        ptr = self._ptr(shard, 1 << (log_n - log_p))
        self._check(self._L.ntt384_fourstep_step1(self._h, ptr, log_n, log_r, log_p, rank, _stream_handle(stream)))
        return shard

    def fourstep_step1_peer(self, shard, peer_ptrs, log_n: int, log_r: int, log_p: int, rank: int, stream=None):
        ptr = self._ptr(shard, 1 << (log_n - log_p))
        arr = (ctypes.c_void_p * len(peer_ptrs))(*[int(v) for v in peer_ptrs])
        self._check(self._L.ntt384_fourstep_step1_peer(self._h, ptr, arr, log_n, log_r, log_p, rank, _stream_handle(stream)))
        return shard

    def fourstep_step2(self, shard, log_n: int, log_r: int, log_p: int, rank: int, stream=None):
        ptr = self._ptr(shard, 1 << (log_n - log_p))
        self._check(self._L.ntt384_fourstep_step2(self._h, ptr, log_n, log_r, log_p, rank, _stream_handle(stream)))
        return shard

    def fourstep_inverse_step1(self, shard, log_n: int, log_r: int, log_p: int, rank: int, stream=None, peer_ptrs=None):
        ptr = self._ptr(shard, 1 << (log_n - log_p))
        if peer_ptrs is None:
            self._check(self._L.ntt384_fourstep_inverse_step1(self._h, ptr, log_n, log_r, log_p, rank, _stream_handle(stream)))
        else:
            arr = (ctypes.c_void_p * len(peer_ptrs))(*[int(v) for v in peer_ptrs])
            self._check(self._L.ntt384_fourstep_inverse_step1_peer(self._h, ptr, arr, log_n, log_r, log_p, rank,
                                                                   _stream_handle(stream)))
        return shard

    def fourstep_inverse_step2(self, shard, log_n: int, log_r: int, log_p: int, rank: int, stream=None):
        ptr = self._ptr(shard, 1 << (log_n - log_p))
        self._check(self._L.ntt384_fourstep_inverse_step2(self._h, ptr, log_n, log_r, log_p, rank, _stream_handle(stream)))
        return shard

    def vec_mul(self, a, b, out=None, mont_form: bool = False, stream=None):
        n = a.numel() * a.element_size() // 48
        out = a if out is None else out
        self._check(self._L.ntt384_vec_mul(self._h, self._ptr(out, n), self._ptr(a, n), self._ptr(b, n), n, int(mont_form),
                                           _stream_handle(stream)))
        return out


class Ntt256MultiGpu:
    """ntt256_mgpu: one host thread driving several peer-connected GPUs (no NCCL, no extra
    processes).  Shards are torch CUDA tensors, one per device, in the four-step layouts of
    ntt_b200.dist (scatter_input / gather_output)."""

    def __init__(self, ngpus: int, field: str = "bn254_fr", max_log_n: int = 24, devices: Optional[Sequence[int]] = None):
        from .lib import FIELD_BLS12_381_FR, FIELD_BN254_FR

        self._L = _load_lib()
        self._h = ctypes.c_void_p()
        fid = {"bn254_fr": FIELD_BN254_FR, "bls12_381_fr": FIELD_BLS12_381_FR}[field]
        devs = (ctypes.c_int * ngpus)(*devices) if devices is not None else None
        rc = self._L.ntt256_mgpu_create(ctypes.byref(self._h), ngpus, devs, fid, max_log_n)
        if rc:
            raise Ntt256Error(rc)
        self.ngpus = ngpus

    def _ptrs(self, shards, n_elements):
        if len(shards) != self.ngpus:
            raise ValueError("one shard per GPU")
        return (ctypes.c_void_p * self.ngpus)(*[_device_ptr(s, n_elements) for s in shards])

    def forward(self, shards, log_n: int):
        rc = self._L.ntt256_mgpu_forward(self._h, self._ptrs(shards, (1 << log_n) // self.ngpus), log_n)
        if rc:
            raise Ntt256Error(rc)

    def inverse(self, shards, log_n: int):
        rc = self._L.ntt256_mgpu_inverse(self._h, self._ptrs(shards, (1 << log_n) // self.ngpus), log_n)
        if rc:
            raise Ntt256Error(rc)

    def forward_batch(self, shards, log_n: int, batch_per_gpu: int):
        rc = self._L.ntt256_mgpu_forward_batch(self._h, self._ptrs(shards, batch_per_gpu << log_n), log_n, batch_per_gpu)
        if rc:
            raise Ntt256Error(rc)

    def synchronize(self):
        rc = self._L.ntt256_mgpu_synchronize(self._h)
        if rc:
            raise Ntt256Error(rc)

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._L.ntt256_mgpu_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
