"""Host-side mirror of include/ntt64.h: forward / inverse NTT over prime fields in 64-bit containers (one u64
per element) -- the width the reference's own CPU NTT runs in (/root/reference/src/big-num.cu:37-55, p =
469762049, g = 3) and of the Goldilocks prime 2^64 - 2^32 + 1.  Same conventions as ntt_b200.api.Ntt256:
device tensors transformed in place, natural order in and out, X[k] = sum_j x[j] w_n^(jk), w_n = g^((p-1)/n).
There is no CPU path: every call goes through the C ABI of libntt256.so."""
from __future__ import annotations

import ctypes
import json

import numpy as np

from .api import _stream_handle
from .lib import Ntt256Error, lib as _load_lib

FIELDS64 = {
    "goldilocks": (0xFFFFFFFF00000001, 7),
    "ref29": (469762049, 3),  # the only field the reference runs (big-num.cu:9-10)
}


def plan_describe64(log_n: int, batch: int, max_log_n: int) -> dict:
    """Host-only: the pass plan a call of `batch` transforms of 2^log_n points gets with the default tuning."""
    L = _load_lib()
    need = L.ntt64_plan_describe(log_n, batch, max_log_n, None, 0)
    buf = ctypes.create_string_buffer(need)
    L.ntt64_plan_describe(log_n, batch, max_log_n, buf, need)
    return json.loads(buf.value.decode())


def plan_describe64_fourstep(step: int, log_n: int, log_r: int, log_p: int, rank: int, max_log_n: int) -> dict:
    """Host-only: the plan of one local step (1 / 2 forward, 3 / 4 inverse) of the four-step decomposition."""
    L = _load_lib()
    a = (step, log_n, log_r, log_p, rank, max_log_n)
    need = L.ntt64_plan_describe_fourstep(*a, None, 0)
    buf = ctypes.create_string_buffer(need)
    L.ntt64_plan_describe_fourstep(*a, buf, need)
    return json.loads(buf.value.decode())


class Ntt64:
    """ntt64_ctx: one GPU, one field (p, generator g), transforms of up to 2^max_log_n points."""

    def __init__(self, field="goldilocks", max_log_n: int = 24, device: int = 0, p=None, g=None):
        L = _load_lib()
        self._L = L
        self._h = ctypes.c_void_p()
        if p is None:
            p, g = FIELDS64[field]
        self.p, self.g = int(p), int(g)
        self.max_log_n = max_log_n
        rc = L.ntt64_ctx_create(ctypes.byref(self._h), device, self.p, self.g, max_log_n)
        if rc:
            raise Ntt256Error(rc)

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._L.ntt64_ctx_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int):
        if rc:
            raise Ntt256Error(rc, self._L.ntt64_last_cuda_error(self._h).decode())

    def set_tuning(self, tile_bits: int, coalesce_bits: int):
        self._check(self._L.ntt64_ctx_set_tuning(self._h, tile_bits, coalesce_bits))

    def set_shift_twiddles(self, enable: bool):
        """Goldilocks: constant twiddles of the register rounds by shifts (default) or on the multiplier."""
        self._check(self._L.ntt64_ctx_set_shift_twiddles(self._h, int(bool(enable))))

    def set_coset_table_limit(self, max_bytes: int):
        """Largest multiplied-out shift^j table; longer coset transforms use two factored tables + a separate pass."""
        self._check(self._L.ntt64_ctx_set_coset_table_limit(self._h, int(max_bytes)))

    @staticmethod
    def _ptr(data, n_elements: int) -> int:
        if data.numel() * data.element_size() != n_elements * 8 or not data.is_cuda or not data.is_contiguous():
            raise ValueError(f"expected a contiguous CUDA tensor of {n_elements} elements of 8 bytes")
        return data.data_ptr()

    @property
    def launch_count(self) -> int:
        return int(self._L.ntt64_launch_count(self._h))

    # ---- device entry points (in place, asynchronous on `stream`)
    def forward(self, data, log_n: int, batch: int = 1, stream=None):
        self._check(self._L.ntt64_forward_batch(self._h, self._ptr(data, batch << log_n), log_n, batch, _stream_handle(stream)))
        return data

    def inverse(self, data, log_n: int, batch: int = 1, stream=None):
        self._check(self._L.ntt64_inverse_batch(self._h, self._ptr(data, batch << log_n), log_n, batch, _stream_handle(stream)))
        return data

    def coset_forward(self, data, log_n: int, shift: int, batch: int = 1, stream=None):
        self._check(self._L.ntt64_coset_forward(self._h, self._ptr(data, batch << log_n), log_n, batch, int(shift), _stream_handle(stream)))
        return data

    def coset_inverse(self, data, log_n: int, shift: int, batch: int = 1, stream=None):
        self._check(self._L.ntt64_coset_inverse(self._h, self._ptr(data, batch << log_n), log_n, batch, int(shift), _stream_handle(stream)))
        return data

    def vec_mul(self, a, b, out=None, stream=None):
        n = a.numel() * a.element_size() // 8
        out = a if out is None else out
        self._check(self._L.ntt64_vec_mul(self._h, self._ptr(out, n), self._ptr(a, n), self._ptr(b, n), n, _stream_handle(stream)))
        return out

    # ---- multi-GPU four-step pieces (same names and meaning as Ntt256's: ntt_b200.dist.FourStep drives either; use
    # its default flag barrier -- the in-kernel rank synchronisation exists for the wide kernels only)
    def set_peer_sync(self, flag_ptrs=None, rank: int = 0, timeout_s: float = 0.0):
        if flag_ptrs is not None:
            raise NotImplementedError("64-bit kernels: use FourStep(barrier='flags') or 'nccl'")

    def _fs(self, fn, shard, log_n, log_r, log_p, rank, stream):
        self._check(fn(self._h, self._ptr(shard, 1 << (log_n - log_p)), log_n, log_r, log_p, rank, _stream_handle(stream)))
        return shard

    def _fs_peer(self, fn, shard, peer_ptrs, log_n, log_r, log_p, rank, stream):
        arr = (ctypes.c_void_p * len(peer_ptrs))(*[int(v) for v in peer_ptrs])
        self._check(fn(self._h, self._ptr(shard, 1 << (log_n - log_p)), arr, log_n, log_r, log_p, rank, _stream_handle(stream)))
        return shard

    def fourstep_step1(self, shard, log_n: int, log_r: int, log_p: int, rank: int, stream=None):
        return self._fs(self._L.ntt64_fourstep_step1, shard, log_n, log_r, log_p, rank, stream)

    def fourstep_step1_peer(self, shard, peer_ptrs, log_n: int, log_r: int, log_p: int, rank: int, stream=None):
        return self._fs_peer(self._L.ntt64_fourstep_step1_peer, shard, peer_ptrs, log_n, log_r, log_p, rank, stream)

    def fourstep_step2(self, shard, log_n: int, log_r: int, log_p: int, rank: int, stream=None):
        return self._fs(self._L.ntt64_fourstep_step2, shard, log_n, log_r, log_p, rank, stream)

    def fourstep_inverse_step1(self, shard, log_n: int, log_r: int, log_p: int, rank: int, stream=None, peer_ptrs=None):
        if peer_ptrs is None:
            return self._fs(self._L.ntt64_fourstep_inverse_step1, shard, log_n, log_r, log_p, rank, stream)
        return self._fs_peer(self._L.ntt64_fourstep_inverse_step1_peer, shard, peer_ptrs, log_n, log_r, log_p, rank, stream)

    def fourstep_inverse_step2(self, shard, log_n: int, log_r: int, log_p: int, rank: int, stream=None):
        return self._fs(self._L.ntt64_fourstep_inverse_step2, shard, log_n, log_r, log_p, rank, stream)

    # ---- host buffers (numpy uint64 arrays; synchronous)
    def _host(self, fn, x: np.ndarray, log_n: int, batch: int, out):
        x = np.ascontiguousarray(x, dtype=np.uint64)
        if x.size != batch << log_n:
            raise ValueError("size mismatch")
        if out is None:
            out = np.empty_like(x)
        if out.dtype != np.uint64 or out.size != x.size or not out.flags.c_contiguous:
            raise ValueError("out must be a contiguous uint64 array of the same size")
        self._check(fn(self._h, x.ctypes.data, out.ctypes.data, log_n, batch))
        return out

    def forward_host(self, x: np.ndarray, log_n: int, batch: int = 1, out=None) -> np.ndarray:
        return self._host(self._L.ntt64_forward_host, x, log_n, batch, out)

    def inverse_host(self, x: np.ndarray, log_n: int, batch: int = 1, out=None) -> np.ndarray:
        return self._host(self._L.ntt64_inverse_host, x, log_n, batch, out)
