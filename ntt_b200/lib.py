"""ctypes binding of libntt256.so (include/ntt256.h).  The CUDA library is the product: there
is no CPU or PyTorch fallback -- a missing or unloadable library raises at import of this
module's `lib()`."""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libntt256.so")
_LIB = None

OK, ERR_ARG, ERR_FIELD, ERR_CUDA, ERR_NOMEM, ERR_PLAN = range(6)
FIELD_BN254_FR, FIELD_BLS12_381_FR = 0, 1

# every symbol include/ntt256.h declares
EXPORTS = [
    "ntt256_ctx_create", "ntt256_ctx_create_named", "ntt256_ctx_destroy", "ntt256_ctx_set_tuning",
    "ntt256_ctx_set_stream_twiddle_limit", "ntt256_ctx_set_stream_twiddle_budget", "ntt256_ctx_table_bytes", "ntt256_ctx_set_small_transform_limit", "ntt256_ctx_set_constant_field_kernel", "ntt256_ctx_set_dependent_launch", "ntt256_ctx_set_elements_per_thread",
    "ntt256_ctx_set_twiddle_mode", "ntt256_ctx_get_twiddle_mode",
    "ntt256_forward", "ntt256_inverse", "ntt256_forward_batch", "ntt256_inverse_batch",
    "ntt256_forward_host", "ntt256_inverse_host", "ntt256_fourstep_step1", "ntt256_fourstep_step2", "ntt256_fourstep_step1_peer", "ntt256_enable_peer_access",
    "ntt256_fourstep_inverse_step1", "ntt256_fourstep_inverse_step1_peer", "ntt256_fourstep_inverse_step2",
    "ntt256_ipc_alloc", "ntt256_ipc_open", "ntt256_ipc_close", "ntt256_ipc_free", "ntt256_peer_barrier",
    "ntt256_set_peer_barrier_timeout", "ntt256_peer_sync_status", "ntt256_ctx_set_peer_sync",
    "ntt256_mgpu_create", "ntt256_mgpu_forward", "ntt256_mgpu_inverse", "ntt256_mgpu_forward_batch",
    "ntt256_mgpu_synchronize", "ntt256_mgpu_destroy",
    "ntt256_vec_mul", "ntt256_to_montgomery", "ntt256_from_montgomery", "ntt256_coset_forward",
    "ntt256_coset_inverse", "ntt256_forward_mul", "ntt256_inverse_mul", "ntt256_ctx_set_fused_companions", "ntt256_ctx_set_dense_pass4",
    "ntt256_plan_describe", "ntt256_plan_describe_ex", "ntt256_plan_describe_fourstep", "ntt256_launch_count", "ntt256_strerror", "ntt256_last_cuda_error",
    "ntt256_version",
]


# every symbol include/ntt384.h declares (the 384-bit companion API, same library)
EXPORTS384 = [
    "ntt384_ctx_create", "ntt384_ctx_destroy", "ntt384_ctx_set_tuning", "ntt384_forward", "ntt384_inverse", "ntt384_forward_batch",
    "ntt384_inverse_batch", "ntt384_forward_host", "ntt384_inverse_host", "ntt384_coset_forward", "ntt384_coset_inverse",
    "ntt384_forward_mul", "ntt384_inverse_mul", "ntt384_fourstep_step1", "ntt384_fourstep_step1_peer", "ntt384_fourstep_step2",
    "ntt384_fourstep_inverse_step1", "ntt384_fourstep_inverse_step1_peer", "ntt384_fourstep_inverse_step2",
    "ntt384_ctx_set_peer_sync", "ntt384_vec_mul", "ntt384_launch_count", "ntt384_last_cuda_error", "ntt384_plan_describe", "ntt384_plan_describe_fourstep",
]


# every symbol include/ntt64.h declares (the 64-bit companion API, same library)
EXPORTS64 = [
    "ntt64_ctx_create", "ntt64_ctx_destroy", "ntt64_ctx_set_tuning", "ntt64_ctx_set_shift_twiddles", "ntt64_forward", "ntt64_inverse", "ntt64_forward_batch",
    "ntt64_inverse_batch", "ntt64_forward_host", "ntt64_inverse_host", "ntt64_coset_forward", "ntt64_coset_inverse", "ntt64_ctx_set_coset_table_limit",
    "ntt64_fourstep_step1", "ntt64_fourstep_step1_peer", "ntt64_fourstep_step2", "ntt64_fourstep_inverse_step1",
    "ntt64_fourstep_inverse_step1_peer", "ntt64_fourstep_inverse_step2",
    "ntt64_vec_mul", "ntt64_launch_count", "ntt64_last_cuda_error", "ntt64_plan_describe", "ntt64_plan_describe_fourstep",
]


class Ntt256Error(RuntimeError):
    def __init__(self, status: int, detail: str = ""):
        self.status = status
        msg = lib().ntt256_strerror(status).decode()
        super().__init__(f"ntt256 status {status}: {msg}{(' -- ' + detail) if detail else ''}")


def lib() -> ctypes.CDLL:
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback)"
            )
        L = ctypes.CDLL(LIB_PATH)
        vp, ci, u32p = ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_uint32)
        L.ntt256_ctx_create.argtypes = [ctypes.POINTER(vp), ci, u32p, u32p, ci]
        L.ntt256_ctx_create_named.argtypes = [ctypes.POINTER(vp), ci, ci, ci]
        L.ntt256_ctx_destroy.argtypes = [vp]
        L.ntt256_ctx_destroy.restype = None
        L.ntt256_ctx_set_tuning.argtypes = [vp, ci, ci]
        L.ntt256_ctx_set_stream_twiddle_limit.argtypes = [vp, ctypes.c_size_t]
        L.ntt256_ctx_set_stream_twiddle_budget.argtypes = [vp, ctypes.c_size_t]
        L.ntt256_ctx_table_bytes.argtypes = [vp, ctypes.POINTER(ctypes.c_size_t)]
        L.ntt256_ctx_table_bytes.restype = ctypes.c_size_t
        L.ntt256_ctx_set_small_transform_limit.argtypes = [vp, ctypes.c_longlong]
        L.ntt256_ctx_set_constant_field_kernel.argtypes = [vp, ci]
        L.ntt256_ctx_set_dependent_launch.argtypes = [vp, ci]
        L.ntt256_ctx_set_elements_per_thread.argtypes = [vp, ci]
        L.ntt256_ctx_set_twiddle_mode.argtypes = [vp, ci]
        L.ntt256_ctx_get_twiddle_mode.argtypes = [vp]
        for name in ("ntt256_forward", "ntt256_inverse"):
            getattr(L, name).argtypes = [vp, vp, ci, vp]
        for name in ("ntt256_forward_batch", "ntt256_inverse_batch"):
            getattr(L, name).argtypes = [vp, vp, ci, ci, vp]
        for name in ("ntt256_forward_host", "ntt256_inverse_host"):
            getattr(L, name).argtypes = [vp, vp, vp, ci, ci]
        for name in ("ntt256_fourstep_step1", "ntt256_fourstep_step2", "ntt256_fourstep_inverse_step1",
                     "ntt256_fourstep_inverse_step2"):
            getattr(L, name).argtypes = [vp, vp, ci, ci, ci, ci, vp]
        L.ntt256_fourstep_inverse_step1_peer.argtypes = [vp, vp, ctypes.POINTER(vp), ci, ci, ci, ci, vp]
        L.ntt256_fourstep_step1_peer.argtypes = [vp, vp, ctypes.POINTER(vp), ci, ci, ci, ci, vp]
        L.ntt256_vec_mul.argtypes = [vp, vp, vp, vp, ctypes.c_ulonglong, ci, vp]
        for name in ("ntt256_to_montgomery", "ntt256_from_montgomery"):
            getattr(L, name).argtypes = [vp, vp, ctypes.c_ulonglong, vp]
        for name in ("ntt256_coset_forward", "ntt256_coset_inverse"):
            getattr(L, name).argtypes = [vp, vp, ci, ci, u32p, vp]
        for name in ("ntt256_forward_mul", "ntt256_inverse_mul"):
            getattr(L, name).argtypes = [vp, vp, vp, ci, ci, vp]
        L.ntt256_ctx_set_fused_companions.argtypes = [vp, ci]
        L.ntt256_ctx_set_dense_pass4.argtypes = [vp, ci]
        L.ntt256_enable_peer_access.argtypes = [ci, ci]
        L.ntt256_mgpu_create.argtypes = [ctypes.POINTER(vp), ci, ctypes.POINTER(ci), ci, ci]
        L.ntt256_mgpu_forward.argtypes = [vp, ctypes.POINTER(vp), ci]
        L.ntt256_mgpu_inverse.argtypes = [vp, ctypes.POINTER(vp), ci]
        L.ntt256_mgpu_forward_batch.argtypes = [vp, ctypes.POINTER(vp), ci, ci]
        L.ntt256_mgpu_synchronize.argtypes = [vp]
        L.ntt256_mgpu_destroy.argtypes = [vp]
        L.ntt256_mgpu_destroy.restype = None
        L.ntt256_ipc_alloc.argtypes = [ci, ctypes.c_size_t, ctypes.POINTER(vp), ctypes.c_char_p]
        L.ntt256_ipc_open.argtypes = [ci, ctypes.c_char_p, ctypes.POINTER(vp)]
        L.ntt256_ipc_close.argtypes = [ci, vp]
        L.ntt256_ipc_free.argtypes = [ci, vp]
        L.ntt256_peer_barrier.argtypes = [ci, ctypes.POINTER(vp), ci, ci, ctypes.c_ulonglong, vp]
        L.ntt256_set_peer_barrier_timeout.argtypes = [ctypes.c_double]
        L.ntt256_peer_sync_status.argtypes = [ci, vp, ctypes.POINTER(ctypes.c_ulonglong)]
        L.ntt256_ctx_set_peer_sync.argtypes = [vp, ctypes.POINTER(vp), ci, ci, ctypes.c_double]
        L.ntt256_plan_describe.argtypes = [ci, ci, ci, ci, ctypes.c_char_p, ctypes.c_size_t]
        L.ntt256_plan_describe.restype = ctypes.c_size_t
        L.ntt256_plan_describe_ex.argtypes = [ci, ci, ci, ci, ci, ctypes.c_char_p, ctypes.c_size_t]
        L.ntt256_plan_describe_ex.restype = ctypes.c_size_t
        L.ntt256_plan_describe_fourstep.argtypes = [ci] * 8 + [ctypes.c_char_p, ctypes.c_size_t]
        L.ntt256_plan_describe_fourstep.restype = ctypes.c_size_t
        L.ntt256_launch_count.argtypes = [vp]
        L.ntt256_launch_count.restype = ctypes.c_ulonglong
        L.ntt256_strerror.argtypes = [ci]
        L.ntt256_strerror.restype = ctypes.c_char_p
        L.ntt256_last_cuda_error.argtypes = [vp]
        L.ntt256_last_cuda_error.restype = ctypes.c_char_p
        L.ntt256_version.restype = ctypes.c_char_p
        L.ntt384_ctx_create.argtypes = [ctypes.POINTER(vp), ci, u32p, u32p, ci]
        L.ntt384_ctx_destroy.argtypes = [vp]
        L.ntt384_ctx_destroy.restype = None
        L.ntt384_ctx_set_tuning.argtypes = [vp, ci, ci]
        for name in ("ntt384_forward", "ntt384_inverse"):
            getattr(L, name).argtypes = [vp, vp, ci, vp]
        for name in ("ntt384_forward_batch", "ntt384_inverse_batch"):
            getattr(L, name).argtypes = [vp, vp, ci, ci, vp]
        L.ntt384_vec_mul.argtypes = [vp, vp, vp, vp, ctypes.c_ulonglong, ci, vp]
        for name in ("ntt384_forward_host", "ntt384_inverse_host"):
            getattr(L, name).argtypes = [vp, vp, vp, ci, ci]
        for name in ("ntt384_coset_forward", "ntt384_coset_inverse"):
            getattr(L, name).argtypes = [vp, vp, ci, ci, u32p, vp]
        for name in ("ntt384_forward_mul", "ntt384_inverse_mul"):
            getattr(L, name).argtypes = [vp, vp, vp, ci, ci, vp]
        for name in ("ntt384_fourstep_step1", "ntt384_fourstep_step2", "ntt384_fourstep_inverse_step1",
                     "ntt384_fourstep_inverse_step2"):
            getattr(L, name).argtypes = [vp, vp, ci, ci, ci, ci, vp]
        for name in ("ntt384_fourstep_step1_peer", "ntt384_fourstep_inverse_step1_peer"):
            getattr(L, name).argtypes = [vp, vp, ctypes.POINTER(vp), ci, ci, ci, ci, vp]
        L.ntt384_ctx_set_peer_sync.argtypes = [vp, ctypes.POINTER(vp), ci, ci, ctypes.c_double]
        L.ntt384_launch_count.argtypes = [vp]
        L.ntt384_launch_count.restype = ctypes.c_ulonglong
        L.ntt384_last_cuda_error.argtypes = [vp]
        L.ntt384_last_cuda_error.restype = ctypes.c_char_p
        L.ntt384_plan_describe.argtypes = [ci, ci, ctypes.c_char_p, ctypes.c_size_t]
        L.ntt384_plan_describe.restype = ctypes.c_size_t
        L.ntt384_plan_describe_fourstep.argtypes = [ci, ci, ci, ci, ci, ci, ctypes.c_char_p, ctypes.c_size_t]
        L.ntt384_plan_describe_fourstep.restype = ctypes.c_size_t
        u64 = ctypes.c_uint64
        L.ntt64_ctx_create.argtypes = [ctypes.POINTER(vp), ci, u64, u64, ci]
        L.ntt64_ctx_destroy.argtypes = [vp]
        L.ntt64_ctx_destroy.restype = None
        L.ntt64_ctx_set_tuning.argtypes = [vp, ci, ci]
        L.ntt64_ctx_set_shift_twiddles.argtypes = [vp, ci]
        L.ntt64_ctx_set_coset_table_limit.argtypes = [vp, ctypes.c_size_t]
        for name in ("ntt64_fourstep_step1", "ntt64_fourstep_step2", "ntt64_fourstep_inverse_step1", "ntt64_fourstep_inverse_step2"):
            getattr(L, name).argtypes = [vp, vp, ci, ci, ci, ci, vp]
        for name in ("ntt64_fourstep_step1_peer", "ntt64_fourstep_inverse_step1_peer"):
            getattr(L, name).argtypes = [vp, vp, ctypes.POINTER(vp), ci, ci, ci, ci, vp]
        for name in ("ntt64_forward", "ntt64_inverse"):
            getattr(L, name).argtypes = [vp, vp, ci, vp]
        for name in ("ntt64_forward_batch", "ntt64_inverse_batch"):
            getattr(L, name).argtypes = [vp, vp, ci, ci, vp]
        for name in ("ntt64_forward_host", "ntt64_inverse_host"):
            getattr(L, name).argtypes = [vp, vp, vp, ci, ci]
        for name in ("ntt64_coset_forward", "ntt64_coset_inverse"):
            getattr(L, name).argtypes = [vp, vp, ci, ci, u64, vp]
        L.ntt64_vec_mul.argtypes = [vp, vp, vp, vp, ctypes.c_ulonglong, vp]
        L.ntt64_launch_count.argtypes = [vp]
        L.ntt64_launch_count.restype = ctypes.c_ulonglong
        L.ntt64_last_cuda_error.argtypes = [vp]
        L.ntt64_last_cuda_error.restype = ctypes.c_char_p
        L.ntt64_plan_describe.argtypes = [ci, ci, ci, ctypes.c_char_p, ctypes.c_size_t]
        L.ntt64_plan_describe.restype = ctypes.c_size_t
        L.ntt64_plan_describe_fourstep.argtypes = [ci, ci, ci, ci, ci, ci, ctypes.c_char_p, ctypes.c_size_t]
        L.ntt64_plan_describe_fourstep.restype = ctypes.c_size_t
        _LIB = L
    return _LIB


def plan_describe(log_n: int, max_log_n: int, tile_bits: int = 11, coalesce_bits: int = 2, group_bits: int = 3) -> dict:
    """Host-only: the pass plan for 2^log_n points (JSON -> dict); group_bits 3 = eight elements per
    thread, 2 = four."""
    import json

    L = lib()
    need = L.ntt256_plan_describe_ex(log_n, max_log_n, tile_bits, coalesce_bits, group_bits, None, 0)
    buf = ctypes.create_string_buffer(need)
    L.ntt256_plan_describe_ex(log_n, max_log_n, tile_bits, coalesce_bits, group_bits, buf, need)
    return json.loads(buf.value.decode())


def plan_describe_fourstep(step: int, log_n: int, log_r: int, log_p: int, rank: int, max_log_n: int,
                           tile_bits: int = 11, coalesce_bits: int = 2) -> dict:
    """Host-only: the plan of local step 1 or 2 of the four-step decomposition on `rank`."""
    import json

    L = lib()
    a = (step, log_n, log_r, log_p, rank, max_log_n, tile_bits, coalesce_bits)
    need = L.ntt256_plan_describe_fourstep(*a, None, 0)
    buf = ctypes.create_string_buffer(need)
    L.ntt256_plan_describe_fourstep(*a, buf, need)
    return json.loads(buf.value.decode())


def plan_describe_384_fourstep(step: int, log_n: int, log_r: int, log_p: int, rank: int, max_log_n: int) -> dict:
    """Host-only: the plan of one local step (1 / 2 forward, 3 / 4 inverse) of the 384-bit four-step decomposition."""
    import json

    L = lib()
    a = (step, log_n, log_r, log_p, rank, max_log_n)
    need = L.ntt384_plan_describe_fourstep(*a, None, 0)
    buf = ctypes.create_string_buffer(need)
    L.ntt384_plan_describe_fourstep(*a, buf, need)
    return json.loads(buf.value.decode())


def plan_describe_384(log_n: int, max_log_n: int) -> dict:
    """Host-only: the pass plan ntt384_forward would run for 2^log_n points (JSON -> dict)."""
    import json

    L = lib()
    need = L.ntt384_plan_describe(log_n, max_log_n, None, 0)
    buf = ctypes.create_string_buffer(need)
    L.ntt384_plan_describe(log_n, max_log_n, buf, need)
    return json.loads(buf.value.decode())
