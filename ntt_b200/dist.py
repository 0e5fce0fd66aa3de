"""Multi-GPU four-step NTT: one process per GPU, torch.distributed (NCCL over NVLink/NVSwitch)
for the single exchange step (SURVEY.md par. 8e; the reference has no multi-GPU path at all).

    n = R*C.  Rank r owns the column block j2 in [r*C/P, (r+1)*C/P) of the R x C row-major view
    of x (packed R x C/P).  step 1: R-point column NTTs with the twiddle w_n^(j2*k1) fused into
    the last pass; one all-to-all (rank r sends the contiguous row block k1 in block s to rank s);
    step 2: C-point row NTTs.  Rank r ends with X[k1 + R*k2], k1 in block r, stored as P blocks
    of (R/P) x (C/P): index = ((k2 mod P)*(R/P) + k1_local)*(C/P) + (k2 div P).

`scatter_input` / `gather_output` convert between a contiguous natural-order host array and
those shards.  Batches of independent transforms shard trivially (`shard_batch`) with no
communication."""
from __future__ import annotations

import numpy as np


def default_log_r(log_n: int) -> int:
    return (log_n + 1) // 2


def _log2(v: int) -> int:
    l = v.bit_length() - 1
    if v != 1 << l:
        raise ValueError("world size must be a power of two")
    return l


def scatter_input(x: np.ndarray, log_n: int, world: int, log_r: int | None = None, limbs: int = 8) -> list[np.ndarray]:
    """Natural-order x [n, limbs] -> per-rank input shards (column blocks of the R x C view); limbs = 8 for 256-bit
    containers, 12 for 384-bit ones."""
    log_r = default_log_r(log_n) if log_r is None else log_r
    R, C = 1 << log_r, 1 << (log_n - log_r)
    m = x.reshape(R, C, limbs)
    w = C // world
    return [np.ascontiguousarray(m[:, r * w:(r + 1) * w]).reshape(-1, limbs) for r in range(world)]


def gather_output(shards: list[np.ndarray], log_n: int, world: int, log_r: int | None = None, limbs: int = 8) -> np.ndarray:
    """Per-rank output shards (element [k2 mod P][k1_local][k2 div P] = X[k1 + R*k2]) -> natural-order X."""
    log_r = default_log_r(log_n) if log_r is None else log_r
    R, C = 1 << log_r, 1 << (log_n - log_r)
    P = world
    # [rank][k2_low][k1_local][k2_high] -> X index k2*R + k1 with k2 = k2_low + P*k2_high, k1 = rank*(R/P) + k1_local
    m = np.stack([s.reshape(P, R // P, C // P, limbs) for s in shards], axis=0)
    m = m.transpose(3, 1, 0, 2, 4)  # [k2_high][k2_low][rank][k1_local]
    return np.ascontiguousarray(m).reshape(-1, limbs)


def scatter_output_layout(X: np.ndarray, log_n: int, world: int, log_r: int | None = None, limbs: int = 8) -> list[np.ndarray]:
    """Natural-order X -> shards in the forward's OUTPUT layout (inverse of gather_output): the input
    of FourStep.inverse()."""
    log_r = default_log_r(log_n) if log_r is None else log_r
    R, C = 1 << log_r, 1 << (log_n - log_r)
    P = world
    m = X.reshape(C // P, P, P, R // P, limbs)       # [k2_high][k2_low][rank][k1_local]
    m = m.transpose(2, 1, 3, 0, 4)                   # [rank][k2_low][k1_local][k2_high]
    return [np.ascontiguousarray(m[r]).reshape(-1, limbs) for r in range(world)]


def gather_input_layout(shards: list[np.ndarray], log_n: int, world: int, log_r: int | None = None, limbs: int = 8) -> np.ndarray:
    """Shards in the forward's INPUT layout (column blocks) -> natural order (inverse of scatter_input)."""
    log_r = default_log_r(log_n) if log_r is None else log_r
    R, C = 1 << log_r, 1 << (log_n - log_r)
    return np.ascontiguousarray(np.concatenate([s.reshape(R, C // world, limbs) for s in shards], axis=1)).reshape(-1, limbs)


def output_location(k: int, log_n: int, world: int, log_r: int | None = None) -> tuple[int, int]:
    """(rank, element index inside that rank's output shard) of X[k] after FourStep.forward()."""
    log_r = default_log_r(log_n) if log_r is None else log_r
    R, C = 1 << log_r, 1 << (log_n - log_r)
    k1, k2 = k % R, k // R
    return k1 // (R // world), ((k2 % world) * (R // world) + k1 % (R // world)) * (C // world) + k2 // world


def geometric_shard(c: int, h: int, p: int, log_n: int, world: int, rank: int, log_r: int | None = None):
    """Factors of the input x[j] = c*h^j restricted to `rank`'s input shard (column block of the R x C
    view): the shard is the outer product col[j1] * row[j2_local] with col[j1] = c*h^(rank*w)*(h^C)^j1 and
    row[j2] = h^j2.  Returned as two uint32 arrays [R, 8] and [C/P, 8] (Python integers, no transform):
    a dense input whose transform has the closed form X[k] = c*(h^n - 1)/(h*w_n^k - 1), which is what lets a
    2^28-point distributed transform be checked exactly without a second one."""
    log_r = default_log_r(log_n) if log_r is None else log_r
    R, C = 1 << log_r, 1 << (log_n - log_r)
    w = C // world

    def powers(first, ratio, count):
        out = np.empty((count, 8), dtype=np.uint32)
        v = first % p
        for i in range(count):
            for l in range(8):
                out[i, l] = (v >> (32 * l)) & 0xFFFFFFFF
            v = v * ratio % p
        return out

    return powers(c * pow(h, rank * w, p), pow(h, C, p), R), powers(1, h, w)


def geometric_output(c: int, h: int, p: int, g: int, log_n: int, k: int) -> int:
    """X[k] of the transform of x[j] = c*h^j:  c*(h^n - 1)/(h*w_n^k - 1) mod p  (h*w_n^k != 1)."""
    n = 1 << log_n
    wk = pow(pow(g, (p - 1) >> log_n, p), k, p)
    den = (h * wk - 1) % p
    return c * (pow(h, n, p) - 1) * pow(den, p - 2, p) % p


def shard_batch(batch: int, world: int, rank: int) -> tuple[int, int]:
    """Transforms [lo, hi) of a batch handled by `rank` (independent units, no collective)."""
    per = (batch + world - 1) // world
    lo = min(rank * per, batch)
    return lo, min(lo + per, batch)


def exchange_reference(step1_out: list[np.ndarray], world: int) -> list[np.ndarray]:
    """What the all-to-all does, on host arrays (used by tests): rank s receives chunk s of every
    rank, concatenated by source rank."""
    chunks = [np.array_split(a, world, axis=0) for a in step1_out]
    return [np.concatenate([chunks[r][s] for r in range(world)], axis=0) for s in range(world)]


class FourStep:
    """Distributed forward NTT of 2^log_n points over `world` ranks (this process is `rank`).

    fused=False: step 1, NCCL all_to_all_single, step 2.
    fused=True : the exchange is the store phase of step 1's last kernel -- every output element
                 goes straight into the destination rank's receive buffer over NVLink
                 (peer-mapped memory shared through CUDA IPC); only a barrier separates the steps.
    barrier    : how the fused path separates the steps.  "flags" (default): the one-warp barrier kernel
                 ntt256_peer_barrier between the steps (the kernel boundary flushes step 1's peer stores);
                 "kernel": no barrier launch at all -- step 1's last kernel publishes an epoch from its last
                 CTA (one system-scope fence per CTA), step 2's first kernel starts with every CTA waiting for
                 all ranks' epochs (ntt256_ctx_set_peer_sync); "nccl": a one-element all-reduce.  Measured
                 at 2^24 on 8 B200 (profiles/r02_diag_8gpu_b.json): 489 / 526 / 507 us per transform.
    """

    def __init__(self, ctx, log_n: int, world: int, rank: int, log_r: int | None = None, group=None,
                 fused: bool = False, barrier: str = "flags"):
        self.ctx, self.log_n, self.world, self.rank, self.group = ctx, log_n, world, rank, group
        self.log_p = _log2(world)
        self.log_r = default_log_r(log_n) if log_r is None else log_r
        if self.log_r < self.log_p or log_n - self.log_r < self.log_p:
            raise ValueError("both R and C must be multiples of the world size")
        self.fused = bool(fused) and world > 1
        # how the fused path separates step 1 (peer stores) from step 2: "flags" = the library's
        # peer-memory barrier kernel (ntt256_peer_barrier), "nccl" = a one-element all-reduce
        if barrier not in ("kernel", "flags", "nccl"):
            raise ValueError("barrier must be 'kernel', 'flags' or 'nccl'")
        self.barrier = barrier
        self._epoch = 0
        self.a = None
        self.b = None
        self._peers = None  # [(ptrs of every rank's buffer a), (ptrs of every rank's buffer b)]
        self._keep, self._owned, self._opened = [], [], []

    def load_shard(self, shard):
        """shard: CUDA tensor with this rank's (R x C/P) input, 32 bytes per element."""
        import torch

        if self.fused:
            nbytes = shard.numel() * shard.element_size()
            self._dev = shard.device.index
            self.a, self._ha = self._ipc_buffer(nbytes, 0)
            self.b, self._hb = self._ipc_buffer(nbytes, 1)
            self._f, self._hf = self._ipc_buffer(4096, 2)
            self._f[1:].zero_()
            torch.cuda.synchronize()
            self._exchange_handles()
            self.a.copy_(shard.contiguous().view(-1).view(torch.int32))
            torch.cuda.synchronize()
            import torch.distributed as dist

            dist.barrier(group=self.group)
        else:
            self.a = shard.contiguous().view(-1)
            self.b = torch.empty_like(self.a)
        return self

    _MAGIC = 0x5A5A0000

    def _ipc_buffer(self, nbytes, which):
        """cudaMalloc'ed by the library (so one IPC handle describes it), viewed as a torch tensor.
        The driver carves small cudaMalloc requests out of shared blocks, and an IPC handle always
        maps the whole block: allocate at least 2 MiB so that the buffer IS the block, and stamp a
        magic word that the peers verify after opening the handle."""
        import ctypes

        import torch

        from .lib import Ntt256Error, lib

        ptr, handle = ctypes.c_void_p(), ctypes.create_string_buffer(64)
        rc = lib().ntt256_ipc_alloc(self._dev, max(nbytes, 2 << 20), ctypes.byref(ptr), handle)
        if rc:
            raise Ntt256Error(rc, "ipc_alloc")
        self._owned.append(ptr.value)

        class _Raw:
            __cuda_array_interface__ = {"shape": (nbytes // 4,), "typestr": "<i4", "data": (ptr.value, False),
                                        "version": 2, "strides": None}

        raw = _Raw()
        t = torch.as_tensor(raw, device=torch.device("cuda", self._dev))
        assert t.data_ptr() == ptr.value
        t[0] = self._MAGIC + 4 * self.rank + which
        torch.cuda.synchronize()
        self._keep.append(raw)
        return t, handle.raw

    def _exchange_handles(self):
        import ctypes

        import torch.distributed as dist

        from .lib import Ntt256Error, lib

        everyone = [None] * self.world
        dist.all_gather_object(everyone, (self._ha, self._hb, self._hf), group=self.group)
        ptrs = [[], [], []]
        for r, pair in enumerate(everyone):
            for k, h in enumerate(pair):
                if r == self.rank:
                    ptrs[k].append((self.a, self.b, self._f)[k].data_ptr())
                else:
                    p = ctypes.c_void_p()
                    rc = lib().ntt256_ipc_open(self._dev, h, ctypes.byref(p))
                    if rc:
                        raise Ntt256Error(rc, f"ipc_open of rank {r}'s buffer")
                    self._opened.append(p.value)
                    ptrs[k].append(p.value)
                    if self._peek(p.value) != self._MAGIC + 4 * r + k:
                        raise RuntimeError(f"IPC mapping of rank {r}'s buffer {k} does not start at the buffer")
        import torch

        self._peers = ptrs[:2]
        self._flag_ptrs = (ctypes.c_void_p * self.world)(*ptrs[2])
        if self.barrier == "kernel":
            self.ctx.set_peer_sync(ptrs[2], self.rank)
        self._flag = torch.zeros(1, dtype=torch.int32, device=torch.device("cuda", self._dev))
        self._cur = 0  # which buffer holds the input
        dist.barrier(group=self.group)

    def _step_barrier(self):
        """Every rank's step-1 kernel (and its peer stores) is complete before any rank's step 2
        starts; stream-ordered, the host does not block."""
        if self.barrier == "kernel":
            self._epoch += 1  # published and awaited inside the pass kernels: nothing to launch
        elif self.barrier == "flags":
            import torch

            from .lib import Ntt256Error, lib

            self._epoch += 1
            rc = lib().ntt256_peer_barrier(self._dev, self._flag_ptrs, self.world, self.rank, self._epoch,
                                           torch.cuda.current_stream().cuda_stream)
            if rc:
                raise Ntt256Error(rc, "peer_barrier")
        else:
            import torch.distributed as dist

            # a 1-element all-reduce completes on a rank only after every rank's NCCL kernel has
            # started, i.e. after every rank's step-1 kernel has finished
            dist.all_reduce(self._flag, group=self.group)

    def check(self):
        """After a synchronisation: raises if a rank-synchronisation wait of this rank ever timed out
        (a peer died or was more than the timeout late), i.e. if a result may be built on missing data."""
        import ctypes

        from .lib import lib

        if not self.fused or self._peers is None:
            return
        e = ctypes.c_ulonglong(0)
        if lib().ntt256_peer_sync_status(self._dev, self._f.data_ptr(), ctypes.byref(e)):
            raise RuntimeError(f"rank {self.rank}: peer synchronisation timed out waiting for epoch {e.value}")

    def _peek(self, ptr):
        import torch

        class _Raw:
            __cuda_array_interface__ = {"shape": (1,), "typestr": "<i4", "data": (ptr, False), "version": 2,
                                        "strides": None}

        return int(torch.as_tensor(_Raw(), device=torch.device("cuda", self._dev)).cpu()[0])

    def close(self):
        from .lib import lib

        import torch

        if self.fused and self._peers is not None:
            torch.cuda.synchronize()
            if self.barrier == "kernel":
                self.ctx.set_peer_sync(None)
            try:
                import torch.distributed as dist

                dist.barrier(group=self.group)
            except Exception:
                pass
            for p in self._opened:
                lib().ntt256_ipc_close(self._dev, p)
            for p in self._owned:
                lib().ntt256_ipc_free(self._dev, p)
            self._opened, self._owned, self._peers = [], [], None
            self.a = self.b = self._f = None

    def inverse(self):
        """Distributed inverse (with 1/n): consumes shards in the forward's OUTPUT layout and
        leaves shards in the forward's INPUT layout, so forward() then inverse() is the identity."""
        import torch.distributed as dist

        ctx = self.ctx
        if self.fused:
            src, dst = (self.a, self.b) if self._cur == 0 else (self.b, self.a)
            ctx.fourstep_inverse_step1(src, self.log_n, self.log_r, self.log_p, self.rank,
                                       peer_ptrs=self._peers[1 - self._cur])
            self._step_barrier()
            ctx.fourstep_inverse_step2(dst, self.log_n, self.log_r, self.log_p, self.rank)
            self._cur = 1 - self._cur
            return dst
        ctx.fourstep_inverse_step1(self.a, self.log_n, self.log_r, self.log_p, self.rank)
        if self.world > 1:
            dist.all_to_all_single(self.b, self.a, group=self.group)
            self.a, self.b = self.b, self.a
        ctx.fourstep_inverse_step2(self.a, self.log_n, self.log_r, self.log_p, self.rank)
        return self.a

    def forward(self):
        """One distributed transform.  The result shard is returned (and becomes the input of the
        next call: buffers ping-pong)."""
        import torch.distributed as dist

        ctx = self.ctx
        if self.fused:
            src, dst = (self.a, self.b) if self._cur == 0 else (self.b, self.a)
            ctx.fourstep_step1_peer(src, self._peers[1 - self._cur], self.log_n, self.log_r, self.log_p, self.rank)
            self._step_barrier()
            ctx.fourstep_step2(dst, self.log_n, self.log_r, self.log_p, self.rank)
            self._cur = 1 - self._cur
            return dst
        ctx.fourstep_step1(self.a, self.log_n, self.log_r, self.log_p, self.rank)
        if self.world > 1:
            dist.all_to_all_single(self.b, self.a, group=self.group)
            self.a, self.b = self.b, self.a
        ctx.fourstep_step2(self.a, self.log_n, self.log_r, self.log_p, self.rank)
        return self.a
