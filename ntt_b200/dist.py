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


def scatter_input(x: np.ndarray, log_n: int, world: int, log_r: int | None = None) -> list[np.ndarray]:
    """Natural-order x [n, 8] -> per-rank input shards (column blocks of the R x C view)."""
    log_r = default_log_r(log_n) if log_r is None else log_r
    R, C = 1 << log_r, 1 << (log_n - log_r)
    m = x.reshape(R, C, 8)
    w = C // world
    return [np.ascontiguousarray(m[:, r * w:(r + 1) * w]).reshape(-1, 8) for r in range(world)]


def gather_output(shards: list[np.ndarray], log_n: int, world: int, log_r: int | None = None) -> np.ndarray:
    """Per-rank output shards (element [k2 mod P][k1_local][k2 div P] = X[k1 + R*k2]) -> natural-order X."""
    log_r = default_log_r(log_n) if log_r is None else log_r
    R, C = 1 << log_r, 1 << (log_n - log_r)
    P = world
    # [rank][k2_low][k1_local][k2_high] -> X index k2*R + k1 with k2 = k2_low + P*k2_high, k1 = rank*(R/P) + k1_local
    m = np.stack([s.reshape(P, R // P, C // P, 8) for s in shards], axis=0)
    m = m.transpose(3, 1, 0, 2, 4)  # [k2_high][k2_low][rank][k1_local]
    return np.ascontiguousarray(m).reshape(-1, 8)


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
    """Distributed forward NTT of 2^log_n points over `world` ranks (this process is `rank`)."""

    def __init__(self, ctx, log_n: int, world: int, rank: int, log_r: int | None = None, group=None):
        self.ctx, self.log_n, self.world, self.rank, self.group = ctx, log_n, world, rank, group
        self.log_p = _log2(world)
        self.log_r = default_log_r(log_n) if log_r is None else log_r
        if self.log_r < self.log_p or log_n - self.log_r < self.log_p:
            raise ValueError("both R and C must be multiples of the world size")
        self.a = None
        self.b = None

    def load_shard(self, shard):
        """shard: CUDA tensor with this rank's (R x C/P) input, 32 bytes per element."""
        import torch

        self.a = shard.contiguous().view(-1)
        self.b = torch.empty_like(self.a)
        return self

    def forward(self):
        """step 1 -> all-to-all -> step 2.  The result shard is returned (and becomes the input of
        the next call: buffers ping-pong)."""
        import torch.distributed as dist

        ctx = self.ctx
        ctx.fourstep_step1(self.a, self.log_n, self.log_r, self.log_p, self.rank)
        if self.world > 1:
            dist.all_to_all_single(self.b, self.a, group=self.group)
            self.a, self.b = self.b, self.a
        ctx.fourstep_step2(self.a, self.log_n, self.log_r, self.log_p, self.rank)
        return self.a
