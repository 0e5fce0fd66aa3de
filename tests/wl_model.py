"""Limb-level Python model of the wide-lazy field policy of Shoup contexts
(ntt_b200/csrc/field_wl.cuh + the generated field_shoup.cuh).  Test infrastructure.

Values are the raw 256-bit integers the registers hold (NOT residues).  Every operation asserts the claims the
CUDA code relies on -- no sum overflows 2^256, no difference goes negative, a correction never subtracts from a
value below the subtrahend -- and records the largest excess over 2p it has seen, so that a whole transform can be
run through tests/plan_sim.py with this arithmetic and compared with the oracle."""
from __future__ import annotations

import importlib.util
import os

M256 = (1 << 256) - 1


def _gen():
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scripts", "gen_shoup_mul.py")
    spec = importlib.util.spec_from_file_location("gen_shoup_mul", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


class WideLazyArith:
    def __init__(self, p: int, exact_products: bool = False):
        assert (1 << 229) < p and 5 * p + (1 << 226) < (1 << 256), "capi.cu shoup_capable"
        self.p = p
        self.p2 = 2 * p
        self.kp = 3 * p
        self.hi2p = self.p2 >> 192
        self.g = None if exact_products else _gen()
        self.max_excess = 0      # over every value handed to the next butterfly
        self.max_operand = 0     # over every product operand
        self.n_ops = {"add": 0, "sub_mul": 0, "sub": 0, "mul": 0, "canon": 0}

    # -- the building block: x += c when the top 64 bits of x exceed those of 2p
    def _above(self, x: int) -> bool:
        return (x >> 192) > self.hi2p

    def _note(self, x: int) -> int:
        assert 0 <= x <= M256
        self.max_excess = max(self.max_excess, x - self.p2)
        return x

    def add(self, a: int, b: int) -> int:
        self.n_ops["add"] += 1
        s = a + b
        assert s <= M256, "add_wl: a + b overflowed 2^256"
        if self._above(s):
            assert s >= self.p2
            s -= self.p2
        return self._note(s)

    def sub(self, a: int, b: int, to_mul: bool) -> int:
        if to_mul:
            self.n_ops["sub_mul"] += 1
            d = a - b + self.kp
            assert 0 < d <= M256, "sub_wl_mul: out of range"
            return d  # reduced by the product that follows; not a lazy-range value
        self.n_ops["sub"] += 1
        d = a - b
        if d < 0:
            d += self.kp
            assert d > 0, "sub_wl: the add-back constant was too small"
        if self._above(d):
            assert d >= self.p2
            d -= self.p2
        return self._note(d)

    def mul(self, x: int, w: int) -> int:
        self.n_ops["mul"] += 1
        assert 0 <= x <= M256 and 0 <= w < self.p
        self.max_operand = max(self.max_operand, x)
        if self.g is None:
            r0 = x * w % self.p  # stand-in with the same range semantics (fast path for big sizes)
            r = r0
        else:
            r, r0, _ = self.g.shoup_mul_model(x, w, self.p)
            assert r0 < 3 * self.p
        assert r % self.p == x * w % self.p
        assert r < self.p2 + (1 << 192)
        return self._note(r)

    def canon(self, x: int) -> int:
        self.n_ops["canon"] += 1
        assert 0 <= x < 6 * self.p and x <= M256
        if self._above(x):
            x -= self.p2
        if x >= self.p2:
            x -= self.p2
        if x >= self.p:
            x -= self.p
        assert 0 <= x < self.p
        return x
