"""Host logic without a GPU: the C++ planner (through ntt256_plan_describe) and the descriptor
semantics, executed by the Python model of the kernel (plan_sim.py) against the oracle."""
import random

import pytest

import plan_sim
from ntt_b200.lib import plan_describe, plan_describe_384
from oracle import py_oracle as po


def _check_plan_invariants(plan, L, T, g):
    assert plan["log_n"] == L
    total = 0
    for pd in plan["passes"]:
        assert 3 <= pd["T"] <= T
        assert pd["T"] + pd["nblk_bits"] == L
        assert sorted(pd["in_pos"] + pd["blk_pos"]) == list(range(L))
        assert sorted(pd["in_pos"]) == sorted(pd["out_pos"])  # in place: a tile stores where it loaded
        if L > T:
            for b in range(min(g, L)):
                assert b in pd["in_pos"], "coalescing bits must be inside every tile"
        total += pd["d"]
        assert sum(pd["round_nb"]) == pd["d"]
        for r in range(pd["nrounds"]):
            sb = pd["round_sb"][r]
            tm = pd["tmap"][r]
            gb = plan.get("group_bits", 3)
            assert sorted(tm + [sb + i for i in range(gb)]) == list(range(pd["T"]))
    assert total == L
    assert plan["passes"][-1]["is_final"] == 1 and plan["passes"][-1]["has_tw"] == 0


@pytest.mark.parametrize("L,T,g", [(3, 3, 0), (5, 5, 0), (8, 8, 2), (6, 4, 1), (8, 4, 1), (9, 5, 1), (10, 6, 2),
                                   (11, 6, 1), (12, 6, 2), (12, 7, 2), (13, 8, 3), (12, 12, 2), (13, 10, 2)])
def test_plan_executes_to_the_oracle(built, L, T, g):
    plan = plan_describe(L, 14, T, g)
    assert "error" not in plan
    _check_plan_invariants(plan, L, T, g)
    rng = random.Random(L * 100 + T)
    for F in (po.BN254_FR, po.REF29):
        x = [rng.randrange(F.p) for _ in range(1 << L)]
        want = po.ntt_ref(x, F)
        assert plan_sim.run_plan(x, plan, F.p, F.g, 14) == want
        assert plan_sim.run_plan(want, plan, F.p, F.g, 14, inverse=True) == x


@pytest.mark.parametrize("L,T,g", [(4, 4, 0), (5, 5, 0), (7, 7, 0), (8, 8, 0), (10, 10, 0), (11, 11, 0), (12, 11, 0),
                                   (13, 7, 0), (12, 6, 1), (13, 11, 2), (9, 4, 0)])
@pytest.mark.parametrize("gb", [2, 1])
def test_small_kernel_plans_execute_to_the_oracle(built, L, T, g, gb):
    """group_bits = 2 / 1: the round schedules of ntt_pass4_kernel / ntt_pass2_kernel (small,
    latency-bound transforms)."""
    plan = plan_describe(L, 14, T, g, group_bits=gb)
    assert "error" not in plan
    assert plan["group_bits"] == gb
    _check_plan_invariants(plan, L, T, g)
    rng = random.Random(L * 100 + T + 7)
    for F in (po.BN254_FR, po.REF29):
        x = [rng.randrange(F.p) for _ in range(1 << L)]
        want = po.ntt_ref(x, F)
        assert plan_sim.run_plan(x, plan, F.p, F.g, 14) == want
        assert plan_sim.run_plan(want, plan, F.p, F.g, 14, inverse=True) == x


@pytest.mark.parametrize("L", [16, 20, 24, 26, 28])
def test_production_plans_exist_and_are_short(built, L):
    for T, g, max_passes in ((10, 2, 5), (11, 2, 5), (12, 2, 4)):
        plan = plan_describe(L, 28, T, g)
        assert "error" not in plan, (L, T, g)
        _check_plan_invariants(plan, L, T, g)
        assert len(plan["passes"]) <= max_passes
    # the headline configuration: 2^24 in three in-place passes with 128-byte runs
    if L == 24:
        assert len(plan_describe(24, 24, 12, 2)["passes"]) == 3


def test_every_small_configuration_executes_to_the_oracle(built):
    """Sweep of (log_n, elements per thread, tile bits, coalescing bits): every plan the planner
    returns must transform correctly (forward and inverse) in the kernel model; requests it cannot
    satisfy must say so (the library then relaxes them, capi.cu get_plan)."""
    F = po.REF29
    ok = noplan = 0
    for L in range(3, 10):
        x = [random.Random(L).randrange(F.p) for _ in range(1 << L)]
        want = po.ntt_ref(x, F)
        for gb in (1, 2, 3):
            for T in range(3, (12 if gb == 3 else 11) + 1):
                for g in range(0, min(T, 4) + 1):
                    plan = plan_describe(L, 14, T, g, group_bits=gb)
                    if "error" in plan:
                        noplan += 1
                        continue
                    assert plan_sim.run_plan(x, plan, F.p, F.g, 14) == want, (L, gb, T, g)
                    assert plan_sim.run_plan(want, plan, F.p, F.g, 14, inverse=True) == x, (L, gb, T, g)
                    ok += 1
    assert ok > 700 and noplan < ok // 8


@pytest.mark.parametrize("L", [3, 4, 5, 7, 9, 10, 11, 12, 13])
def test_384_bit_plans_execute_to_the_oracle(built, L):
    """The plans ntt384_forward runs (four elements per thread, 2^10-element tiles), executed by the
    kernel model over BLS12-377 Fq."""
    F = po.BLS12_377_FQ
    plan = plan_describe_384(L, 14)
    assert "error" not in plan and plan["group_bits"] == 2
    rng = random.Random(L)
    x = [rng.randrange(F.p) for _ in range(1 << L)]
    want = po.ntt_ref(x, F)
    assert plan_sim.run_plan(x, plan, F.p, F.g, 14) == want


@pytest.mark.parametrize("L,batch", [(3, 1), (4, 1), (5, 1), (8, 1), (10, 1), (12, 1), (13, 1), (14, 1), (15, 1),
                                     (16, 1), (3, 100000), (6, 10000), (9, 4096), (12, 256)])
def test_64_bit_plans_execute_to_the_oracle(built, L, batch):
    """The plans ntt64_forward runs for a call of `batch` transforms of 2^L points (sixteen elements per thread
    in radix-16 rounds, or four for small calls), executed by the kernel model over the Goldilocks field and
    over the reference's own modulus."""
    from ntt_b200.api64 import plan_describe64

    max_log_n = 16
    plan = plan_describe64(L, batch, max_log_n)
    assert "error" not in plan and plan["group_bits"] in (2, 4)
    assert all(pd["T"] >= plan["group_bits"] for pd in plan["passes"])
    _check_plan_invariants(plan, L, 14, 0)
    rng = random.Random(L * 7 + batch)
    for F in (po.Field("goldilocks", 0xFFFFFFFF00000001, 7), po.REF29):
        x = [rng.randrange(F.p) for _ in range(1 << L)]
        want = po.ntt_ref(x, F)
        assert plan_sim.run_plan(x, plan, F.p, F.g, max_log_n) == want
        assert plan_sim.run_plan(want, plan, F.p, F.g, max_log_n, inverse=True) == x
