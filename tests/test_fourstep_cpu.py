"""Multi-GPU host logic on CPU: the four-step data distribution (scatter, per-rank step plans,
all-to-all, gather) executed with the kernel model (plan_sim) must equal the oracle; and the
same exchange is run for real across two gloo ranks."""
import os
import random
import subprocess
import sys

import numpy as np
import pytest

import plan_sim
from ntt_b200 import dist as nd
from ntt_b200.lib import plan_describe_fourstep
from oracle import c_oracle as co
from oracle import py_oracle as po

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run_rank_step(step, shard_ints, log_n, log_r, log_p, rank, F, max_log_n, T, g):
    plan = plan_describe_fourstep(step, log_n, log_r, log_p, rank, max_log_n, T, g)
    assert "error" not in plan
    data = list(shard_ints)
    w_max = pow(F.g, (F.p - 1) >> max_log_n, F.p)
    for pd in plan["passes"]:
        plan_sim.run_pass(data, pd, F.p, w_max, max_log_n)
    return data


@pytest.mark.parametrize("log_n,log_r,world,T,g", [(8, 4, 2, 4, 1), (9, 5, 2, 5, 1), (10, 5, 4, 5, 1), (10, 6, 4, 8, 2),
                                                   (8, 4, 1, 5, 1), (11, 6, 8, 6, 1)])
def test_fourstep_model_equals_oracle(built, log_n, log_r, world, T, g):
    F = po.BLS12_381_FR
    rng = random.Random(log_n * 7 + world)
    x = [rng.randrange(F.p) for _ in range(1 << log_n)]
    want = po.ntt_ref(x, F)
    log_p = world.bit_length() - 1
    shards = nd.scatter_input(co.ints_to_array(x), log_n, world, log_r)
    s1 = [co.ints_to_array(_run_rank_step(1, co.array_to_ints(s), log_n, log_r, log_p, r, F, 12, T, g))
          for r, s in enumerate(shards)]
    recv = nd.exchange_reference(s1, world)
    s2 = [co.ints_to_array(_run_rank_step(2, co.array_to_ints(s), log_n, log_r, log_p, r, F, 12, T, g))
          for r, s in enumerate(recv)]
    got = co.array_to_ints(nd.gather_output(s2, log_n, world, log_r))
    assert got == want


@pytest.mark.parametrize("log_n,log_r,world", [(8, 4, 2), (10, 5, 4), (11, 6, 8)])
def test_fused_peer_store_equals_all_to_all(built, log_n, log_r, world):
    """The store phase of step 1's last pass in peer mode (ntt256_fourstep_step1_peer) must put
    every element exactly where step 1 + all-to-all would."""
    F = po.BN254_FR
    log_p = world.bit_length() - 1
    rng = random.Random(3)
    x = [rng.randrange(F.p) for _ in range(1 << log_n)]
    shards = nd.scatter_input(co.ints_to_array(x), log_n, world, log_r)
    w_max = pow(F.g, (F.p - 1) >> 12, F.p)
    plain, recv = [], [[None] * (1 << (log_n - log_p)) for _ in range(world)]
    for r, s in enumerate(shards):
        plan = plan_describe_fourstep(1, log_n, log_r, log_p, r, 12, 6, 1)
        a, b = co.array_to_ints(s), co.array_to_ints(s)
        for i, pd in enumerate(plan["passes"]):
            last = i == len(plan["passes"]) - 1
            plan_sim.run_pass(a, pd, F.p, w_max, 12)
            plan_sim.run_pass(b, pd, F.p, w_max, 12,
                              peer={"shift": log_n - 2 * log_p, "rank": r, "recv": recv} if last else None)
        plain.append(co.ints_to_array(a))
    want = nd.exchange_reference(plain, world)
    for r in range(world):
        assert recv[r] == co.array_to_ints(want[r])


@pytest.mark.parametrize("log_n,log_r,world,T,g", [(8, 4, 2, 5, 1), (10, 5, 4, 6, 1), (9, 5, 1, 5, 1), (11, 6, 8, 6, 1), (10, 6, 2, 8, 2)])
def test_distributed_inverse_undoes_the_forward_layouts(built, log_n, log_r, world, T, g):
    """Inverse four-step (plan kinds 3, 4): input = the forward's OUTPUT shards, output = the
    forward's INPUT shards, values = the oracle's inverse NTT."""
    F = po.BLS12_381_FR
    log_p = world.bit_length() - 1
    rng = random.Random(11 * log_n + world)
    x = [rng.randrange(F.p) for _ in range(1 << log_n)]
    X = po.ntt_ref(x, F)
    shards = nd.scatter_output_layout(co.ints_to_array(X), log_n, world, log_r)
    assert np.array_equal(nd.gather_output(shards, log_n, world, log_r), co.ints_to_array(X))
    w_max = pow(F.g, (F.p - 1) >> 12, F.p)
    scale = pow(1 << log_n, F.p - 2, F.p)

    def run(step, arr, rank):
        plan = plan_describe_fourstep(step, log_n, log_r, log_p, rank, 12, T, g)
        assert "error" not in plan
        data = co.array_to_ints(arr)
        for i, pd in enumerate(plan["passes"]):
            pd = dict(pd)
            sc = None
            if step == 4 and i == len(plan["passes"]) - 1:
                sc = scale
                if pd["nrounds"] >= 2:
                    pd["scaled_round"] = 0
                else:
                    pd["scale_out"] = 1
            plan_sim.run_pass(data, pd, F.p, w_max, 12, inverse=True, scale=sc)
        return co.ints_to_array(data)

    s1 = [run(3, s, r) for r, s in enumerate(shards)]
    recv = nd.exchange_reference(s1, world)
    s2 = [run(4, s, r) for r, s in enumerate(recv)]
    assert co.array_to_ints(nd.gather_input_layout(s2, log_n, world, log_r)) == x
    for r in range(world):
        assert np.array_equal(s2[r], nd.scatter_input(co.ints_to_array(x), log_n, world, log_r)[r])


def test_scatter_gather_are_inverse_views():
    log_n, world = 8, 4
    x = np.arange((1 << log_n) * 8, dtype=np.uint32).reshape(-1, 8)
    sh = nd.scatter_input(x, log_n, world)
    R, C = 16, 16
    assert sh[1].reshape(R, C // world, 8)[3, 2, 0] == x[3 * C + 1 * (C // world) + 2, 0]
    assert nd.shard_batch(256, 8, 3) == (96, 128)
    assert nd.shard_batch(5, 4, 3) == (5, 5)


_WORKER = r'''
import os, sys
sys.path.insert(0, os.environ["NTT_ROOT"]); sys.path.insert(0, os.path.join(os.environ["NTT_ROOT"], "tests"))
import numpy as np, torch, torch.distributed as dist
import plan_sim
from ntt_b200 import dist as nd
from ntt_b200.lib import plan_describe_fourstep
from oracle import c_oracle as co, py_oracle as po
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
log_n, log_r, log_p = 8, 4, 1
F = po.BN254_FR
x = co.splitmix_elements(1 << log_n, F.p, 4242)
shard = nd.scatter_input(x, log_n, world, log_r)[rank]
def run(step, arr):
    plan = plan_describe_fourstep(step, log_n, log_r, log_p, rank, 12, 5, 1)
    data = co.array_to_ints(arr); w = pow(F.g, (F.p - 1) >> 12, F.p)
    for pd in plan["passes"]:
        plan_sim.run_pass(data, pd, F.p, w, 12)
    return co.ints_to_array(data)
a = torch.from_numpy(run(1, shard).view(np.int32)).contiguous().view(-1)
b = torch.empty_like(a)
dist.all_to_all_single(b, a)           # the same call ntt_b200.dist.FourStep.forward issues over NCCL
out = run(2, b.numpy().view(np.uint32).reshape(-1, 8))
gathered = [torch.empty_like(a) for _ in range(world)]
dist.all_gather(gathered, torch.from_numpy(out.view(np.int32)).contiguous().view(-1))
if rank == 0:
    X = nd.gather_output([g.numpy().view(np.uint32).reshape(-1, 8) for g in gathered], log_n, world, log_r)
    assert np.array_equal(X, co.ntt(x, F.p, F.g)), "four-step over gloo differs from the oracle"
    print("FOURSTEP_GLOO_OK")
dist.destroy_process_group()
'''


def test_fourstep_two_gloo_ranks(built, tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    env = dict(os.environ, NTT_ROOT=ROOT)
    out = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
         "--master-port", "29611", str(script)], env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "FOURSTEP_GLOO_OK" in out.stdout


def test_output_location_and_geometric_input_helpers():
    """ntt_b200.dist.output_location / geometric_shard / geometric_output (used by bench.py's config-5 check
    and by the 8-GPU 2^28 test) against the oracle at sizes where everything can be enumerated."""
    import numpy as np

    from ntt_b200 import dist as nd
    from ntt_b200.api import FIELDS
    from oracle import c_oracle as co

    p, g = FIELDS["bn254_fr"]
    c, h = 0x1234567890ABCDEF1122334455667788 % p, 7
    for log_n, world in ((8, 4), (9, 2), (10, 8)):
        n = 1 << log_n
        xs = [c * pow(h, j, p) % p for j in range(n)]
        x = np.array([[(v >> (32 * l)) & 0xFFFFFFFF for l in range(8)] for v in xs], dtype=np.uint32)
        X = co.ntt(x, p, g)
        Xi = co.array_to_ints(X)
        for k in (0, 1, 5, n // 2, n - 1):
            assert nd.geometric_output(c, h, p, g, log_n, k) == Xi[k]
        shards = nd.scatter_input(x, log_n, world)
        for r in range(world):
            col, row = nd.geometric_shard(c, h, p, log_n, world, r)
            ci, ri = co.array_to_ints(col), co.array_to_ints(row)
            outer = np.array([[((a * b % p) >> (32 * l)) & 0xFFFFFFFF for l in range(8)] for a in ci for b in ri], dtype=np.uint32)
            assert np.array_equal(outer, shards[r]), (log_n, world, r)
        outs = nd.scatter_output_layout(X, log_n, world)
        for k in range(n):
            r, i = nd.output_location(k, log_n, world)
            assert np.array_equal(outs[r][i], X[k])


def test_shard_layout_helpers_for_other_limb_counts():
    """scatter_input / gather_input_layout and scatter_output_layout / gather_output are inverse pairs for 12-limb
    (384-bit) elements as well, and move whole elements (every limb of an element travels together)."""
    import numpy as np

    from ntt_b200 import dist as nd

    for limbs in (8, 12):
        for log_n, world in ((6, 2), (8, 4), (10, 8)):
            n = 1 << log_n
            x = (np.arange(n, dtype=np.uint32)[:, None] * 16 + np.arange(limbs, dtype=np.uint32)[None, :]).astype(np.uint32)
            shards = nd.scatter_input(x, log_n, world, limbs=limbs)
            assert all(s.shape == (n // world, limbs) for s in shards)
            assert all(np.array_equal(s % 16, np.tile(np.arange(limbs, dtype=np.uint32), (n // world, 1))) for s in shards)
            assert np.array_equal(nd.gather_input_layout(shards, log_n, world, limbs=limbs), x)
            outs = nd.scatter_output_layout(x, log_n, world, limbs=limbs)
            assert np.array_equal(nd.gather_output(outs, log_n, world, limbs=limbs), x)
            for k in (0, 1, n // 2 + 1, n - 1):
                r, idx = nd.output_location(k, log_n, world)
                assert np.array_equal(outs[r][idx], x[k])


@pytest.mark.parametrize("log_n,world", [(8, 2), (10, 4), (12, 8), (13, 2), (16, 4)])
def test_64_bit_fourstep_plans_execute_to_the_oracle(built, log_n, world):
    """The four-step plans the 64-bit library runs (sixteen or four elements per thread, its own tile choice),
    executed by the kernel model over Goldilocks: forward equals the oracle, the inverse pair returns the input
    shards (1/n through the scaled round table or the output scale, as the library does it)."""
    from ntt_b200.api64 import plan_describe64_fourstep

    F = po.Field("goldilocks", 0xFFFFFFFF00000001, 7)
    max_log_n = 16
    log_p, log_r = world.bit_length() - 1, nd.default_log_r(log_n)
    w_max = pow(F.g, (F.p - 1) >> max_log_n, F.p)
    rng = random.Random(64 * log_n + world)
    x = [rng.randrange(F.p) for _ in range(1 << log_n)]
    want = po.ntt_ref(x, F)

    def run(step, shard, rank, inverse=False):
        plan = plan_describe64_fourstep(step, log_n, log_r, log_p, rank, max_log_n)
        assert "error" not in plan and plan["group_bits"] in (2, 4)
        data = [int(v) for v in shard]
        passes = plan["passes"]
        for i, pd in enumerate(passes):
            pd = dict(pd)
            scale = None
            if step == 4 and i == len(passes) - 1:
                scale = pow(1 << log_n, F.p - 2, F.p)
                if pd["nrounds"] >= 2:
                    pd["scaled_round"] = 0
                else:
                    pd["scale_out"] = 1
            plan_sim.run_pass(data, pd, F.p, w_max, max_log_n, inverse, scale, group_bits=plan["group_bits"])
        return np.array(data, dtype=np.uint64).reshape(-1, 1)

    xs = np.array(x, dtype=np.uint64).reshape(-1, 1)
    ins = nd.scatter_input(xs, log_n, world, log_r, limbs=1)
    s1 = [run(1, s.reshape(-1), r) for r, s in enumerate(ins)]
    s2 = [run(2, s.reshape(-1), r) for r, s in enumerate(nd.exchange_reference(s1, world))]
    got = nd.gather_output(s2, log_n, world, log_r, limbs=1).reshape(-1)
    assert [int(v) for v in got] == want
    b1 = [run(3, s.reshape(-1), r, inverse=True) for r, s in enumerate(s2)]
    b2 = [run(4, s.reshape(-1), r, inverse=True) for r, s in enumerate(nd.exchange_reference(b1, world))]
    for r in range(world):
        assert np.array_equal(b2[r], ins[r]), r


@pytest.mark.parametrize("log_n,world", [(8, 2), (10, 4), (12, 8)])
def test_384_bit_fourstep_plans_execute_to_the_oracle(built, log_n, world):
    """The four-step plans the 384-bit library runs (four elements per thread, 2^9-element tiles), executed by the
    kernel model over BLS12-377 Fq: forward equals the oracle, the inverse pair returns the input shards."""
    from ntt_b200.lib import plan_describe_384_fourstep

    F = po.BLS12_377_FQ
    max_log_n = 14
    log_p, log_r = world.bit_length() - 1, nd.default_log_r(log_n)
    w_max = pow(F.g, (F.p - 1) >> max_log_n, F.p)
    rng = random.Random(384 * log_n + world)
    x = [rng.randrange(F.p) for _ in range(1 << log_n)]
    want = po.ntt_ref(x, F)

    def run(step, shard, rank, inverse=False):
        plan = plan_describe_384_fourstep(step, log_n, log_r, log_p, rank, max_log_n)
        assert "error" not in plan and plan["group_bits"] == 2
        data = list(shard)
        passes = plan["passes"]
        for i, pd in enumerate(passes):
            pd = dict(pd)
            scale = None
            if step == 4 and i == len(passes) - 1:
                scale = pow(1 << log_n, F.p - 2, F.p)
                if pd["nrounds"] >= 2:
                    pd["scaled_round"] = 0
                else:
                    pd["scale_out"] = 1
            plan_sim.run_pass(data, pd, F.p, w_max, max_log_n, inverse, scale, group_bits=2)
        return data

    # shards as Python integers: the layout helpers work on any [n, limbs] array, object dtype included
    xs = np.array(x, dtype=object).reshape(-1, 1)
    ins = nd.scatter_input(xs, log_n, world, log_r, limbs=1)
    s1 = [np.array(run(1, list(s.reshape(-1)), r), dtype=object).reshape(-1, 1) for r, s in enumerate(ins)]
    s2 = [np.array(run(2, list(s.reshape(-1)), r), dtype=object).reshape(-1, 1)
          for r, s in enumerate(nd.exchange_reference(s1, world))]
    assert list(nd.gather_output(s2, log_n, world, log_r, limbs=1).reshape(-1)) == want
    b1 = [np.array(run(3, list(s.reshape(-1)), r, inverse=True), dtype=object).reshape(-1, 1) for r, s in enumerate(s2)]
    b2 = [run(4, list(s.reshape(-1)), r, inverse=True) for r, s in enumerate(nd.exchange_reference(b1, world))]
    for r in range(world):
        assert b2[r] == list(ins[r].reshape(-1)), r
