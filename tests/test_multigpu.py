"""Real multi-GPU run of the four-step NTT (NCCL all-to-all) and of batch sharding.  Needs >= 2
GPUs: `gpurun --gpus 2 -- python -m pytest tests/test_multigpu.py -m gpu -q`; skipped on a 1-GPU box."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_WORKER = r'''
import os, sys
sys.path.insert(0, os.environ["NTT_ROOT"])
import numpy as np, torch, torch.distributed as dist
from ntt_b200 import dist as nd
from ntt_b200.api import Ntt256, FIELDS
from oracle import c_oracle as co
rank, world, lr_ = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr_)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr_))
p, g = FIELDS["bls12_381_fr"]
ctx = Ntt256("bls12_381_fr", max_log_n=22, device=lr_)
for fused, barrier in ((False, "nccl"), (True, "nccl"), (True, "flags"), (True, "kernel")):
  for log_n in (14, 20):
    x = co.splitmix_elements(1 << log_n, p, 99 + log_n)          # same on every rank
    shard = nd.scatter_input(x, log_n, world)[rank]
    fs = nd.FourStep(ctx, log_n, world, rank, fused=fused, barrier=barrier).load_shard(torch.from_numpy(shard.view(np.int32)).cuda().view(-1))
    out = fs.forward().clone()
    torch.cuda.synchronize()
    gathered = [torch.empty_like(out) for _ in range(world)]
    dist.all_gather(gathered, out)
    if rank == 0:
        X = nd.gather_output([t.cpu().numpy().view(np.uint32).reshape(-1, 8) for t in gathered], log_n, world)
        assert np.array_equal(X, co.ntt(x, p, g, threads=8)), f"four-step 2^{log_n} over {world} GPUs (fused={fused}) differs from the oracle"
    # distributed inverse consumes the forward's output shards and must return the input shard
    back = fs.inverse().clone()
    torch.cuda.synchronize()
    assert np.array_equal(back.cpu().numpy().view(np.uint32).reshape(-1, 8), shard), f"inverse(forward(shard)) != shard (fused={fused}, 2^{log_n})"
    # another forward reuses the ping-pong buffers
    out2 = fs.forward().clone()
    torch.cuda.synchronize()
    assert torch.equal(out2, out)
    # back-to-back transforms without host synchronisation: forward/inverse pairs must keep
    # restoring the shard (exercises the barrier's epoch counting and the ping-pong buffers)
    for _ in range(6):
        fs.inverse(); fs.forward()
    out3 = fs.inverse().clone()
    torch.cuda.synchronize()
    assert np.array_equal(out3.cpu().numpy().view(np.uint32).reshape(-1, 8), shard)
    dist.barrier()
    fs.check()
    if fused:
        fs.close()
# batch sharding: independent transforms, no collective
batch, log_n = 16, 12
lo, hi = nd.shard_batch(batch, world, rank)
x = co.splitmix_elements(batch << log_n, p, 7)
mine = torch.from_numpy(x[lo << log_n: hi << log_n].view(np.int32)).cuda().view(-1)
ctx.forward(mine, log_n, batch=hi - lo)
torch.cuda.synchronize()
got = mine.cpu().numpy().view(np.uint32).reshape(-1, 8)
for b in range(lo, hi):
    assert np.array_equal(got[(b - lo) << log_n:(b - lo + 1) << log_n], co.ntt(x[b << log_n:(b + 1) << log_n], p, g))
dist.barrier()
if rank == 0:
    print("MULTIGPU_OK", world)
dist.destroy_process_group()
'''


_WORKER_24 = r'''
import os, sys
sys.path.insert(0, os.environ["NTT_ROOT"])
import numpy as np, torch, torch.distributed as dist
from ntt_b200 import dist as nd
from ntt_b200.api import Ntt256, FIELDS
from oracle import c_oracle as co
rank, world, lr_ = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr_)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr_))
# BASELINE config 4 at full size on every GPU of the box: all 2^24 outputs against the oracle
# (/root/reference/src/big-num.cu:446-467 compares every element of every size it runs)
log_n = 24
for field in ("bn254_fr", "bls12_381_fr"):
    p, g = FIELDS[field]
    ctx = Ntt256(field, max_log_n=log_n, device=lr_)
    shards = [co.splitmix_elements((1 << log_n) // world, p, 2400 + r) for r in range(world)] if rank == 0 else None
    mine = co.splitmix_elements((1 << log_n) // world, p, 2400 + rank)
    fs = nd.FourStep(ctx, log_n, world, rank, fused=True).load_shard(torch.from_numpy(mine.view(np.int32)).cuda().view(-1))
    out = fs.forward().clone()
    torch.cuda.synchronize()
    fs.check()
    gathered = [torch.empty_like(out) for _ in range(world)] if rank == 0 else None
    dist.gather(out, gathered, dst=0)
    if rank == 0:
        X = nd.gather_output([t.cpu().numpy().view(np.uint32).reshape(-1, 8) for t in gathered], log_n, world)
        x = nd.gather_input_layout(shards, log_n, world)
        want = co.ntt(x, p, g, threads=max(1, min(co.max_threads(), os.cpu_count() or 1)))
        assert np.array_equal(X, want), f"2^24 {field} over {world} GPUs differs from the oracle"
    back = fs.inverse().clone()
    torch.cuda.synchronize()
    assert np.array_equal(back.cpu().numpy().view(np.uint32).reshape(-1, 8), mine)
    dist.barrier()
    fs.close()
    ctx.close()
if rank == 0:
    print("MULTIGPU24_OK", world)
dist.destroy_process_group()
'''

_WORKER_28 = r'''
import os, sys
sys.path.insert(0, os.environ["NTT_ROOT"])
import numpy as np, torch, torch.distributed as dist
from ntt_b200 import dist as nd
from ntt_b200.api import Ntt256, FIELDS
rank, world, lr_ = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr_)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr_))
# BASELINE config 5: 2^28 points (8 GiB) over 8 GPUs.  x[j] = c*h^j, so X[k] = c(h^n - 1)/(h w^k - 1):
# 512 sampled outputs per rank are checked against the closed form, then the distributed inverse must
# return the input shard bit for bit.
p, g = FIELDS["bn254_fr"]
log_n = 28
n = 1 << log_n
c, h = 0x1234567890ABCDEF1122334455667788 % p, 7
ctx = Ntt256("bn254_fr", max_log_n=log_n, device=lr_)
fs = nd.FourStep(ctx, log_n, world, rank, fused=True)
col, row = nd.geometric_shard(c, h, p, log_n, world, rank, fs.log_r)
R, w = col.shape[0], row.shape[0]
a = torch.from_numpy(col.view(np.int32)).cuda().view(R, 1, 8).expand(R, w, 8).contiguous().view(-1)
b = torch.from_numpy(row.view(np.int32)).cuda().view(1, w, 8).expand(R, w, 8).contiguous().view(-1)
ctx.vec_mul(a, b, out=a)
del b
src = a.clone()
fs.load_shard(a)
out = fs.forward()
torch.cuda.synchronize()
fs.check()
rng = np.random.default_rng(28 + rank)
ks = [0, 1, n - 1, n // 2, n // 2 + 7] + [int(v) for v in rng.integers(0, n, size=4096)]
checked = 0
for k in ks:
    owner, idx = nd.output_location(k, log_n, world, fs.log_r)
    if owner != rank:
        continue
    v = out.view(-1)[8 * idx: 8 * idx + 8].cpu().numpy().view(np.uint32)
    assert sum(int(v[i]) << (32 * i) for i in range(8)) == nd.geometric_output(c, h, p, g, log_n, k), (rank, k)
    checked += 1
assert checked >= 256
back = fs.inverse()
torch.cuda.synchronize()
assert torch.equal(back.view(-1), src.view(-1)), "inverse(forward(shard)) != shard at 2^28"
t = torch.tensor([checked], device="cuda")
dist.all_reduce(t)
fs.check()
dist.barrier()
fs.close()
if rank == 0:
    print("MULTIGPU28_OK", world, int(t.item()))
dist.destroy_process_group()
'''


_WORKER_384 = r'''
import os, sys
sys.path.insert(0, os.environ["NTT_ROOT"])
import numpy as np, torch, torch.distributed as dist
from ntt_b200 import dist as nd
from ntt_b200.api import Ntt384, FIELDS384
from oracle import c_oracle as co
rank, world, lr_ = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr_)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr_))
# 384-bit containers over every GPU of the box (BLS12-377 Fq): the same FourStep driver, the 12-limb instances of
# the pass kernels; NCCL all-to-all and the fused peer-store exchange with the synchronisation inside the kernels
p, g = FIELDS384["bls12_377_fq"]
ctx = Ntt384("bls12_377_fq", max_log_n=22, device=lr_)
def rand384(seed, n):
    x = np.random.default_rng(seed).integers(0, 1 << 32, size=(n, 12), dtype=np.uint64).astype(np.uint32)
    x[:, 11] &= 0x00FFFFFF
    return x
for fused, barrier in ((False, "nccl"), (True, "nccl"), (True, "kernel")):
  for log_n in (14, 20):
    x = rand384(384 + log_n, 1 << log_n)                       # same on every rank
    shard = nd.scatter_input(x, log_n, world, limbs=12)[rank]
    fs = nd.FourStep(ctx, log_n, world, rank, fused=fused, barrier=barrier).load_shard(torch.from_numpy(shard.view(np.int32)).cuda().view(-1))
    out = fs.forward().clone()
    torch.cuda.synchronize()
    gathered = [torch.empty_like(out) for _ in range(world)]
    dist.all_gather(gathered, out)
    if rank == 0:
        X = nd.gather_output([t.cpu().numpy().view(np.uint32).reshape(-1, 12) for t in gathered], log_n, world, limbs=12)
        assert np.array_equal(X, co.ntt384(x, p, g, threads=8)), f"384-bit four-step 2^{log_n} over {world} GPUs (fused={fused}) differs from the oracle"
    back = fs.inverse().clone()
    torch.cuda.synchronize()
    assert np.array_equal(back.cpu().numpy().view(np.uint32).reshape(-1, 12), shard), f"inverse(forward(shard)) != shard (fused={fused}, 2^{log_n})"
    for _ in range(4):
        fs.forward(); fs.inverse()
    torch.cuda.synchronize()
    src = fs.a if not fused else (fs.a if fs._cur == 0 else fs.b)
    assert np.array_equal(src.cpu().numpy().view(np.uint32).reshape(-1, 12), shard)
    dist.barrier()
    fs.check()
    if fused:
        fs.close()
if rank == 0:
    print("MULTIGPU384_OK", world)
dist.destroy_process_group()
'''


_WORKER_64 = r'''
import os, sys
sys.path.insert(0, os.environ["NTT_ROOT"])
import numpy as np, torch, torch.distributed as dist
from ntt_b200 import dist as nd
from ntt_b200.api64 import Ntt64, FIELDS64
from oracle import c_oracle as co
rank, world, lr_ = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr_)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr_))
# 64-bit containers (Goldilocks) over every GPU of the box: the same FourStep driver; NCCL all-to-all and the fused
# peer-store exchange behind the flag barrier
p, g = FIELDS64["goldilocks"]
ctx = Ntt64("goldilocks", max_log_n=24, device=lr_)
for fused, barrier in ((False, "nccl"), (True, "nccl"), (True, "flags")):
  for log_n in (14, 20, 24):
    x = co.splitmix64_elements(1 << log_n, p, 64 + log_n)     # same on every rank
    shard = nd.scatter_input(x.reshape(-1, 1), log_n, world, limbs=1)[rank].reshape(-1)
    fs = nd.FourStep(ctx, log_n, world, rank, fused=fused, barrier=barrier).load_shard(torch.from_numpy(shard.view(np.int64)).cuda())
    out = fs.forward().clone()
    torch.cuda.synchronize()
    gathered = [torch.empty_like(out) for _ in range(world)]
    dist.all_gather(gathered, out)
    if rank == 0:
        X = nd.gather_output([t.cpu().numpy().view(np.uint64).reshape(-1, 1) for t in gathered], log_n, world, limbs=1).reshape(-1)
        assert np.array_equal(X, co.ntt64(x, p, g, threads=16)), f"64-bit four-step 2^{log_n} over {world} GPUs (fused={fused}) differs from the oracle"
    back = fs.inverse().clone()
    torch.cuda.synchronize()
    assert np.array_equal(back.cpu().numpy().view(np.uint64).reshape(-1), shard), f"inverse(forward(shard)) != shard (fused={fused}, 2^{log_n})"
    for _ in range(4):
        fs.forward(); fs.inverse()
    torch.cuda.synchronize()
    src = fs.a if not fused else (fs.a if fs._cur == 0 else fs.b)
    assert np.array_equal(src.cpu().numpy().view(np.uint64).reshape(-1), shard)
    dist.barrier()
    fs.check()
    if fused:
        fs.close()
if rank == 0:
    print("MULTIGPU64_OK", world)
dist.destroy_process_group()
'''


def _run_workers(tmp_path, source, world, port, timeout=1500):
    script = tmp_path / "worker.py"
    script.write_text(source)
    env = dict(os.environ, NTT_ROOT=ROOT)
    return subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr",
         "127.0.0.1", "--master-port", str(port), str(script)], env=env, capture_output=True, text=True, timeout=timeout)


def test_config4_full_2p24_over_all_gpus(tmp_path):
    """Every output of the 2^24-point transform on 2/4/8 GPUs against the oracle, both fields, the default
    exchange (peer stores + synchronisation folded into the kernels), then the distributed inverse."""
    import torch

    if not torch.cuda.is_available():
        pytest.fail("needs CUDA")
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("single-GPU box")
    out = _run_workers(tmp_path, _WORKER_24, 1 << (n.bit_length() - 1), 29641)
    assert out.returncode == 0, (out.stdout[-1500:] + out.stderr[-3000:])
    assert "MULTIGPU24_OK" in out.stdout


def test_config5_2p28_over_8_gpus_closed_form(tmp_path):
    import torch

    if not torch.cuda.is_available():
        pytest.fail("needs CUDA")
    if torch.cuda.device_count() < 8:
        pytest.skip("needs the 8 GPUs of one box")
    out = _run_workers(tmp_path, _WORKER_28, 8, 29651)
    assert out.returncode == 0, (out.stdout[-1500:] + out.stderr[-3000:])
    assert "MULTIGPU28_OK" in out.stdout


def test_384_bit_fourstep_over_all_gpus(tmp_path):
    import torch

    if not torch.cuda.is_available():
        pytest.fail("needs CUDA")
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("single-GPU box: covered by the emulated ranks of tests/test_gpu_384.py")
    out = _run_workers(tmp_path, _WORKER_384, 1 << (n.bit_length() - 1), 29661)
    assert out.returncode == 0, (out.stdout[-1500:] + out.stderr[-3000:])
    assert "MULTIGPU384_OK" in out.stdout


def test_64_bit_fourstep_over_all_gpus(tmp_path):
    import torch

    if not torch.cuda.is_available():
        pytest.fail("needs CUDA")
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("single-GPU box: covered by the emulated ranks of tests/test_gpu_64.py")
    out = _run_workers(tmp_path, _WORKER_64, 1 << (n.bit_length() - 1), 29671)
    assert out.returncode == 0, (out.stdout[-1500:] + out.stderr[-3000:])
    assert "MULTIGPU64_OK" in out.stdout


def test_fourstep_and_batch_over_nccl(tmp_path):
    import torch

    if not torch.cuda.is_available():
        pytest.fail("needs CUDA")
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("single-GPU box: the multi-rank data flow is covered by test_fourstep_ranks_emulated_on_one_gpu")
    world = 1 << (n.bit_length() - 1)
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    env = dict(os.environ, NTT_ROOT=ROOT)
    out = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr",
         "127.0.0.1", "--master-port", "29631", str(script)], env=env, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, (out.stdout[-1500:] + out.stderr[-3000:])
    assert "MULTIGPU_OK" in out.stdout


def test_single_process_mgpu_driver():
    """ntt256_mgpu_*: one host thread, all GPUs of the box, fused peer-store exchange ordered by
    cross-device events.  Forward equals the oracle, inverse restores the input shards, batches run
    independently per GPU."""
    import numpy as np
    import torch

    if not torch.cuda.is_available():
        pytest.fail("needs CUDA")
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("single-GPU box")
    world = 1 << (n.bit_length() - 1)
    sys.path.insert(0, ROOT)
    from ntt_b200 import dist as nd
    from ntt_b200.api import FIELDS, Ntt256MultiGpu
    from oracle import c_oracle as co

    p, g = FIELDS["bn254_fr"]
    mg = Ntt256MultiGpu(world, "bn254_fr", max_log_n=22)
    for log_n in (12, 20, 21):
        x = co.splitmix_elements(1 << log_n, p, 555 + log_n)
        host = nd.scatter_input(x, log_n, world)
        shards = [torch.from_numpy(h.view(np.int32)).to(f"cuda:{i}").view(-1) for i, h in enumerate(host)]
        for rep in range(2):  # twice: buffers and events are reused
            mg.forward(shards, log_n)
            mg.synchronize()
            outs = [s.cpu().numpy().view(np.uint32).reshape(-1, 8) for s in shards]
            assert np.array_equal(nd.gather_output(outs, log_n, world), co.ntt(x, p, g, threads=8)), (log_n, rep)
            mg.inverse(shards, log_n)
            mg.synchronize()
            for i in range(world):
                assert np.array_equal(shards[i].cpu().numpy().view(np.uint32).reshape(-1, 8), host[i]), (log_n, rep, i)
    log_n, per = 12, 3
    xs = [co.splitmix_elements(per << log_n, p, 900 + i) for i in range(world)]
    shards = [torch.from_numpy(h.view(np.int32)).to(f"cuda:{i}").view(-1) for i, h in enumerate(xs)]
    mg.forward_batch(shards, log_n, per)
    mg.synchronize()
    for i in range(world):
        got = shards[i].cpu().numpy().view(np.uint32).reshape(-1, 8)
        for b in range(per):
            assert np.array_equal(got[b << log_n:(b + 1) << log_n], co.ntt(xs[i][b << log_n:(b + 1) << log_n], p, g))
    mg.close()
