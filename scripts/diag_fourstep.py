"""Where an N-GPU four-step transform spends its time (run under torchrun, one rank per GPU):
the fused step, its pieces alone, and the barrier alone; device-timed, max over ranks.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29511 scripts/diag_fourstep.py [log_n]"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from ntt_b200.api import Ntt256
from ntt_b200.dist import FourStep

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
ctx = Ntt256("bn254_fr", max_log_n=log_n, device=local)
fs = FourStep(ctx, log_n, world, rank, fused=True)
fs.load_shard(torch.randint(0, 2**28, ((1 << log_n) // world * 8,), dtype=torch.int32, device="cuda"))
flag = torch.zeros(1, device="cuda")
lp, lr = fs.log_p, fs.log_r


def timed(fn, steps=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps): fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / steps * 1e3], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return round(float(t.item()), 1)


def s1_peer():
    ctx.fourstep_step1_peer(fs.a, fs._peers[1], log_n, lr, lp, rank)


def s1_peer_ar():
    s1_peer(); dist.all_reduce(flag)


res = dict(gpus=world, log_n=log_n)
res["full_us"] = timed(fs.forward)
res["step1_local_us"] = timed(lambda: ctx.fourstep_step1(fs.b, log_n, lr, lp, rank))
res["step1_peer_us"] = timed(s1_peer)
res["step2_us"] = timed(lambda: ctx.fourstep_step2(fs.b, log_n, lr, lp, rank))
res["allreduce_us"] = timed(lambda: dist.all_reduce(flag))
res["step1_peer_allreduce_us"] = timed(s1_peer_ar)
res["full_again_us"] = timed(fs.forward)
dist.barrier()
fs.close()
# the same transform with the other two ways of separating the steps
for name in ("flags", "nccl"):
    f2 = FourStep(ctx, log_n, world, rank, fused=True, barrier=name)
    f2.load_shard(torch.randint(0, 2**28, ((1 << log_n) // world * 8,), dtype=torch.int32, device="cuda"))
    res[f"full_{name}_barrier_us"] = timed(f2.forward)
    dist.barrier()
    f2.close()
f3 = FourStep(ctx, log_n, world, rank, fused=False)
f3.load_shard(torch.randint(0, 2**28, ((1 << log_n) // world * 8,), dtype=torch.int32, device="cuda"))
res["full_nccl_all_to_all_us"] = timed(f3.forward)
if rank == 0:
    print(json.dumps(res), flush=True)
dist.barrier()
dist.destroy_process_group()
