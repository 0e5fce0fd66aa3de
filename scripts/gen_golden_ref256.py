"""Golden vectors from the REFERENCE'S OWN GPU DRIVERS on the big fields (run on a GPU box).

oracle/_ref/libref256.so wraps /root/reference/src/big-num.cu's NTT_GPU_Naive<8,256> and
NTT_GZKP<8,256> (compiled unmodified, oracle/ref256_harness.cu).  The reference's main() only
exercises p = 469762049; its drivers take the modulus and generator as 256-bit values, so here they
are run on BN254 Fr and BLS12-381 Fr as well, on x[j] = j and on seeded random elements.  For every
case the digest (SHA-256 over the output's in-memory image) and a few leading elements are written to
gpurun_out/ref256_gpu_kat.json -> committed as tests/golden/ref256_gpu_kat.json, where the CPU
tests pin the oracle to it and the GPU tests pin the CUDA path to it.  The script also records, live,
whether this repo's library and the C oracle produced the same bytes."""
import ctypes, hashlib, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
R = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "libref256.so"))
u32p = ctypes.POINTER(ctypes.c_uint32)
R.ref256_ntt.argtypes = [u32p, ctypes.c_uint, u32p, u32p, ctypes.c_int]

from oracle import c_oracle as co
from ntt_b200.api import FIELDS, Ntt256
import torch


def limbs(v):
    return np.array([(v >> (32 * k)) & 0xFFFFFFFF for k in range(8)], dtype=np.uint32)


def ref_run(x, log_n, p, g, which):
    a = np.array(x, dtype=np.uint32, copy=True)
    pl, gl = limbs(p), limbs(g)
    rc = R.ref256_ntt(a.ctypes.data_as(u32p), log_n, pl.ctypes.data_as(u32p), gl.ctypes.data_as(u32p), which)
    if rc:
        raise RuntimeError(f"reference driver {which} failed with CUDA error {rc}")
    return a


def hexel(row):
    return hex(sum(int(row[k]) << (32 * k) for k in range(8)))


cases = []
for field in ("ref29", "bn254_fr", "bls12_381_fr"):
    p, g = FIELDS[field]
    ctx = Ntt256(field, max_log_n=16)
    for log_n in (5, 6, 8, 10, 12, 14, 16):
        n = 1 << log_n
        for kind in ("iota", "random"):
            seed = 0x256000 + log_n
            if kind == "iota":
                x = np.zeros((n, 8), dtype=np.uint32); x[:, 0] = np.arange(n, dtype=np.uint32)
            else:
                x = co.splitmix_elements(n, p, seed)
            naive = ref_run(x, log_n, p, g, 0)
            gzkp = ref_run(x, log_n, p, g, 1)
            d = torch.from_numpy(x.view(np.int32).copy()).cuda()
            ctx.forward(d, log_n); torch.cuda.synchronize()
            ours = d.cpu().numpy().view(np.uint32).reshape(-1, 8)
            orc = co.ntt(x, p, g, threads=8)
            cases.append(dict(field=field, log_n=log_n, input=kind, seed=(seed if kind == "random" else None),
                              sha256_naive=hashlib.sha256(naive.tobytes()).hexdigest(),
                              sha256_gzkp=hashlib.sha256(gzkp.tobytes()).hexdigest(),
                              first=[hexel(naive[i]) for i in range(4)],
                              drivers_agree=bool(np.array_equal(naive, gzkp)),
                              ours_equals_reference=bool(np.array_equal(ours, naive)),
                              oracle_equals_reference=bool(np.array_equal(orc, naive))))
            print(field, log_n, kind, cases[-1]["drivers_agree"], cases[-1]["ours_equals_reference"],
                  cases[-1]["oracle_equals_reference"], flush=True)
    ctx.close()
out = dict(source="NTT_GPU_Naive<8,256> / NTT_GZKP<8,256>(...,5,8) of /root/reference/src/big-num.cu, compiled unmodified "
                  "(oracle/ref256_harness.cu) and run on a B200 by scripts/gen_golden_ref256.py",
           digest="SHA-256 over the n*32 output bytes (8 little-endian u32 limbs per element)",
           cases=cases)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "ref256_gpu_kat.json"), "w"), indent=1)
print("ALL_AGREE", all(c["drivers_agree"] and c["ours_equals_reference"] and c["oracle_equals_reference"] for c in cases))
