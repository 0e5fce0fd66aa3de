"""Golden vectors for the 384-bit path from the REFERENCE'S OWN GZKP driver (run on a GPU box).

oracle/_ref/libref512.so = NTT_GZKP<16,512> of /root/reference/src/big-num.cu compiled unmodified
(oracle/ref512_harness.cu; 512 bits is the only wider container its one-limb-per-thread kernels accept).
A BLS12-377 Fq element is the same integer in a 384- or 512-bit container, so the reference's outputs
pin this repo's ntt384_* path: digests over the 48-byte-per-element image are written to
gpurun_out/ref512_gpu_kat.json (-> tests/golden/), together with whether this library and the
Python oracle produced the same bytes on the box."""
import ctypes, hashlib, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
R = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "libref512.so"))
u32p = ctypes.POINTER(ctypes.c_uint32)
R.ref512_ntt_gzkp.argtypes = [u32p, ctypes.c_uint, u32p, u32p]

from oracle import py_oracle as po
from ntt_b200.api import FIELDS384, Ntt384

p, g = FIELDS384["bls12_377_fq"]
F = po.BLS12_377_FQ


def limbs(v, nl):
    return np.array([(v >> (32 * k)) & 0xFFFFFFFF for k in range(nl)], dtype=np.uint32)


def to_arr(vals, nl):
    a = np.zeros((len(vals), nl), dtype=np.uint32)
    for i, v in enumerate(vals):
        a[i] = limbs(v, nl)
    return a


ctx = Ntt384("bls12_377_fq", max_log_n=16)
cases = []
for log_n in (5, 6, 8, 10, 12, 14):
    n = 1 << log_n
    for kind in ("iota", "random"):
        seed = 0x512000 + log_n
        if kind == "iota":
            vals = list(range(n))
        else:
            rng = np.random.default_rng(seed)
            vals = [int.from_bytes(rng.bytes(48), "little") % p for _ in range(n)]
        a512 = to_arr(vals, 16)
        rc = R.ref512_ntt_gzkp(a512.ctypes.data_as(u32p), log_n, limbs(p, 16).ctypes.data_as(u32p), limbs(g, 16).ctypes.data_as(u32p))
        assert rc == 0, rc
        assert not a512[:, 12:].any(), "reference output not reduced below 2^384"
        ref384 = np.ascontiguousarray(a512[:, :12])
        d = torch.from_numpy(to_arr(vals, 12).view(np.int32)).cuda()
        ctx.forward(d, log_n); torch.cuda.synchronize()
        ours = d.cpu().numpy().view(np.uint32).reshape(-1, 12)
        orc = to_arr(po.ntt_ref(vals, F), 12) if log_n <= 12 else None
        cases.append(dict(field="bls12_377_fq", log_n=log_n, input=kind, seed=(seed if kind == "random" else None),
                          sha256=hashlib.sha256(ref384.tobytes()).hexdigest(),
                          first=[hex(sum(int(ref384[i][k]) << (32 * k) for k in range(12))) for i in range(3)],
                          ours_equals_reference=bool(np.array_equal(ours, ref384)),
                          oracle_equals_reference=(None if orc is None else bool(np.array_equal(orc, ref384)))))
        print(log_n, kind, cases[-1]["ours_equals_reference"], cases[-1]["oracle_equals_reference"], flush=True)
out = dict(source="NTT_GZKP<16,512>(...,5,4) of /root/reference/src/big-num.cu, compiled unmodified (oracle/ref512_harness.cu), "
                  "run on a B200 by scripts/gen_golden_ref512.py over BLS12-377 Fq (g = 15)",
           digest="SHA-256 over the low 12 limbs (48 bytes) of every output element; the upper 4 limbs are zero",
           random_input="numpy default_rng(seed).bytes(48) little-endian mod p, element by element",
           cases=cases)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "ref512_gpu_kat.json"), "w"), indent=1)
print("ALL_AGREE", all(c["ours_equals_reference"] and c["oracle_equals_reference"] in (True, None) for c in cases))
