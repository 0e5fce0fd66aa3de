"""Generates tests/golden/*.json.  Run in the build container (needs /root/reference for the
int64 vectors, which come from the reference's own compiled CPU code via oracle/_ref).
    python scripts/gen_golden.py
"""
import ctypes, hashlib, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from oracle import c_oracle as co, py_oracle as po

out_dir = os.path.join(ROOT, "tests", "golden")
os.makedirs(out_dir, exist_ok=True)

# ---- 1. vectors produced by the reference's own CPU functions (p = 469762049, g = 3, x[j] = j)
R = co.ref_lib()
assert R is not None, "oracle/_ref is missing: run `make -C oracle ref` where /root/reference exists"
i64p = ctypes.POINTER(ctypes.c_longlong)
ref = {"source": "NTT()/NTT_pro1+NTT_pro2 compiled from /root/reference/src/self-sort-in-place.cu via oracle/ref_harness.cu",
       "p": int(R.ref_modulus()), "g": int(R.ref_generator()), "input": "x[j] = j", "cases": []}
for log_n in (1, 2, 3, 4, 5, 8, 10, 12, 15, 16):
    n = 1 << log_n
    a = np.arange(n, dtype=np.int64); R.ref_ntt_i64(a.ctypes.data_as(i64p), log_n)
    b = np.arange(n, dtype=np.int64); R.ref_ssip_i64(b.ctypes.data_as(i64p), log_n)
    assert (a == b).all()
    case = {"log_n": log_n, "sha256_i64le": hashlib.sha256(a.tobytes()).hexdigest(), "head": [int(v) for v in a[:8]],
            "tail": [int(v) for v in a[-4:]]}
    if log_n <= 4:
        case["full"] = [int(v) for v in a]
    ref["cases"].append(case)
with open(os.path.join(out_dir, "ref_i64_kat.json"), "w") as f:
    json.dump(ref, f, indent=1)

# ---- 2. big-field vectors from the Python big-int restatement (ground truth)
big = {"source": "oracle/py_oracle.py::ntt_ref (restatement of big-num.cu:37-55)", "cases": []}
for F in (po.BN254_FR, po.BLS12_381_FR):
    for log_n, kind in ((3, "iota"), (10, "iota"), (12, "splitmix")):
        n = 1 << log_n
        x = list(range(n)) if kind == "iota" else po.splitmix_elements(n, F, 0x4E5454 + log_n)
        X = po.ntt_ref(x, F)
        assert po.ntt_ref(X, F, inverse=True) == x
        case = {"field": F.name, "log_n": log_n, "input": kind, "seed": 0x4E5454 + log_n,
                "sha256_image": po.sha256_image(X), "head": [hex(v) for v in X[:4]]}
        if log_n == 3:
            case["full"] = [hex(v) for v in X]
        big["cases"].append(case)
with open(os.path.join(out_dir, "bigfield_kat.json"), "w") as f:
    json.dump(big, f, indent=1)

# ---- 3. anchors quoted in SURVEY.md par. 8c (survey-time values, kept to detect oracle drift)
anchors = {
 "source": "SURVEY.md par. 8c, x[j] = j",
 "ref29": {"8": [28, 135891481, 78440360, 448772802, 469762045, 20989239, 391321681, 333870560],
           "16_head": [120, 32665148, 271782962, 334675195, 156880720, 418632167, 427783555, 428623306],
           "1024_head": [523776, 164093200, 443523204, 378291826, 30097604, 146615184, 328108782, 301427606],
           "omega_256": 338628632},
 "bn254_fr": {"sha256_2^10": "e232adf436abfe1394f922e14db59aef78b8212670f28053fba192695b6a3e4d",
              "sha256_2^16": "64f3b89e091af55095164643715a1ff775694622eab2ba8764a92a509847a9bf",
              "n8_X1": "0x002701a4fd3f1d3e7a309cdc72c7c8fcb5c94af009cb48e6e51461367a2f1796",
              "2^16_X1": "0x2dc947817fbfd6493f4a028587edcce87f8c99f3b94dee3129094775fecd5462"},
 "bls12_381_fr": {"sha256_2^10": "b0aa27cde7a1b8d1dcb802b0a5ab59270c1e62c5042a758086e62e1829b15489",
                  "sha256_2^16": "08f36b8fffc678fd1f973a7a61fc35f288e4231e852f8b6d8c3ae244ae3697e9",
                  "n8_X1": "0x3d9c9167f96a9b25495c51a9576083ab432e241ab8def899b6781127e7c9c15f",
                  "2^16_X1": "0x5fc118d58a2df300d36bb1591a262f7a183de9db22e88fd7a0282e462646cf9e"}}
with open(os.path.join(out_dir, "anchors.json"), "w") as f:
    json.dump(anchors, f, indent=1)
print("golden fixtures written to", out_dir)
