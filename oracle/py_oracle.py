"""TEST INFRASTRUCTURE ONLY -- never imported by the product path (ntt_b200/).

Python big-int restatement of the reference's NTT, the ground truth every other
implementation in this repo is compared against.  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline leg may import this module.

Parity status
-------------
* custom field p = 469762049, g = 3 (the only field the reference ever runs,
  /root/reference/src/big-num.cu:9-10): PINNED -- reproduces the known-answer vectors
  generated from the reference's own code (SURVEY.md par. 8c, tests/golden/ref_i64_kat.json)
  and, in this container, the reference's compiled CPU functions (oracle/_ref).
* BN254 Fr / BLS12-381 Fr: PINNED against the reference's own GPU drivers -- big-num.cu's
  NTT_GPU_Naive / NTT_GZKP take the modulus as a 256-bit argument; compiled unmodified
  (oracle/ref256_harness.cu) and run on a B200 they give the digests in
  tests/golden/ref256_gpu_kat.json, which this oracle reproduces (tests/test_oracle.py).
  Also anchored by the direct DFT definition and tests/golden/anchors.json.

Each function cites the reference lines it restates.
"""
from __future__ import annotations

import hashlib
from dataclasses import dataclass


@dataclass(frozen=True)
class Field:
    name: str
    p: int
    g: int  # multiplicative generator handed to the NTT as `omega` (big-num.cu:434-436)

    @property
    def two_adicity(self) -> int:
        t, m = 0, self.p - 1
        while m % 2 == 0:
            m //= 2
            t += 1
        return t

    def omega(self, n: int) -> int:
        """omega_n = g^((p-1)/n): big-num.cu:128-134 (cgbn_sub_ui32, div_ui32, modular_power)."""
        assert (self.p - 1) % n == 0
        return pow(self.g, (self.p - 1) // n, self.p)


REF29 = Field("ref29", 469762049, 3)  # big-num.cu:9-10
BN254_FR = Field(
    "bn254_fr", 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001, 5
)
BLS12_381_FR = Field(
    "bls12_381_fr", 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001, 7
)
# 384-bit containers (include/ntt384.h): BLS12-377 base field = BW6-761 scalar field, generator 15
BLS12_377_FQ = Field(
    "bls12_377_fq",
    0x01AE3A4617C510EAC63B05C06CA1493B1A22D9F300F5138F1EF3622FBA094800170B5D44300000008508C00000000001, 15
)
FIELDS = {f.name: f for f in (REF29, BN254_FR, BLS12_381_FR)}


def bit_reverse_table(log_n: int) -> list[int]:
    """big-num.cu:393-395."""
    n = 1 << log_n
    rev = [0] * n
    for i in range(n):
        rev[i] = (rev[i >> 1] >> 1) | ((i & 1) << (log_n - 1)) if log_n else 0
    return rev


def ntt_ref(x: list[int], field: Field, inverse: bool = False) -> list[int]:
    """Literal restatement of NTT() at big-num.cu:37-55: bit-reverse permute, then radix-2
    DIT stages with gap = omega^((P-1)/(2*stride)) and w stepped by gap inside each block.
    inverse=True follows the commented intent at big-num.cu:470-477 / GZKP-NTT.cu:1725-1732:
    same transform with inv(root), then multiply by inv(length)."""
    p = field.p
    n = len(x)
    log_n = n.bit_length() - 1
    assert n == 1 << log_n
    data = [v % p for v in x]
    rev = bit_reverse_table(log_n)
    for i in range(n):
        if i < rev[i]:
            data[i], data[rev[i]] = data[rev[i]], data[i]
    omega = pow(field.g, p - 2, p) if inverse else field.g
    stride = 1
    while stride < n:
        gap = pow(omega, (p - 1) // (stride << 1), p)
        for start in range(0, n, stride << 1):
            w = 1
            for offset in range(stride):
                a = data[start + offset]
                b = w * data[start + offset + stride] % p
                data[start + offset] = (a + b) % p
                data[start + offset + stride] = (a - b + p) % p
                w = gap * w % p
        stride <<= 1
    if inverse:
        co = pow(n, p - 2, p)
        data = [v * co % p for v in data]
    return data


def ntt_ssip(x: list[int], field: Field) -> list[int]:
    """Second, independent statement: self-sort-in-place ordering,
    NTT_pro1 + NTT_pro2 at /root/reference/src/self-sort-in-place.cu:79-128.
    Natural order in and out, no bit-reversal pass."""
    p = field.p
    n = len(x)
    log_n = n.bit_length() - 1
    data = [v % p for v in x]
    omega = field.g
    # NTT_pro1: plain DIF stages i = log_n .. log_n/2+1
    for i in range(log_n, log_n // 2, -1):
        stride = 1 << (i - 1)
        gap = pow(omega, (p - 1) // (stride << 1), p)
        for start in range(0, n, stride << 1):
            w = 1
            for offset in range(stride):
                a, b = data[start + offset], data[start + offset + stride]
                data[start + offset] = (a + b) % p
                data[start + offset + stride] = ((a - b + p) % p) * w % p
                w = gap * w % p
    # NTT_pro2: DIF stage on bit i-1 fused with the swap of bits i-1 <-> log_n-i
    for i in range(log_n // 2, 0, -1):
        stride = 1 << (i - 1)
        pair_stride = 1 << (log_n - i)
        gap = pow(omega, (p - 1) // (stride << 1), p)
        for start in range(0, n, pair_stride << 1):
            for offset0 in range(0, pair_stride, stride << 1):
                w = 1
                for offset in range(stride):
                    q = start + offset0 + offset
                    a, b = data[q], data[q + stride]
                    c, d = data[q + pair_stride], data[q + pair_stride + stride]
                    data[q] = (a + b) % p
                    data[q + stride] = (c + d) % p
                    data[q + pair_stride] = ((a - b + p) % p) * w % p
                    data[q + pair_stride + stride] = ((c - d + p) % p) * w % p
                    w = gap * w % p
    return data


def dft_direct(x: list[int], field: Field) -> list[int]:
    """The definition X[k] = sum_j x[j] w^(jk) (SURVEY.md par. 0) -- O(n^2), tiny n only."""
    p = field.p
    n = len(x)
    w = field.omega(n)
    return [sum(x[j] * pow(w, j * k, p) for j in range(n)) % p for k in range(n)]


def dft_sample(x: list[int], field: Field, k: int) -> int:
    """One output X[k] by Horner, for spot checks at sizes no full oracle can reach."""
    p = field.p
    wk = pow(field.omega(len(x)), k, p)
    acc = 0
    for v in reversed(x):
        acc = (acc * wk + v) % p
    return acc


# ---------------------------------------------------------------- memory image helpers
def to_limbs(v: int) -> list[int]:
    """cgbn_mem_t<256>: 8 little-endian u32 limbs (include/cgbn/cgbn_cuda.h:51-55)."""
    return [(v >> (32 * i)) & 0xFFFFFFFF for i in range(8)]


def from_limbs(limbs) -> int:
    return sum(int(l) << (32 * i) for i, l in enumerate(limbs))


def to_bytes(values: list[int]) -> bytes:
    return b"".join(int(v).to_bytes(32, "little") for v in values)


def from_bytes(buf: bytes) -> list[int]:
    return [int.from_bytes(buf[i : i + 32], "little") for i in range(0, len(buf), 32)]


def sha256_image(values: list[int]) -> str:
    return hashlib.sha256(to_bytes(values)).hexdigest()


def fnv1a64_i64(values: list[int]) -> str:
    """FNV-1a-64 over int64 outputs as used for the KATs in SURVEY.md par. 8c
    (xor the whole value, then multiply)."""
    h = 0xCBF29CE484222325
    for v in values:
        h ^= v & 0xFFFFFFFFFFFFFFFF
        h = (h * 0x100000001B3) & 0xFFFFFFFFFFFFFFFF
    return f"{h:016x}"


def splitmix_elements(n: int, field: Field, seed: int) -> list[int]:
    """Synthetic inputs: uniform canonical field elements from a counter-based SplitMix64
    stream (4 words per draw, top bits masked to the field's bit length, rejection on >= p).
    The C oracle and ntt_b200.testing generate the identical stream."""
    p = field.p
    bits = p.bit_length()
    mask = (1 << bits) - 1
    out = []
    state = seed & 0xFFFFFFFFFFFFFFFF
    M = 0xFFFFFFFFFFFFFFFF

    def nxt():
        nonlocal state
        state = (state + 0x9E3779B97F4A7C15) & M
        z = state
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M
        return z ^ (z >> 31)

    while len(out) < n:
        v = nxt() | (nxt() << 64) | (nxt() << 128) | (nxt() << 192)
        v &= mask
        if v < p:
            out.append(v)
    return out
