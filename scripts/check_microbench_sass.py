"""Opcode histogram of the hot loop of every kernel in a microbenchmark binary (cuobjdump -sass).

Round 1's pipe microbenchmark timed 64-bit adds where it meant IMAD.WIDE (ptxas hoisted the loop-invariant
product); this script is the guard: run it on the binary, read which SASS instructions each loop really
contains, and commit the table next to the timings.
Usage: python scripts/check_microbench_sass.py scripts/microbench_issue [out.json]
"""
import collections
import json
import re
import subprocess
import sys


def loop_histograms(binary):
    sass = subprocess.run(["cuobjdump", "-sass", binary], capture_output=True, text=True, check=True).stdout
    cur, body = None, collections.defaultdict(list)
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            continue
        m = re.search(r"/\*([0-9a-f]{4,5})\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)(.*?);", line)
        if m and cur:
            body[cur].append((int(m.group(1), 16), m.group(3), m.group(4)))
    out = {}
    for f, ins in body.items():
        best = None
        for a, op, rest in ins:
            if op.startswith("BRA"):
                t = re.search(r"0x([0-9a-f]+)", rest)
                if t:
                    tgt = int(t.group(1), 16)
                    if tgt < a and (best is None or a - tgt > best[1] - best[0]):
                        best = (tgt, a)
        name = subprocess.run(["c++filt", f], capture_output=True, text=True).stdout.strip().split("(")[0]
        if not best:
            out[name] = {"loop": None}
            continue
        c = collections.Counter(op for a, op, rest in ins if best[0] <= a <= best[1])
        out[name] = dict(sorted(c.items(), key=lambda kv: -kv[1]))
    return out


if __name__ == "__main__":
    h = loop_histograms(sys.argv[1])
    for k, v in h.items():
        print(k, v)
    if len(sys.argv) > 2:
        json.dump(h, open(sys.argv[2], "w"), indent=1)
