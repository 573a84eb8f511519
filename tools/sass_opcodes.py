#!/usr/bin/env python3
"""SASS evidence of the hardware paths, per kernel: python tools/sass_opcodes.py > profiles/r2_sass_opcodes.txt
Runs `cuobjdump -sass` on the in-tree library (no GPU needed) and counts the opcodes that prove a path is really used."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "orbslamm_b200", "liborbslamm_b200.so")
MARK = ["ATOM", "DMMA", "IDP.2A", "IDP.4A", "LDGSTS", "LDG.E.STRONG.SYS", "STG.E.STRONG.SYS", "ST.E.STRONG.SYS", "LD.E.STRONG.SYS", "MEMBAR.SC.SYS", "MEMBAR.ALL.SYS", "POPC", "RED.",
        "REDUX", "SYNCS", "UBLKCP", "UTMALDG", "VIMNMX3"]


def main():
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    names = {}
    dem = subprocess.run(["cu++filt"], input="\n".join(re.findall(r"Function : (\S+)", txt)), capture_output=True, text=True).stdout.split("\n")
    for m, d in zip(re.findall(r"Function : (\S+)", txt), dem):
        names[m] = re.sub(r"^void ", "", d).replace("(bool)", "").replace("(int)", "").split("(")[0]
    per = collections.OrderedDict()
    cur = None
    for line in txt.split("\n"):
        f = re.search(r"Function : (\S+)", line)
        if f:
            cur = names.get(f.group(1), f.group(1)); per[cur] = [0, collections.Counter()]
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d\s+)?([A-Z][A-Z0-9_.]*)", line)
        if cur and m:
            op = m.group(1)
            per[cur][0] += 1
            for k in MARK:
                if op.startswith(k):
                    per[cur][1][k] += 1
    print("# cuobjdump -sass orbslamm_b200/liborbslamm_b200.so (sm_100a, tools/sass_opcodes.py): occurrences of the opcodes that prove the hardware paths, per kernel (+ instruction count)")
    print("# DMMA = fp64 tensor-core MMA (mma.sync.m8n8k4.f64); UTMALDG = TMA tensor-map box load; UBLKCP = TMA bulk copy; VIMNMX3 = 3-input DPX min/max;")
    print("# IDP.4A/2A = dp4a/dp2a; LDGSTS = cp.async; SYNCS = mbarrier; *.STRONG.SYS / MEMBAR.*.SYS = system-scope accesses and fences of the NVLink peer exchange")
    for k, (n, c) in sorted(per.items(), key=lambda kv: (len(kv[0]), kv[0])):
        if k.startswith("orbs::") or "orbs::" in k:
            print(f"{k[:56]:56s} instrs={n:6d}  " + "  ".join(f"{o}={v}" for o, v in sorted(c.items())))


if __name__ == "__main__":
    sys.exit(main())
