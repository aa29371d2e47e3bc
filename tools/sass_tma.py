#!/usr/bin/env python
"""Which kernels of the built library stage through the TMA unit: counts of the SASS mnemonics of
cp.async.bulk / cp.async.bulk.tensor (UBLKCP / UTMALDG / UTMASTG), of the mbarrier waits (SYNCS) and of the
proxy fence (FENCE.VIEW.ASYNC), per kernel, from `cuobjdump -sass cup3d_b200/libcup3d_b200.so`.
    python tools/sass_tma.py > profiles/rNN_sass_tma.tsv      (no GPU needed)"""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "cup3d_b200", "libcup3d_b200.so")
PAT = ["UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "FENCE.VIEW.ASYNC", "MEMBAR.SC.SYS", "MEMBAR.ALL.SYS"]


def main():
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    names, counts = [], []
    for ln in txt.splitlines():
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            names.append(m.group(1))
            counts.append(dict.fromkeys(PAT, 0))
            continue
        if not names:
            continue
        for p in PAT:
            if re.search(r"\b" + re.escape(p), ln):
                counts[-1][p] += 1
    dem = subprocess.run(["c++filt"], input="\n".join(re.sub(r"^.*?(_Z\w+)$", r"\1", n) for n in names),
                         capture_output=True, text=True).stdout.splitlines()
    print("# SASS mnemonic counts per kernel of libcup3d_b200.so (sm_100a): " + ", ".join(PAT))
    print("# kernel\t" + "\t".join(PAT))
    for d, c in zip(dem, counts):
        if not any(c.values()):
            continue
        d = d.replace("(anonymous namespace)::", "")
        d = re.sub(r"^void ", "", d)
        depth, cut = 0, len(d)
        for i, ch in enumerate(d):
            if ch == "<":
                depth += 1
            elif ch == ">":
                depth -= 1
            elif ch == "(" and depth == 0:
                cut = i
                break
        print(d[:cut] + "\t" + "\t".join(str(c[p]) for p in PAT))


if __name__ == "__main__":
    main()
