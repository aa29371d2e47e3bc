#!/usr/bin/env python
"""Registers, static shared memory and spills of every kernel (ptxas -v, sm_100a, the Makefile's flags):
    python tools/ptxas_summary.py > profiles/rNN_ptxas_v.tsv
No GPU needed (nvcc cross-compiles)."""
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "cup3d_b200", "csrc")
FILES = ["smooth_tma", "stencil7_tma", "advdiff_tma", "prhs_tma", "mg_kernels", "amr_kernels", "amr_advdiff",
         "adapt_kernels", "stencil_kernels", "blas_kernels", "comm", "obstacle", "capi"]


def demangle(name):
    m = re.search(r"(_Z\w+)$", name)
    d = subprocess.run(["c++filt", m.group(1) if m else name], capture_output=True, text=True).stdout.strip()
    d = d.replace("(anonymous namespace)::", "")
    d = re.sub(r"^void ", "", d)
    depth, cut = 0, len(d)
    for i, ch in enumerate(d):  # drop the argument list: the first '(' outside the template brackets
        if ch == "<":
            depth += 1
        elif ch == ">":
            depth -= 1
        elif ch == "(" and depth == 0:
            cut = i
            break
    return d[:cut]


def main():
    print("# ptxas -v, sm_100a, -O3: file, kernel, registers, static smem bytes, stack bytes, spill stores, spill loads")
    with tempfile.TemporaryDirectory() as tmp:
        procs = []
        for f in FILES:
            cmd = ["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo",
                   "-std=c++17", "--expt-relaxed-constexpr", "-Xptxas", "-v", "-dc", "-o",
                   os.path.join(tmp, f + ".o"), os.path.join(SRC, f + ".cu")]
            procs.append((f, subprocess.Popen(cmd, stderr=subprocess.PIPE, text=True)))
        for f, p in procs:
            txt = p.communicate()[1]
            if p.returncode != 0:
                sys.stderr.write(txt)
                sys.exit(1)
            for m in re.finditer(r"Compiling entry function '([^']+)' for 'sm_100a'\n(.*?)(?=ptxas info\s+: Compiling|\Z)",
                                 txt, re.S):
                body = m.group(2)
                regs = re.search(r"Used (\d+) registers", body)
                sm = re.search(r"(\d+) bytes smem", body)
                sp = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", body)
                print("\t".join([f, demangle(m.group(1)), regs.group(1) if regs else "?", sm.group(1) if sm else "0"] +
                                (list(sp.groups()) if sp else ["0", "0", "0"])))


if __name__ == "__main__":
    main()
