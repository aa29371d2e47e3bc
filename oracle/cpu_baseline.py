#!/usr/bin/env python3
"""oracle/cpu_baseline.py -- TEST/BENCH INFRASTRUCTURE ONLY.

Times the reference's own CPU V-cycle (mg_vcycle, main.c:4831) on the host
cores: oracle/_ref (the unmodified reference, OpenMP over all cores) when it
was built, else the C restatement oracle/libcup_oracle.so.  Run as a
subprocess by bench.py because the reference keeps its mesh in file-statics.
Prints one JSON line.
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--level", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--dump", default="", help="write the V-cycle output of the point-source RHS to this .npy")
    a = ap.parse_args()
    L = a.level
    from oracle import refbind as R
    if R.available():
        R.init(levelStart=L, levelMax=L + 1)
        n = R.nblk()
        ib, rb = R.blocks()
        b = np.zeros((n, 512))
        lo = rb[:, 1:4]
        hi = lo + 8 * rb[:, 0:1]
        for p, v in ((0.25, 1.0), (0.75, -1.0)):
            i = int(np.nonzero(np.all((lo <= p) & (p < hi), axis=1))[0][0])
            b[i, 0] = v
        # The reference's schedule(dynamic,1) block loops do not scale monotonically with the thread
        # count (they collapse when the host is oversubscribed or crosses sockets), so the count is
        # chosen by a scan: per candidate one warm-up cycle and the best of two timed ones, every
        # candidate measured (no early exit), the scan reported.  CUP_REF_THREADS / OMP_NUM_THREADS pin it.
        import ctypes
        gomp = ctypes.CDLL("libgomp.so.1")
        ncpu = os.cpu_count() or 1
        cand = sorted({t for t in (8, 16, 24, 32, 48, 64, 96, 128, 192, ncpu // 2, ncpu) if 1 <= t <= ncpu})
        if os.environ.get("OMP_NUM_THREADS"):
            cand = [int(os.environ["OMP_NUM_THREADS"])]
        scan = {}
        best, best_t = None, None
        y = None
        bad = 0
        for t in cand:
            gomp.omp_set_num_threads(t)
            s1 = R.time_vcycle(b, 1, 1) if len(cand) > 1 else 0.0
            if len(cand) > 1 and (best is None or s1 < 1.5 * best):
                s1 = min(s1, R.time_vcycle(b, 0, 1))  # a contender: best of two
            scan[t] = round(1e3 * s1, 2)
            if best is None or s1 < best:
                best, best_t = s1, t
            # oversubscription only gets worse (128 threads: 20 s per 256^3 cycle on the r02 box, 140x the
            # best): after TWO consecutive counts more than 1.5x slower than the best the larger ones are
            # not tried -- one bad count alone does not end the scan
            bad = bad + 1 if s1 > 1.5 * best else 0
            if bad >= 2:
                scan["stopped_after"] = t
                break
        gomp.omp_set_num_threads(best_t)
        sec = R.time_vcycle(b, a.warmup, a.steps)
        if a.dump:
            np.save(a.dump, R.mg_vcycle(b))
        kind, threads = "reference", best_t
        what = ("unmodified reference main.c (oracle/_ref), gcc -O3 -fopenmp, %d OpenMP threads (thread scan, ms "
                "per cycle: %s; %d logical CPUs), 1 rank" % (threads, scan, ncpu))
    else:
        from oracle import portbind as P
        sec, n, threads = P.time_vcycle_uniform(L, a.warmup, a.steps)
        kind = "port"
        what = "C restatement oracle/cup_oracle.c, %d OpenMP threads" % threads
    cells = n * 512
    print(json.dumps({"cell_updates_per_s": cells * a.steps / sec, "ms_per_cycle": 1e3 * sec / a.steps,
                      "threads": threads, "kind": kind, "what": what, "cells": cells}))


if __name__ == "__main__":
    main()
