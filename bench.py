#!/usr/bin/env python3
"""bench.py -- headline benchmark of the CUP3D hot path on B200.

Metric (BASELINE.json): cell-updates/s of the geometric-multigrid Poisson
V-cycle on a 512^3 uniform grid (bpd 1, levelStart 6, levelMax 7; 262 144
blocks of 8^3), fp64, zero right-hand side + a +1/-1 point-source pair.  One
"step" = one complete mg_vcycle over the whole grid; one cell-update = one
finest-level cell processed by one V-cycle.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

ours      : the CUDA library through its C ABI.  `value` is measured with the
            vectors resident in HBM (CUDA events on the library's stream);
            `e2e` calls the host-pointer entry cup_mg_vcycle() with pinned
            host buffers, H2D + V-cycle + D2H inside the timed region.
reference : the reference's own CPU implementation (oracle/_ref, the
            unmodified main.c + single-rank MPI shim, OpenMP over all host
            cores) timed on a bounded sample of the same workload.

Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

B_PER_CELL_VCYCLE = 171.0  # algorithmic bytes per cell-update, fp64 (SURVEY.md 8d / DESIGN.md)
B_PER_CELL_SMOOTH = 24.0   # smoother: read u, f, write u' (3 Reals)


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic():
    """per-launch DRAM bytes of the dominant kernel from the committed ncu capture, if any"""
    p = os.path.join(ROOT, "profiles", "smooth_traffic.json")
    try:
        with open(p) as f:
            return json.load(f)
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows = []
        self.gpu = gpu_index
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([s.strip() for s in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for k, nm in enumerate(names):
                    if r[5 + k].lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def point_sources(ib, rb):
    """b = +1 at cell 0 of the block containing (.25,.25,.25), -1 at (.75,.75,.75)"""
    import numpy as np
    out = []
    for p in (0.25, 0.75):
        lo = rb[:, 1:4]
        hi = lo + 8 * rb[:, 0:1]
        hit = np.all((lo <= p) & (p < hi), axis=1)
        out.append(int(np.nonzero(hit)[0][0]))
    return out


def free_ram_gb():
    try:
        for line in open("/proc/meminfo"):
            if line.startswith("MemAvailable"):
                return int(line.split()[1]) / 1e6
    except Exception:
        pass
    return 0.0


def cpu_reference_run(level, warmup, steps, timeout=1500, dump=None):
    """time the reference's CPU V-cycle in a subprocess (its state is process-global); dump: also write the
    V-cycle output of the point-source right-hand side to that .npy (the parity check reads it)"""
    cmd = [sys.executable, os.path.join(ROOT, "oracle", "cpu_baseline.py"), "--level", str(level), "--warmup",
           str(warmup), "--steps", str(steps)]
    if dump:
        cmd += ["--dump", dump]
    env = dict(os.environ)
    # torchrun exports OMP_NUM_THREADS=1 to every worker when the user has not set it; the CPU arm must
    # use all the host threads it can (it picks the best count itself), so that default is dropped.
    # CUP_REF_THREADS pins a count explicitly.
    if "TORCHELASTIC_RUN_ID" in env and env.get("OMP_NUM_THREADS") == "1":
        env.pop("OMP_NUM_THREADS")
    if env.get("CUP_REF_THREADS"):
        env["OMP_NUM_THREADS"] = env["CUP_REF_THREADS"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=timeout, env=env)
    for line in r.stdout.splitlines()[::-1]:
        if line.startswith("{"):
            return json.loads(line)
    raise RuntimeError("cpu baseline failed: " + r.stderr[-2000:])


def run_reference(args, rank, world):
    if rank != 0:
        return
    # the configuration itself (512^3: ~20 GB of host RAM, about a minute of the reference's mesh_init) when the
    # host has the memory, else a bounded 256^3 sample of the same workload; --cpu-level forces one
    lvl = args.cpu_level if args.cpu_level >= 0 else (args.level if free_ram_gb() >= 48 else 5)
    steps = min(args.steps, 10) if lvl >= 6 else args.steps
    res = cpu_reference_run(lvl, min(args.warmup, 2), steps)
    args = argparse.Namespace(**dict(vars(args), steps=steps))
    cells = (8 << lvl) ** 3
    line = {
        "impl": "reference", "metric": "poisson_vcycle_cell_updates_per_s", "value": res["cell_updates_per_s"],
        "unit": "cell-updates/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": res["ms_per_cycle"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "512^3 uniform Poisson V-cycle (bpd 1, levelStart 6, levelMax 7), fp64, point-source "
                               "pair; CPU arm timed on %s" % ("the full 512^3 grid" if lvl >= 6 else
                                                              "a bounded %d^3 sample of the same workload" % (8 << lvl)),
                   "cpu_grid": 8 << lvl, "same_config": lvl >= 6},
        "cpu_baseline": {"value": res["cell_updates_per_s"], "unit": "cell-updates/s", "cores": res["threads"],
                         "kind": res["kind"], "sample": "%d V-cycles of the %d^3 grid (%d cells), %s" %
                         (args.steps, 8 << lvl, cells, res["what"])},
        "e2e": {"value": res["cell_updates_per_s"], "unit": "cell-updates/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def sweep_rates(ctx, torch, capi, N, b, z, stream):
    """ms per sweep (CUDA events on the library's stream, 5 launches after 2 warm-ups) and GB/s at the
    algorithmic traffic of SURVEY 8(d); the state is filled with uniform random numbers on the device"""
    import ctypes
    rt = ctypes.cdll.LoadLibrary("libcudart.so.12")
    rng = torch.Generator(device="cuda").manual_seed(1)
    t = torch.empty(N, dtype=torch.float64, device="cuda")
    for f in range(9):
        t.uniform_(-1, 1, generator=rng)
        torch.cuda.synchronize()
        rc = rt.cudaMemcpy(ctypes.c_void_p(ctx.state_dev(f)), ctypes.c_void_p(t.data_ptr()), ctypes.c_size_t(N * 8), 3)
        if rc != 0:
            raise RuntimeError("cudaMemcpy into the state failed: %d" % rc)
    del t
    ctx.set_params(dt=1e-3, nu=1e-3, uinf=(0.1, 0.0, 0.0), step=5)
    reals = {"advdiff": 9, "prhs": 8, "divp": 2, "gradp": 4, "pois_op": 2}
    calls = {"advdiff": lambda: ctx.stencil_apply(capi.ST_ADVDIFF), "prhs": lambda: ctx.stencil_apply(capi.ST_PRHS),
             "divp": lambda: ctx.stencil_apply(capi.ST_DIVP), "gradp": lambda: ctx.stencil_apply(capi.ST_GRADP),
             "pois_op": lambda: ctx.pois_op_dev(b, z)}
    out = {}
    for name, fn in calls.items():
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(5):
            fn()
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        out[name] = {"ms": round(ms, 4), "Gcells_per_s": round(N / ms / 1e6, 2),
                     "GBs_at_algorithmic": round(N * 8 * reals[name] / ms / 1e6, 1)}
    return out


def parity_vs_reference(ctx, torch, dist, capi, mesh, args, rank, world):
    """The SAME point-source V-cycle on the grid of the CPU sample (256^3), on all N ranks, against the
    reference run live on the host (oracle/_ref in a subprocess on rank 0): max |z - z_ref| / max |z_ref|.
    This is the driver-visible correctness evidence for every N (decomposition may only change rounding).
    -> (parity dict, cpu_baseline dict or None)"""
    import tempfile
    import numpy as np
    Lp = args.cpu_level if args.cpu_level >= 0 else 5
    gib, grb = mesh.uniform_blocks(Lp)
    owner = capi.split_owner(len(gib), world)
    mine = np.nonzero(owner == rank)[0]
    cpu, zref, err = None, None, None
    if rank == 0:
        d = tempfile.mkdtemp(prefix="cup_bench_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
        path = os.path.join(d, "zref.npy")
        try:
            full = world == 1 and not args.no_cpu
            r = cpu_reference_run(Lp, 1, 3 if full else 1, dump=path)
            zref = np.load(path)
            if full:
                cpu = {"value": r["cell_updates_per_s"], "unit": "cell-updates/s", "cores": r["threads"],
                       "kind": r["kind"],
                       "sample": "3 V-cycles of a %d^3 grid (same block-structured workload, bounded), %s" %
                                 (8 << Lp, r["what"])}
        except Exception as ex:
            err = str(ex)[:300]
        finally:
            import shutil
            shutil.rmtree(d, ignore_errors=True)
    ok = torch.tensor([0 if (rank == 0 and zref is None) else 1], device="cuda")
    if dist is not None:
        dist.broadcast(ok, src=0)
    if int(ok.item()) == 0:
        return {"rel_err": None, "error": err or "reference run failed"}, cpu
    zr = torch.empty(len(gib) * 512, dtype=torch.float64, device="cuda")
    if rank == 0:
        zr.copy_(torch.from_numpy(zref.reshape(-1)))
    if dist is not None:
        dist.broadcast(zr, src=0)
    ctx.mesh_upload(gib[mine], grb[mine], (1, 1, 1), Lp + 1)
    ctx.set_params(mean_constraint=2)
    n = len(mine)
    b = torch.zeros(n * 512, dtype=torch.float64, device="cuda")
    for g, val in zip(point_sources(gib, grb), (1.0, -1.0)):
        if owner[g] == rank:
            b[int(g - mine[0]) * 512] = val
    z = torch.empty_like(b)
    for _ in range(2):  # the second call replays the captured CUDA graph: both paths must agree with the reference
        ctx.mg_vcycle_dev(b, z)
        torch.cuda.synchronize()
        ctx.synchronize()
    lo = int(mine[0]) * 512
    diff = (z - zr[lo:lo + n * 512]).abs().max().reshape(1)
    if dist is not None:
        dist.all_reduce(diff, op=dist.ReduceOp.MAX)
    ref_max = float(zr.abs().max().item())
    rel = float(diff.item()) / ref_max
    par = {"rel_err": rel, "grid": 8 << Lp, "tolerance": 1e-10, "ref_max": ref_max,
           "against": "oracle/_ref (unmodified reference main.c) run live on the host, same RHS"}
    # secondary: the full Krylov solve on the same grid (pois_solve, main.c:4875; the reference's 62 %-of-step
    # function): cosine right-hand side to 1e-10, and how much of it is spent outside the V-cycles
    try:
        import ctypes
        rt = ctypes.cdll.LoadLibrary("libcudart.so.12")
        X, Y, Z = None, None, None
        rhs = torch.empty(n * 512, dtype=torch.float64, device="cuda")
        step = 2048
        for s0 in range(0, n, step):
            X, Y, Z = mesh.cell_centers(gib[mine][s0:s0 + step], grb[mine][s0:s0 + step])
            h = grb[mine][s0:s0 + step, 0][:, None, None, None]
            blk = (h ** 3 * np.cos(np.pi * X) * np.cos(2 * np.pi * Y) * np.cos(3 * np.pi * Z)).reshape(-1)
            rhs[s0 * 512:s0 * 512 + blk.size] = torch.from_numpy(blk).cuda()
        ctx.set_params(mean_constraint=2, ptol=1e-10, ptol_rel=1e-14)
        best = None
        for _ in range(2):
            torch.cuda.synchronize()
            rt.cudaMemcpy(ctypes.c_void_p(ctx.state_dev(capi.F_LHS)), ctypes.c_void_p(rhs.data_ptr()),
                          ctypes.c_size_t(n * 512 * 8), 3)
            rt.cudaMemset(ctypes.c_void_p(ctx.state_dev(capi.F_PRES)), 0, ctypes.c_size_t(n * 512 * 8))
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            info = ctx.pois_solve()  # synchronous
            dt_s = time.perf_counter() - t0
            best = dt_s if best is None else min(best, dt_s)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(5):
            ctx.mg_vcycle_dev(b, z)
        e1.record()
        torch.cuda.synchronize()
        vc_ms = e0.elapsed_time(e1) / 5
        par["pois_solve"] = {"grid": 8 << Lp, "rhs": "h^3 cos(pi x) cos(2 pi y) cos(3 pi z)", "tolerance": 1e-10,
                             "ms": best * 1e3, "iterations": info.iterations, "vcycles": info.vcycles,
                             "residual": info.residual, "ms_per_vcycle": vc_ms,
                             "frac_outside_vcycles": 1.0 - info.vcycles * vc_ms / (best * 1e3),
                             "reference_cpu_s_survey_8_threads": 5.97}
    except Exception as ex:
        par["pois_solve"] = {"error": str(ex)[:200]}
    return par, cpu


def run_ours(args, rank, world, local_rank):
    import numpy as np
    import torch
    import cup3d_b200
    from cup3d_b200 import mesh

    from cup3d_b200 import capi
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    L = args.level
    gib, grb = mesh.uniform_blocks(L)
    # strong scaling: the SAME 512^3 grid, contiguous ranges of the Hilbert-ordered block list
    # per rank (the reference's mesh_init split, main.c:3306-3322)
    owner = capi.split_owner(len(gib), world)
    mine = np.nonzero(owner == rank)[0]
    ib, rb = gib[mine], grb[mine]
    ctx = cup3d_b200.Context(local_rank, 8)
    if world > 1:
        box = [capi.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        ctx.comm_init(rank, world, box[0])
    parity, cpu = ({"rel_err": None, "error": "skipped (--no-parity)"}, None)
    if not args.no_parity:
        parity, cpu = parity_vs_reference(ctx, torch, dist, capi, mesh, args, rank, world)
    ctx.mesh_upload(ib, rb, (1, 1, 1), L + 1)
    ctx.set_params(mean_constraint=2)
    n = len(ib)
    N = n * 512
    gcells = len(gib) * 512
    gsrc = point_sources(gib, grb)
    b = torch.zeros(N, dtype=torch.float64, device="cuda")
    src = []
    for g, val in zip(gsrc, (1.0, -1.0)):
        if owner[g] == rank:
            src.append((int(g - mine[0]), val))
            b[int(g - mine[0]) * 512] = val
    z = torch.empty_like(b)
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(v):
        if dist is None:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # clocks / throttle reasons are sampled from the warm-up to the end of the kernel timing so
    # that even the sub-millisecond multi-GPU steps are covered by several samples
    clk = ClockSampler(local_rank)
    clk.start()
    for _ in range(args.warmup):
        ctx.mg_vcycle_dev(b, z)
    barrier()
    l0 = ctx.kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        ctx.mg_vcycle_dev(b, z)
    e1.record(stream)
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))  # device time, max over ranks
    launches = ctx.kernel_launches() - l0
    if os.environ.get("CUP_STAMP"):
        # per-phase device timestamps of the last (graph-replayed) cycle, every rank (diagnostics)
        rep = ctx.trace_report()
        tot = sum(v for _, v in rep)
        sys.stderr.write("[stamp rank %d] total %.1f us (incl. ~%d stamp kernels): %s\n" %
                         (rank, tot / 1e3, len(rep), " ".join("%s=%.1f" % (k, v / 1e3) for k, v in rep)))
    # dominant kernel: finest-level smoother, timed live with CUDA events on the same stream
    # (single-rank kernel time; ghost faces of other ranks are whatever the last exchange left)
    sm_ms = ctx.time_smooth(L, 40)
    clocks = clk.stop()
    checksum = float(z.abs().sum().item())
    # result fingerprint, comparable across N by the driver: pois_dot(z, z) over all ranks and four fixed cells
    zz = ctx.pois_dot_dev(z, z)
    G = len(gib)
    samp = torch.zeros(4, dtype=torch.float64, device="cuda")
    for k, g in enumerate((0, G // 3, G // 2 + 17, G - 1)):
        if owner[g] == rank:
            samp[k] = z[int(g - mine[0]) * 512 + 77]
    if dist is not None:
        dist.all_reduce(samp)
    fingerprint = {"pois_dot_zz": zz, "cells": [float(v) for v in samp.tolist()]}

    # end to end through the host-pointer C ABI: pinned host in/out, H2D + V-cycle + D2H per step
    hb = torch.zeros(N, dtype=torch.float64).pin_memory()
    for i, val in src:
        hb[i * 512] = val
    hz = torch.empty(N, dtype=torch.float64).pin_memory()
    e2e_steps = max(3, min(args.steps, 5))
    ctx.mg_vcycle(hb, hz)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        ctx.mg_vcycle(hb, hz)  # synchronous: returns after the D2H
    barrier()
    t_e2e = max_over_ranks((time.perf_counter() - t0) / e2e_steps)
    assert abs(float(hz.abs().sum().item()) - checksum) <= 1e-9 * max(checksum, 1e-300)
    if rank != 0:
        ctx.close()
        if dist is not None:
            dist.destroy_process_group()
        return

    cells = gcells
    value = cells * args.steps / (ms * 1e-3)
    peak, peak_src = measured_peak()
    smooth_gbs = N * B_PER_CELL_SMOOTH / (sm_ms * 1e-3) / 1e9
    traffic = ncu_traffic()
    # secondary numbers SURVEY 8(d) asks for: per-sweep rates of the time-step stencils and of pois_op on
    # the same 512^3 grid (after everything above was measured; a failure here cannot touch the line)
    sweeps = None
    if world == 1 and not args.no_sweeps:
        try:
            sweeps = sweep_rates(ctx, torch, capi, N, b, z, stream)
        except Exception as ex:
            sweeps = {"error": str(ex)[:200]}
    line = {
        "metric": "poisson_vcycle_cell_updates_per_s", "value": value, "unit": "cell-updates/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "%d^3 uniform Poisson V-cycle (bpd 1, levelStart %d, levelMax %d), fp64, zero RHS + "
                               "point-source pair" % (8 << L, L, L + 1),
                   "blocks": len(gib), "mg_levels": L + 1,
                   "l2_policy": "inputs larger than L2 (%.2f GB per vector per rank)" % (N * 8 / 1e9),
                   "parallelism": "%d rank(s), one per GPU, contiguous Hilbert ranges of the block list; ghost faces "
                                  "pushed into peer windows over NVLink by the sweep kernels (NCCL for setup and "
                                  "allreduce)" % world},
        "hbm_gbs_vcycle_per_gpu": cells * B_PER_CELL_VCYCLE * args.steps / (ms * 1e-3) / 1e9 / world,
        "roofline": {"bound": "hbm", "kernel": "k_smooth_tma<double> (finest-level smoother, this rank's blocks)",
                     "achieved": smooth_gbs, "peak": peak, "unit": "GB/s", "frac": smooth_gbs / peak,
                     "traffic": (traffic["bytes_per_launch"] * (N / float(traffic.get("cells_per_launch", 512 ** 3)))
                                 if traffic else None),
                     "traffic_source": ("ncu --set full capture at N=1 (profiles/), scaled by this rank's share of "
                                        "the cells" if traffic else None),
                     "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": N * B_PER_CELL_SMOOTH, "ms_per_launch": sm_ms,
                     "vcycle_frac_per_gpu_at_171B_per_cell":
                         cells * B_PER_CELL_VCYCLE * args.steps / (ms * 1e-3) / 1e9 / (peak * world)},
        "parity": parity,
        "fingerprint": fingerprint,
        "cpu_baseline": cpu,
        "e2e": {"value": cells / t_e2e, "unit": "cell-updates/s", "h2d_bytes_per_step": gcells * 8,
                "d2h_bytes_per_step": gcells * 8, "ms_per_step": t_e2e * 1e3},
        "gpu_launches": launches,
        "clocks": clocks,
        "sweeps": sweeps,
    }
    try:
        ctx.close()
    except Exception:
        pass
    # secondary, N = 1: BASELINE.json configs[2] (AMR full time step, fp32) measured by this same script in a
    # fresh process (bench.py --config amr), after everything above; a failure there cannot touch the line
    if world == 1 and not args.no_amr:
        try:
            torch.cuda.empty_cache()
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--config", "amr", "--steps", "3",
                                "--warmup", "2"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
            got = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
            line["amr_step"] = json.loads(got[-1]) if got else {"error": r.stderr[-300:]}
        except Exception as ex:
            line["amr_step"] = {"error": str(ex)[:200]}
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()
    if parity.get("rel_err") is not None and parity["rel_err"] > 1e-10:
        sys.stderr.write("bench.py: PARITY FAILURE: V-cycle differs from the reference by %.3e (> 1e-10)\n" %
                         parity["rel_err"])
        sys.exit(3)


def run_amr(args, rank, world, local_rank):
    """BASELINE.json configs[2]: 256^3 base (bpd 32, level 0) + 3 refinement levels, one full time step =
    advdiff() (3 k_advdiff sweeps + RK3 updates, main.c:5027) + projection() (k_prhs, k_divp, GMRES with the
    V-cycle preconditioner, k_gradp, updates, main.c:5828), fp32.  The static multi-level mesh is synthetic:
    every block that meets a spherical shell (radius 0.2 around the centre) is refined three times and the
    reference's 2:1 balance rule enforced (cup3d_b200/mesh.py:amr_blocks).  Under torchrun the Hilbert-ordered
    leaf list is split into contiguous ranges (coarse-fine interfaces cross ranks: ghost blocks)."""
    import numpy as np
    import torch
    import cup3d_b200
    from cup3d_b200 import capi, mesh
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    base = args.amr_base  # blocks per dimension of the base grid (32 -> 256^3 cells)
    nlev = args.amr_levels
    gib, grb = mesh.amr_blocks(0, nlev, mesh.sphere_shell((0.5, 0.5, 0.5), 0.2, 0.5), bpd=(base,) * 3)
    owner = capi.split_owner(len(gib), world)
    mine = np.nonzero(owner == rank)[0]
    ib, rb = gib[mine], grb[mine]
    n = len(ib)
    rbytes = 4 if args.amr_dtype == "f32" else 8
    ctx = cup3d_b200.Context(local_rank, rbytes)
    if world > 1:
        box = [capi.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        ctx.comm_init(rank, world, box[0])
    t0 = time.perf_counter()
    ctx.mesh_upload(ib, rb, (base,) * 3, nlev + 1)
    setup_s = time.perf_counter() - t0
    st = torch.zeros((n, 9, 512), dtype=torch.float64).pin_memory()
    stn = st.numpy()
    for s0 in range(0, n, 8192):
        X, Y, Z = mesh.cell_centers(ib[s0:s0 + 8192], rb[s0:s0 + 8192])
        m = len(X)
        stn[s0:s0 + m, 2] = (np.sin(2 * np.pi * X) * np.cos(2 * np.pi * Y)).reshape(m, 512)
        stn[s0:s0 + m, 3] = (-np.cos(2 * np.pi * X) * np.sin(2 * np.pi * Y)).reshape(m, 512)
        stn[s0:s0 + m, 4] = (0.1 * np.sin(2 * np.pi * Z)).reshape(m, 512)
    hmin = float(grb[:, 0].min())
    ptol, ptol_rel = (1e-4, 1e-3) if rbytes == 4 else (1e-6, 1e-4)
    nu, cfl = 1e-3, 0.4
    ctx.set_params(dt=1e-6, nu=nu, uinf=(0.0, 0.0, 0.0), step=5, mean_constraint=2, ptol=ptol, ptol_rel=ptol_rel)
    ctx.state_h2d(stn)
    # the reference's time-step control (sta_dt, main.c:5941-5961) for this velocity field, then fixed
    um = ctx.umax()
    dt = min(cfl * hmin / (um + 1e-8), (1.0 / 6.0) * hmin * hmin / (nu + (1.0 / 6.0) * hmin * um))
    ctx.set_params(dt=dt)
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    def maxr(v):
        if dist is None:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def step():
        ctx.advdiff()
        return ctx.projection()

    clk = ClockSampler(local_rank)
    clk.start()
    for _ in range(args.warmup):
        step()
    barrier()
    l0 = ctx.kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    its, vcs = 0, 0
    its_list = []  # a step that runs into the solver's 1000-iteration cap (fp32 residual floor) shows up here
    marks = []  # (before advdiff, between, after projection) of every step: the two halves of the step
    e0.record(stream)
    for _ in range(args.steps):
        m = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        m[0].record(stream)
        ctx.advdiff()
        m[1].record(stream)
        info = ctx.projection()
        m[2].record(stream)
        marks.append(m)
        its += info.iterations
        vcs += info.vcycles
        its_list.append(int(info.iterations))
    e1.record(stream)
    barrier()
    ms = maxr(e0.elapsed_time(e1))
    launches = ctx.kernel_launches() - l0
    phase_ms = {"advdiff": maxr(sum(m[0].elapsed_time(m[1]) for m in marks) / len(marks)),
                "projection": maxr(sum(m[1].elapsed_time(m[2]) for m in marks) / len(marks))}
    clocks = clk.stop()
    umax = ctx.umax()
    # end to end: the velocity goes up from pinned host memory every step and the new velocity comes back
    # (the velocity that goes up is the one the previous step brought down, so the state stays consistent)
    ctx.state_d2h(stn, 2, 3)
    barrier()
    t0 = time.perf_counter()
    e2e_steps = 3
    for _ in range(e2e_steps):
        ctx.state_h2d(stn, 2, 3)
        step()
        ctx.state_d2h(stn, 2, 3)
    barrier()
    t_e2e = maxr((time.perf_counter() - t0) / e2e_steps)
    # algorithmic traffic: V-cycle = sum_L nact_L * 512 * 17 Reals + 2 N; k_advdiff stage 9 + RK update 12
    nact = torch.tensor([float(ctx.mg_nact(L)) for L in range(nlev + 1)], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(nact)
    gcells = len(gib) * 512
    vc_bytes = (float(nact.sum().item()) * 512 * 17 + 2 * gcells) * rbytes
    adv_bytes = 3 * 21 * gcells * rbytes
    if rank != 0:
        ctx.close()
        if dist is not None:
            dist.destroy_process_group()
        return
    peak, peak_src = measured_peak()
    value = gcells * args.steps / (ms * 1e-3)
    line = {
        "metric": "amr_time_step_cell_updates_per_s", "value": value, "unit": "cell-updates/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": args.amr_dtype, "data": "synthetic",
        "config": {"workload": "%d^3 base + %d AMR levels (static spherical-shell refinement, 2:1 balanced), one "
                               "advdiff() + projection() per step, %s" % (8 * base, nlev, args.amr_dtype),
                   "blocks": len(gib), "cells": gcells, "blocks_per_level": np.bincount(gib[:, 0]).tolist(),
                   "mg_active_blocks_per_level": [int(v) for v in nact.tolist()],
                   "poisson_tolerance": [ptol, ptol_rel], "dt": dt, "nu": nu, "cfl": cfl,
                   "l2_policy": "inputs larger than L2 (%.2f GB per field per rank)" % (n * 512 * rbytes / 1e9),
                   "parallelism": "%d rank(s); coarse-fine interfaces across ranks through ghost blocks" % world},
        "krylov_iterations_per_step": its / args.steps, "vcycles_per_step": vcs / args.steps,
        "krylov_iterations": its_list,
        "phases_ms": {k: round(v, 3) for k, v in phase_ms.items()},
        "roofline": {"bound": "hbm", "kernel": "whole step, algorithmic traffic of its two dominant parts",
                     "vcycle_bytes": vc_bytes, "advdiff_bytes": adv_bytes,
                     "advdiff_gbs_per_gpu": adv_bytes / (phase_ms["advdiff"] * 1e-3) / 1e9 / world,
                     "vcycles_gbs_per_gpu_if_all_projection_time": vc_bytes * (vcs / args.steps) /
                     (phase_ms["projection"] * 1e-3) / 1e9 / world,
                     "peak": peak, "unit": "GB/s", "peak_source": peak_src},
        "e2e": {"value": gcells / t_e2e, "unit": "cell-updates/s", "h2d_bytes_per_step": len(gib) * 3 * 512 * 8,
                "d2h_bytes_per_step": len(gib) * 3 * 512 * 8, "ms_per_step": t_e2e * 1e3},
        "gpu_launches": launches, "clocks": clocks, "setup_s": round(setup_s, 2), "umax": umax,
    }
    print(json.dumps(line), flush=True)
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--level", type=int, default=6, help="uniform level: grid = (8<<level)^3; 6 = 512^3")
    ap.add_argument("--cpu-level", type=int, default=-1,
                    help="grid of the CPU runs: 5 = 256^3, 6 = 512^3 (default: 5 for the parity / cpu_baseline leg of "
                         "ours; for --impl reference 6 when the host has >= 48 GB free, else 5)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-parity", action="store_true", help="skip the live parity check against the reference")
    ap.add_argument("--no-amr", action="store_true", help="skip the secondary AMR time-step measurement (N = 1)")
    ap.add_argument("--config", default="vcycle", choices=["vcycle", "amr"],
                    help="vcycle: BASELINE.json's headline (512^3 V-cycle); amr: configs[2], the full AMR time step")
    ap.add_argument("--amr-base", type=int, default=32, help="base blocks per dimension (32 = 256^3 cells)")
    ap.add_argument("--amr-levels", type=int, default=3, help="refinement levels above the base")
    ap.add_argument("--amr-dtype", default="f32", choices=["f32", "f64"])
    ap.add_argument("--no-sweeps", action="store_true", help="skip the per-sweep secondary timings")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    elif args.config == "amr":
        run_amr(args, rank, world, local_rank)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
