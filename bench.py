#!/usr/bin/env python
"""bench.py -- throughput of the MG-CFD per-cycle solver loop on B200 (and of the reference CPU solver beside it).

    python bench.py --gpus N --steps K --warmup W [--workload c2|c1|c3|c3s|tiny] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A STEP is one multigrid V-cycle of main()'s loop (src/euler3d_cpu_double.cpp:371-694) over the whole mesh: for a
4-level mesh 6 smoothing visits (step factor, 3 Runge-Kutta stages of flux + boundary/wall flux + time_step, residual),
3 restrictions and 3 prolongations.  The unit of work is one internal-edge flux evaluation ("flux edge-update"):
a cycle performs  sum_l E_I(l) * RK(3) * visits(l)  of them (BASELINE.json metric).

  value  = edge-updates of all ranks / device time of the K timed cycles (CUDA events on the solver's own stream,
           state resident in HBM, L2 flushed before every timed cycle, max over ranks)
  e2e    = the same metric through the public API with HOST buffers: every step copies level-0 `variables` from pinned
           host memory (set_field), runs one cycle (run_cycles -> RMS back), and reads `variables` back (get_field)
  roofline = the dominant kernel, the fused stage kernel on level 0 (k_stage_pipe: compute_flux_edge + boundary + wall flux +
           time_step in one launch per Runge-Kutta stage): algorithmic bytes per launch 32*E_I + 28*(E_B+E_W) + 128*N
           / its mean launch duration, measured with CUDA events in a second pass over the same K cycles (the three stage
           launches of a smoothing visit share one event pair, so that they overlap as in the replayed graph); peak =
           MEASURED_PEAKS.json hbm_gbs; traffic = DRAM bytes per launch from the ncu capture (profiles/traffic.json).  With the
           optional visit kernel (mgcfd_options.visit) one launch covers the three stages and is reported the same way.
  roofline_other = the same for the multigrid transfers on level 0/1 (prolong: 8*E_I + 148*N_f + 64*N_c; restrict: 44*N_f + 40*N_c)
  sustained = ms per step and SM clock over a >= 2 s back-to-back replay of the same cycle (the headline region is short enough to
           run at burst clocks)
  cpu_baseline = the UNMODIFIED reference (oracle/_ref/libmgcfd_ref_omp.so: its own sources built -DOMP -DOMP_SCATTERS)
           on the host cores, mesh duplicated once per thread as its assess-memory protocol does (gen_job.py:360-365), plus its
           serial build and the per-level flux rates from its own loop timers
  N > 1: "parity" = the same data plane on N ranks against one GPU on a small 4-level mesh before anything is timed (the run
           aborts when it fails); "north_star_c4" (and, on 8 GPUs, "north_star_c3" = the 64 M-node mesh) = the weak-scaling
           unit BASELINE.json's north star names, with the 1-GPU unit timed by rank 0 in the same run

`--impl reference` times that reference build alone (all host threads) and prints the same line with "impl": "reference"; its
timed sample is bounded to ~45 s of CPU work (MGCFD_REFERENCE_BUDGET_S) whatever --steps says -- the metric is a rate.
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

RK = 3
WORKLOADS = {
    # name: (generator kind, per-level node dims, mesh variant, description)
    "c2": (0, [[67] * 3, [55] * 3, [48] * 3, [43] * 3], 2,
           "C2: Onera-M6-shaped 4-level multigrid, hex-dual box 300763/166375/110592/79507 nodes, 888822 fine internal edges, mesh_name=m6wing"),
    "c1": (2, [[26, 25, 25]], 0, "C1: fvcorr.domn.097K-shaped single level, 97500 cell-centred tets, mesh_name=fvcorr"),
    "c3": (1, [[201] * 3, [101] * 3, [51] * 3, [26] * 3], 2, "C3: 8.1M-node Kuhn-tet box, 4 levels, 56.4M fine internal edges, mesh_name=m6wing"),
    "c4": (1, [[161] * 3, [81] * 3, [41] * 3, [21] * 3], 2, "C4 (weak scaling unit): 4.2M-node Kuhn-tet box per GPU, 4 levels (8 GPUs: 33M nodes; --workload c3 on 8 GPUs is the 64M-node mesh)"),
    "c3s": (1, [[129] * 3, [65] * 3, [33] * 3, [17] * 3], 2, "2.1M-node Kuhn-tet box, 4 levels (reduced C3)"),
    "tiny": (0, [[21, 19, 17], [11, 10, 9], [6, 5, 5]], 2, "tiny 3-level hex box (smoke)"),
    "parity": (0, [[40, 24, 22], [20, 12, 11], [10, 6, 6], [5, 3, 3]], 2, "small 4-level hex box (multi-GPU parity check)"),
}


def visits(level, levels):
    return 1 if (levels == 1 or level == 0 or level == levels - 1) else 2


def units_per_cycle(dims):
    nl = len(dims)
    return sum(d[1] * RK * visits(l, nl) for l, d in enumerate(dims))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons, pw = [], [], set(), []
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1])); pw.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": max(pw) if pw else None}


def measured_peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def host_threads():
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


def cpu_model_name():
    """What the reference's get_cpu_model_name reports (src/Base/common.h:114-143): the `model name` line of /proc/cpuinfo."""
    try:
        for line in open("/proc/cpuinfo"):
            if line.lower().startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def reference_cpu_run(workload, steps, warmup, threads, with_serial=False):
    """The unmodified reference on `threads` host threads over `threads` copies of the mesh (its OMP_SCATTERS protocol).
    Returns a dict: value (edge-updates/s aggregated over the copies), seconds per step, kind, cores, sample, steps timed,
    per-level flux rates from the reference's own loop timers and (with_serial) the serial build's rate."""
    os.environ["OMP_NUM_THREADS"] = str(threads)
    os.environ.setdefault("OMP_PROC_BIND", "spread")
    import mgcfd_b200 as M
    from oracle.loader import Oracle, Reference, reference_available
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from conftest import mesh_levels
    kind, dims, variant, desc = WORKLOADS[workload]
    mesh = M.Mesh.generate(kind, dims, mesh_variant=variant)
    ldims = [mesh.dims(l) for l in range(mesh.levels)]
    units = units_per_cycle(ldims)
    budget_s = float(os.environ.get("MGCFD_REFERENCE_BUDGET_S", "45"))
    build = "-O3 -fno-fast-math (-DPRECISE_FP: the oracle build; the reference's default -ffast-math build is ~13% faster, SURVEY.md 6), -march=x86-64-v3, without the indirect_rw probe the stock main() also runs"
    if reference_available(omp=True):
        ref = Reference(omp=True)
        t = ref.threads()
        sess = ref.session(variant, mesh_levels(mesh))
        if t > 1:
            sess.duplicate(t)
        sess.prepare()
        # bounded sample: a CPU V-cycle of C2 takes ~0.7 s on 16 threads, so at most `budget_s` seconds of cycles are timed (the
        # metric is a rate: it does not depend on how many cycles the sample holds); the warm-up is capped likewise
        _, _, first = sess.run(1)
        if warmup > 1:
            sess.run(min(warmup - 1, max(0, int(0.25 * budget_s / max(first, 1e-9)))))
        n = max(1, min(steps, int(budget_s / max(first, 1e-9))))
        _, _, secs = sess.run(n)
        tf = sess.times()["flux"]
        per_level = {f"L{l}": t * ldims[l][1] * RK * visits(l, len(ldims)) * n / tf[l] for l in range(len(ldims)) if tf[l] > 0}
        sess.close()
        out = dict(value=t * units * n / secs, s_per_step=secs / n, kind="reference", cores=t, steps_timed=n, cpu_model=cpu_model_name(),
                   flux_edge_updates_per_sec_by_level=per_level,
                   sample=f"{n} V-cycle(s) timed (of {steps} requested; bounded to ~{budget_s:.0f} s of CPU work, +warm-up) of {workload} duplicated x{t} "
                          f"(one copy per thread, -DOMP -DOMP_SCATTERS), oracle/_ref/libmgcfd_ref_omp.so = the reference's own sources, {build}; "
                          "per-level rates from its loop timers (flux<l>, src/Monitoring/timer.cpp:106-195)")
        if with_serial and reference_available(omp=False):
            sref = Reference(omp=False)
            ss = sref.session(variant, mesh_levels(mesh))
            ss.prepare()
            _, _, first = ss.run(1)
            ns = max(1, min(3, int(0.3 * budget_s / max(first, 1e-9))))
            _, _, ssecs = ss.run(ns)
            tfs = ss.times()["flux"]
            out["serial"] = {"value": units * ns / ssecs, "unit": "edge-updates/s", "cores": 1, "ms_per_step": ssecs / ns * 1e3, "steps_timed": ns,
                             "flux_edge_updates_per_sec_by_level": {f"L{l}": ldims[l][1] * RK * visits(l, len(ldims)) * ns / tfs[l] for l in range(len(ldims)) if tfs[l] > 0},
                             "build": "oracle/_ref/libmgcfd_ref.so (serial, same flags)"}
            ss.close()
        return out
    orc = Oracle()                      # scalar port, 1 thread
    lv = mesh_levels(mesh, apply_ewt_with=orc)
    t0 = time.perf_counter()
    orc.run_cycles(variant, lv, 1)
    first = time.perf_counter() - t0
    n = max(1, min(steps, int(budget_s / max(first, 1e-9))))
    t0 = time.perf_counter()
    orc.run_cycles(variant, lv, n)
    secs = time.perf_counter() - t0
    return dict(value=units * n / secs, s_per_step=secs / n, kind="port", cores=1, steps_timed=n, cpu_model=cpu_model_name(),
                sample=f"{n} V-cycle(s) timed (of {steps} requested; bounded to ~{budget_s:.0f} s) of {workload}, oracle/libmgcfd_oracle.so (scalar C port), 1 thread")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--flux-mode", type=int, default=None)
    ap.add_argument("--tile-nodes", type=int, default=None)
    ap.add_argument("--cpu-baseline-steps", type=int, default=2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-north-star", action="store_true", help="N > 1: skip the c4 / c3 weak-scaling units")
    ap.add_argument("--with-serial", action="store_true", help="(reference arm) also time the serial build")
    args = ap.parse_args()
    t_start = time.perf_counter()
    # NCCL prints "NCCL version ..." on STDOUT at NCCL_DEBUG=VERSION: keep stdout for the one JSON line
    if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", ""):
        os.environ["NCCL_DEBUG"] = "WARN"
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    kind, dims, variant, desc = WORKLOADS[args.workload]
    W = max(args.warmup, 0)

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return
        thr = host_threads()
        r = reference_cpu_run(args.workload, args.steps, W, thr, with_serial=args.with_serial)
        cb = {"value": r["value"], "unit": "edge-updates/s", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"], "cpu_model": r["cpu_model"]}
        for k in ("flux_edge_updates_per_sec_by_level", "serial"):
            if k in r:
                cb[k] = r[k]
        print(json.dumps({
            "impl": "reference", "metric": "flux edge-updates/s", "value": r["value"], "unit": "edge-updates/s", "n_gpus": args.gpus,
            "steps": args.steps, "steps_timed": r["steps_timed"], "warmup": W, "ms_per_step": r["s_per_step"] * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": desc, "step": "one V-cycle", "note": "CPU only; the mesh is duplicated once per host thread; the timed sample is bounded (cpu_baseline.sample)"},
            "cpu_baseline": cb,
            "e2e": {"value": r["value"], "unit": "edge-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}))
        return

    # ------------------------------------------------------------------ B200 arm
    import numpy as np
    import torch
    import mgcfd_b200 as M
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; mgcfd_b200 has no CPU fallback (use --impl reference for the CPU solver)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        # control plane (barriers, the max over ranks, the NCCL id broadcast) on gloo; the data plane -- halo rows and scalar
        # all-reduces -- is the library's own: peer-to-peer stores over NVLink from the kernels that produce the rows, or NCCL
        import torch.distributed as dist
        dist.init_process_group("gloo")

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    kw = {}
    if args.flux_mode is not None:
        kw["flux_mode"] = args.flux_mode
    if args.tile_nodes is not None:
        kw["tile_nodes"] = args.tile_nodes
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=f"cuda:{local}")      # > 126 MB of L2
    plane = {"text": "single GPU"}

    def grown(d, n, knd):
        # N > 1: weak scaling -- the box grows N-fold along x (N times the nodes and edges)
        return [[x[0] * n - (n - 1), x[1], x[2]] for x in d] if (n > 1 and knd != 2) else [[x[0] * n, x[1], x[2]] for x in d]

    def make_solver(workload, nranks):
        """this rank's solver for `workload` grown nranks-fold; (solver, global (nodes, internal edges) per level, nodes held on level 0)"""
        knd, dm, var, _ = WORKLOADS[workload]
        gd = grown(dm, nranks, knd)
        if nranks == 1:
            mesh = M.Mesh.generate(knd, gd, mesh_variant=var, lengths=(1.0, 1.0, 1.0))
            ld = [mesh.dims(l) for l in range(mesh.levels)]
            s = M.Solver.from_mesh(mesh, device=local, **kw)
            mesh.close()
            return s, ld, ld[0][0]
        idt = torch.zeros(128, dtype=torch.uint8)
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)                     # NCCL prints its version banner on stdout: keep stdout for the one JSON line
        try:
            if rank == 0:
                idt = torch.frombuffer(bytearray(M.dist_unique_id()), dtype=torch.uint8).clone()
            dist.broadcast(idt, 0)
            # every rank generates ITS part of the N-fold mesh (mgcfd_generate_upload_partition: bit for bit the partition of the
            # assembled mesh, tests/test_partition.py) -- no rank ever holds the global edge list, so 64 M nodes fit 8 ranks' hosts
            s = M.Solver.generate_distributed(knd, gd, rank, nranks, bytes(idt.numpy().tobytes()), mesh_variant=var,
                                              lengths=(float(nranks), 1.0, 1.0), device=local, **kw)
            ld = []
            for l in range(len(dm)):
                info = s.dist_level_info(l)
                ld.append((info["global_nodes"], info["global_internal_edges"]))          # GLOBAL counts
            if os.environ.get("MGCFD_NO_P2P", "0") != "1":      # direct peer-to-peer data path (CUDA IPC) instead of NCCL
                mine = s.p2p_prepare()
                allp = [None] * nranks
                dist.all_gather_object(allp, mine)
                ok = 1
                try:
                    s.p2p_attach([a[0] for a in allp], [a[1] for a in allp])
                except M.MgcfdError as e:          # no peer access between these GPUs: stay on NCCL (every rank must agree)
                    print(f"rank {rank}: peer-to-peer attach failed ({e}); using NCCL", file=sys.stderr)
                    ok = 0
                flag = torch.tensor([ok])
                dist.all_reduce(flag, op=dist.ReduceOp.MIN)
                if int(flag) == 0:
                    raise SystemExit("bench.py: the ranks disagree on the peer-to-peer data path; rerun with MGCFD_NO_P2P=1")
                plane["text"] = ("peer-to-peer stores over NVLink (one CUDA IPC slab per rank): every kernel delivers the rows it produces into the "
                                 "other ranks' arrays itself, the visit kernel's grid barriers double as the halo exchange; cycle replayed as a CUDA graph")
            else:
                plane["text"] = "NCCL send/recv + all-reduce between the stage kernels"
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)
        return s, ld, s._nel[0]

    def timed_pass(s, K):
        stream = torch.cuda.ExternalStream(s.cuda_stream(), device=local)
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
        l0 = s.launch_count()
        barrier()
        with torch.cuda.stream(stream):
            for a, b in ev:
                flush.zero_()
                a.record(stream)
                s.enqueue_cycles(1)
                b.record(stream)
        ra, _ = s.collect()
        barrier()
        return sum(a.elapsed_time(b) for a, b in ev), s.launch_count() - l0, ra

    def max_over_ranks(vals):
        if dist is None:
            return list(vals)
        t = torch.tensor(list(vals), dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.tolist()

    # ------------------------------------------------------------------ N > 1: parity of this data plane before anything is timed
    parity = None
    if world > 1:
        cycles = 6
        s, _, _ = make_solver("parity", world)
        ra, _ = s.run_cycles(cycles)
        pieces = []
        nl = len(WORKLOADS["parity"][1])
        for l in range(nl):
            info = s.dist_level_info(l)
            pieces.append((s.global_ids(l)[:info["owned"]], s.get_field(l, M.FIELD_VARIABLES)[:info["owned"]].copy()))
        gathered = [None] * world
        dist.all_gather_object(gathered, pieces)
        s.close()
        err = 0.0
        if rank == 0:
            knd, dm, var, _ = WORKLOADS["parity"]
            ref = M.Solver.from_mesh(M.Mesh.generate(knd, grown(dm, world, knd), mesh_variant=var, lengths=(float(world), 1.0, 1.0)), device=local, **kw)
            rra, _ = ref.run_cycles(cycles)
            err = float(np.max(np.abs(ra - rra) / rra))
            for l in range(nl):
                want = ref.get_field(l, M.FIELD_VARIABLES)
                got = np.full_like(want, np.nan)
                for p in gathered:
                    got[p[l][0]] = p[l][1]
                e = np.max(np.abs(got - want), axis=0) / np.max(np.abs(want), axis=0)
                err = max(err, float(np.max(e)) if np.all(np.isfinite(e)) else float("inf"))
            ref.close()
        (err,) = max_over_ranks([err])
        parity = {"max_rel_err": err, "tol": 1e-11, "nranks": world, "cycles": cycles,
                  "what": "rms history + final variables of every level, N ranks vs one GPU, " + WORKLOADS["parity"][3] + f" grown {world}-fold"}
        if not (err < 1e-11):
            if rank == 0:
                print(json.dumps({"error": "multi-GPU parity check failed", "parity": parity}))
            raise SystemExit(3)

    # ------------------------------------------------------------------ the headline unit
    t0 = time.perf_counter()
    s, ldims, n_local0 = make_solver(args.workload, world)
    units = units_per_cycle(ldims)
    setup_s = time.perf_counter() - t0
    K = args.steps
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()           # clocks are sampled (20 ms) from the warm-up to the end of the second timed pass
    s.run_cycles(W)
    ms_total, launches, rms = timed_pass(s, K)
    # second pass, same K cycles, every kernel bracketed by its own CUDA events (graphs bypassed): per-kernel durations
    s.set_timing(True)
    s.reset_times()
    s.run_cycles(2)
    s.reset_times()
    ms_total_timed, _, _ = timed_pass(s, K)
    t_ms, t_it = s.times()
    s.set_timing(False)
    clocks = sampler.stop() if rank == 0 else None
    # sustained: the same cycle replayed back to back for >= 2 s (no event pairs, no flush between cycles: the steady-state rate
    # at the clocks the device settles to)
    sus = None
    if world == 1:
        n_sus = int(min(4000, max(50, 2.2e3 / max(ms_total / K, 1e-3))))
        sampler2 = ClockSampler(local)
        sampler2.start()
        stream = torch.cuda.ExternalStream(s.cuda_stream(), device=local)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.run_cycles(5)
        with torch.cuda.stream(stream):
            e0.record(stream)
            s.enqueue_cycles(n_sus)
            e1.record(stream)
        s.collect()
        torch.cuda.synchronize()
        c2 = sampler2.stop()
        sus_ms = e0.elapsed_time(e1)
        sus = {"ms_per_step": sus_ms / n_sus, "steps": n_sus, "seconds": sus_ms * 1e-3, "value": units * n_sus / (sus_ms * 1e-3),
               "sm_mhz_median": c2.get("sm_mhz"), "reasons": c2.get("reasons"), "l2": "not flushed between cycles (back-to-back replay)"}

    # e2e: host buffers through the public API
    n0 = n_local0
    host_in = torch.empty(5 * n0, dtype=torch.float64).pin_memory()
    host_out = torch.empty(5 * n0, dtype=torch.float64).pin_memory()
    host_in.numpy()[:] = s.get_field(0, M.FIELD_VARIABLES).reshape(-1)
    e2e_steps = max(3, min(K, 50))
    for _ in range(2):
        s.set_field(0, M.FIELD_VARIABLES, host_in.numpy()); s.run_cycles(1); s.get_field(0, M.FIELD_VARIABLES, out=host_out.numpy())
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        s.set_field(0, M.FIELD_VARIABLES, host_in.numpy())
        s.run_cycles(1)
        s.get_field(0, M.FIELD_VARIABLES, out=host_out.numpy())
        host_in, host_out = host_out, host_in
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    ms_total, e2e_s, ms_total_timed = max_over_ranks([ms_total, e2e_s, ms_total_timed])

    nl = len(ldims)
    info0 = s.level_info(0)
    vinfo = [s.visit_info(l) for l in range(nl)]
    linfo = [s.level_info(l) for l in range(nl)]
    s.close()

    # ------------------------------------------------------------------ N > 1: the north-star weak-scaling units
    north = {}
    if world > 1 and not args.no_north_star:
        todo = ["c4"] + (["c3"] if world == 8 else [])
        for wl in todo:
            key = "north_star_" + wl
            if time.perf_counter() - t_start > (420 if wl == "c4" else 520):
                north[key] = {"skipped": "time budget of this run"}
                continue
            try:
                nk = 10 if wl == "c4" else 6
                one = None
                if rank == 0:                      # the 1-GPU unit, timed by rank 0 alone in this same run
                    s1, ld1, _ = make_solver(wl, 1)
                    s1.run_cycles(3)
                    stream = torch.cuda.ExternalStream(s1.cuda_stream(), device=local)
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    with torch.cuda.stream(stream):
                        e0.record(stream); s1.enqueue_cycles(nk); e1.record(stream)
                    s1.collect(); torch.cuda.synchronize()
                    one = (e0.elapsed_time(e1) / nk, units_per_cycle(ld1), ld1[0][0])
                    s1.close()
                sN, ldN, _ = make_solver(wl, world)
                sN.run_cycles(3)
                stream = torch.cuda.ExternalStream(sN.cuda_stream(), device=local)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                barrier()
                with torch.cuda.stream(stream):
                    e0.record(stream); sN.enqueue_cycles(nk); e1.record(stream)
                raN, _ = sN.collect(); barrier()
                (msN,) = max_over_ranks([e0.elapsed_time(e1) / nk])
                vis = [sN.visit_info(l)["visit"] for l in range(len(ldN))]
                sN.close()
                if rank == 0:
                    valN = units_per_cycle(ldN) / (msN * 1e-3)
                    val1 = one[1] / (one[0] * 1e-3)
                    north[key] = {"workload": WORKLOADS[wl][3], "nodes": int(ldN[0][0]), "internal_edges_level0": int(ldN[0][1]), "n_gpus": world, "steps": nk,
                                  "ms_per_step": msN, "value": valN, "unit": "edge-updates/s", "mg_cycles_per_sec": 1e3 / msN,
                                  "one_gpu_unit": {"nodes": int(one[2]), "ms_per_step": one[0], "value": val1},
                                  "efficiency_vs_same_run_1gpu_unit": valN / (world * val1), "final_rms": float(raN[-1]), "visit_kernel_levels": vis,
                                  "l2": "not flushed (working set far above L2)"}
            except Exception as e:          # the headline line is printed whatever happens here
                north[key] = {"error": repr(e)}
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    nel, nI, nB, nW = info0["nel"], info0["nI"], info0["nB"], info0["nW"]      # this rank's level 0 (the whole level when N = 1)
    flux_ms0, flux_it0 = float(t_ms[1, 0]), int(t_it[1, 0])
    peak, peak_src = measured_peak()
    use_visit = bool(vinfo[0]["visit"])
    stage_launches0 = flux_it0 // max(nI, 1)                 # RK stages timed on level 0
    stage_bytes = 32 * nI + 28 * (nB + nW) + 128 * nel
    if use_visit:
        launches_timed = stage_launches0 // RK                # one launch = one visit = three stages
        alg_bytes = RK * stage_bytes
        kname = ("k_visit level 0 (persistent: minimum dt + 3 x [compute_flux_edge + boundary + wall flux + time_step] + residual/RMS, ONE launch per "
                 "smoothing visit; algorithmic bytes = 3 stages x (32 E_I + 28 (E_B+E_W) + 128 N))")
    else:
        launches_timed = stage_launches0
        alg_bytes = stage_bytes
        kname = "k_stage_pipe level 0 (compute_flux_edge + boundary + wall flux + time_step fused, one launch per RK stage)"
    avg_launch_ms = flux_ms0 / max(launches_timed, 1)
    achieved = alg_bytes / (avg_launch_ms * 1e-3) / 1e9 if avg_launch_ms > 0 else 0.0
    traffic = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        traffic = tj.get(args.workload + ("_visit" if use_visit else ""), {}).get("dram_bytes_per_launch")
    except Exception:
        pass
    per_level = {}
    for l in range(nl):
        if t_ms[1, l] > 0:
            per_level[f"L{l}"] = float(t_it[1, l]) / (float(t_ms[1, l]) * 1e-3)
    kernel_share = {name: float(t_ms[k].sum()) for k, name in enumerate(M.KERNEL_NAMES) if t_ms[k].sum() > 0}
    other = {}
    if nl > 1 and world == 1:
        nf, nc = linfo[0]["nel"], linfo[1]["nel"]
        # prolong level 1 -> 0 (timer index 6, level 0) and restrict level 0 -> 1 (timer index 5, level 1); iters = nI resp. nel_fine per launch
        for name, k, lev, nbytes, per in (("k_prolong level 1->0 (prolong_residuals_interpolate_proper)", 6, 0, 8 * linfo[0]["nI"] + 148 * nf + 64 * nc, linfo[0]["nI"]),
                                          ("k_restrict level 0->1 (mg_restrict)", 5, 1, 44 * nf + 40 * nc, nf)):
            n_l = int(t_it[k, lev]) // max(per, 1)
            if n_l > 0 and t_ms[k, lev] > 0:
                us = float(t_ms[k, lev]) / n_l * 1e3
                other[name] = {"achieved": nbytes / (us * 1e-6) / 1e9, "unit": "GB/s", "frac": nbytes / (us * 1e-6) / 1e9 / peak, "algorithmic_bytes_per_launch": nbytes,
                               "avg_launch_us": us, "launches_timed": n_l}
    out = {
        "metric": "flux edge-updates/s", "value": units * K / (ms_total * 1e-3), "unit": "edge-updates/s", "n_gpus": world,
        "steps": K, "warmup": W, "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": desc, "step": "one V-cycle (euler3d_cpu_double.cpp:371-694), all levels",
                   "edge_updates_per_step": units, "parallelism": "single GPU" if world == 1 else f"{world} ranks, mesh split by recursive coordinate bisection (box {world}x longer in x): " + plane["text"],
                   "l2": "256 MiB buffer written before every timed cycle (L2 flush)", "flux_mode": {0: "tiled coloured scatter", 1: "tiled sorted segment", 2: "atomic"}[int(kw.get("flux_mode", 1))],
                   "visit_kernel": [{"level": l, "on": bool(v["visit"]), "supers_per_cta": int(v["supers_per_cta"]), "ctas": int(v["ctas"]), "ring_rounds": int(v["ring_rounds"]), "ring_entries": int(v["ring_entries"]),
                                     "resident": bool(v["resident"]), "smem_bytes": int(v["smem_bytes"])} for l, v in enumerate(vinfo)],
                   "pipelined": bool(info0["pipe_grid"]), "tile_nodes": int(info0["tile_nodes"]), "setup_s": round(setup_s, 2)},
        "mg_cycles_per_sec": K / (ms_total * 1e-3),
        "flux_edge_updates_per_sec_by_level": per_level,
        "kernel_ms_timed_pass": kernel_share, "ms_per_step_timed_pass": ms_total_timed / K,
        "final_rms": float(rms[-1]) if len(rms) else None,
        "roofline": {"bound": "hbm", "kernel": kname, "achieved": achieved,
                     "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": alg_bytes, "avg_launch_us": avg_launch_ms * 1e3, "launches_timed": launches_timed,
                     "edge_updates_per_sec": flux_it0 / (flux_ms0 * 1e-3) if flux_ms0 > 0 else 0.0},
        "roofline_other": other,
        "e2e": {"value": units * e2e_steps / e2e_s, "unit": "edge-updates/s", "h2d_bytes_per_step": 40 * n0, "d2h_bytes_per_step": 40 * n0 + 48,
                "steps": e2e_steps, "ms_per_step": e2e_s / e2e_steps * 1e3,
                "path": "Solver.set_field(pinned host) -> Solver.run_cycles(1) -> Solver.get_field(pinned host), C ABI mgcfd_set_field/mgcfd_run_cycles/mgcfd_get_field"},
        "gpu_launches": int(launches), "clocks": clocks,
    }
    if sus is not None:
        out["sustained"] = sus
    if parity is not None:
        out["parity"] = parity
    out.update(north)
    del flush
    if not args.no_cpu_baseline:
        try:
            # in a fresh process: the OpenMP runtime must see OMP_NUM_THREADS before it starts, and torch has started it here
            env = dict(os.environ)
            for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
                env.pop(k, None)
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--workload", args.workload, "--steps",
                                str(args.cpu_baseline_steps), "--warmup", "1", "--with-serial"], capture_output=True, text=True, env=env, timeout=900)
            line = [ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1]
            cb = json.loads(line)
            out["cpu_baseline"] = dict(cb["cpu_baseline"], ms_per_step=cb["ms_per_step"])
        except Exception as e:  # the baseline is reported, never required
            out["cpu_baseline"] = {"value": None, "unit": "edge-updates/s", "cores": 0, "kind": "unavailable", "sample": repr(e)}
    print(json.dumps(out))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
