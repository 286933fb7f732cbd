#!/usr/bin/env python
"""bench.py -- particle pushes/s (fused move+deposit) of the Starfish kinetic hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload auto|b|b_beam|c|e]

One "step" = one KineticMaterial.updateFields() pass (move + deposit + moment sums) over the whole resident
particle population; one "push" = one particle advanced one step including its deposit contribution.

Workloads (BASELINE.json configs):
  b  (default at N=1)  XY 512x512 nodes, 16,777,216 uniformly loaded particles, all faces periodic; with N>1 ranks
                        (--workload b) the same shard on every rank: weak scaling
  e  (default at N>1)  XY 2048x2048 nodes, 2^30 particles partitioned by index over the ranks (north_star's multi-GPU
                        config): STRONG scaling, per-step NCCL allreduce of the deposit
  c                     RZ 1024x1024, beam over r < 0.25 Rmax, LEFT symmetry, other faces open
Particle arrays (1 GiB at 16M) are far larger than the 126 MB L2, so no explicit L2 flush is needed.

JSON keys: see the task contract; `value` is device-resident throughput (CUDA events on the library's stream,
max over ranks), `e2e` is the same metric through the plugin API with host buffers for the fields (E upload,
nd/u/v/w + mover sums download every step; the running moment sums stay on the device), `roofline` is the fused step kernel against the measured HBM copy peak,
`cpu_baseline` the oracle restatement of the Java algorithm timed on this box's host cores.
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

import numpy as np  # noqa: E402

ALGO_BYTES_PER_PUSH = 104  # read pos[3],vel[3],mpw + write pos[3],vel[3], FP64


def workload(name, world):
    from starfish_b200 import synthetic as S
    if name == "b":
        return S.config_b(), 1 << 24
    if name == "b_beam":
        return S.config_b(beam=True), 1 << 24
    if name == "c":
        return S.config_c(), 1 << 27
    if name == "e":
        return S.config_e(), (1 << 30) // world
    raise ValueError(name)


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.proc = index, [], None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc:
            self.proc.terminate()
        self.join(timeout=2)
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference_run(wl, n_sample, steps, warmup, threads):
    """The reference's CPU algorithm (oracle port: T mover threads, serial deposit + sampling like KM:180/:1580)."""
    from oracle import oracle as O
    ok = O.OracleKM(wl.charge, wl.mass, [wl.mesh], threads=threads)
    ok.addParticles(0, wl.particles(0, n_sample), wl.dt)
    for _ in range(warmup):
        ok.updateFields(wl.dt)
    t0 = time.perf_counter()
    for _ in range(steps):
        ok.updateFields(wl.dt)
    t = time.perf_counter() - t0
    return n_sample * steps / t, t


def run_reference(args, rank, world, wname):
    if rank != 0:
        return
    wl, _n = workload(wname, world)
    threads = os.cpu_count() or 1
    n_sample = args.ref_particles
    value, t = cpu_reference_run(wl, n_sample, args.steps, args.warmup, threads)
    line = {
        "impl": "reference", "metric": "particle pushes/sec (move+deposit)", "value": value, "unit": "pushes/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "strong" if wname == "e" else "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl.name, "mesh_nodes": [wl.mesh.ni, wl.mesh.nj], "particles": n_sample,
                   "note": "bounded sample of the workload on the host cores; JVM unavailable, C restatement of the Java algorithm"},
        "cpu_baseline": {"value": value, "unit": "pushes/s", "cores": threads, "kind": "port",
                         "sample": f"{n_sample} particles x {args.steps} steps, {threads} mover threads + serial deposit (as KineticMaterial.java:180,:1580)"},
        "e2e": {"value": value, "unit": "pushes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


_RESULT_FD = None


def quiet_stdout():
    """stdout carries exactly ONE line, the result: file descriptor 1 is pointed at stderr for everything else (NCCL prints its version banner on stdout
    when NCCL_DEBUG is set in the environment, torch.distributed warns there too), the result line goes to the saved descriptor."""
    global _RESULT_FD
    if _RESULT_FD is None:
        sys.stdout.flush()
        _RESULT_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    sys.stdout.flush()
    if _RESULT_FD is None:
        os.write(1, data)
    else:
        os.write(_RESULT_FD, data)


def main():
    quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=0, help="timed steps (default: 340 for config B, 40 for the 2^30-particle config E)")
    ap.add_argument("--warmup", type=int, default=0)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="auto", choices=["auto", "b", "b_beam", "c", "e"])
    ap.add_argument("--particles", type=int, default=0, help="particles per rank (default: the config's)")
    ap.add_argument("--ref-particles", type=int, default=1 << 21)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--step-flags", type=int, default=0)
    ap.add_argument("--trace", action="store_true", help="stderr: per timed step, host ms of the injection and of the step, device ms of the step and its kernel")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    wname = args.workload if args.workload != "auto" else ("b" if world == 1 else "e")
    if args.impl == "reference":
        args.steps = args.steps or 3
        args.warmup = args.warmup or 1
        run_reference(args, rank, world, wname)
        return
    if args.steps <= 0:  # timed region >= 0.5 s
        args.steps = {"b": 340, "b_beam": 340, "c": 60, "e": 40}[wname]
    if args.warmup < 3:
        # config C injects every step: the store grows past its capacity hint once and the first sorts after that re-allocate their buffers
        # (several GB of cudaMalloc) -- two sort intervals of warm-up keep those one-time allocations out of the timed steps
        args.warmup = 6 if wname.startswith("b") else (7 if wname == "c" else 3)

    import torch
    import torch.distributed as dist
    from starfish_b200 import KineticMaterial, Particles
    from starfish_b200.parallel import attach_communicator

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: starfish_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    wl, n_rank = workload(wname, world)
    if args.particles:
        n_rank = args.particles
    m = wl.mesh
    km = KineticMaterial("O+", wl.charge, wl.mass, [m], m.domain_type, device=local_rank, capacity_hint=n_rank, step_flags=args.step_flags)
    km.dt = wl.dt
    # the host keeps its E field in page-locked memory (a Java host would use direct buffers from sfgpu_host_alloc)
    for name in ("efi", "efj"):
        pinned = km.hostArray((m.ni, m.nj))
        pinned[...] = getattr(m, name)
        setattr(m, name, pinned)
    attach_communicator(km)  # NCCL communicator of the library: rank 0 makes the id, torch.distributed carries it
    km.download_fields = rank == 0  # the deposit is identical on every rank after the allreduce: read back once
    # this rank's shard: particle indices [rank*n_rank, (rank+1)*n_rank) of the global population
    chunk = 1 << 23
    for first in range(0, n_rank, chunk):
        c = min(chunk, n_rank - first)
        arr = wl.particles(rank * n_rank + first, c)
        km.addParticles(m, Particles(c, **arr), wl.dt)
    n_local = km.getNp()

    def barrier():
        km.sync()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def allmax(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # config C: a beam leaves through the open faces and is re-injected at z = 0+ at the same rate (a pool of
    # pre-generated injection batches in host memory; uploading one is part of every step of this workload)
    inject = None
    if wname == "c":
        n_inj = max(1, int(round(n_rank * 0.5 / (m.nj - 1))))  # drift 0.5 cells/step over nj-1 cells
        pool = []
        for k in range(8):
            arr = wl.particles((world + rank) * n_rank + k * n_inj, n_inj)
            arr["y"] = m.x0[1] + (arr["y"] - m.x0[1]) * (0.5 / (m.nj - 1))  # inside the first half cell
            pool.append(Particles(n_inj, **arr))
        state = {"k": 0}

        def inject():
            km.addParticles(m, pool[state["k"] % len(pool)], wl.dt)
            state["k"] += 1

    def one_step():
        if args.trace:
            t_a = time.perf_counter()
        if inject:
            inject()
        if args.trace:
            km.sync()
            t_b = time.perf_counter()
        km.step_raw(wl.dt)
        if args.trace:
            km.sync()
            tot, ker, nl = km.lastStepTiming()
            print("trace: inject %.3f ms host, step %.3f ms host, %.3f ms device (kernel %.3f), %d launches, np %d" %
                  (1e3 * (t_b - t_a), 1e3 * (time.perf_counter() - t_b), tot, ker, nl, km.getNp()), file=sys.stderr)

    # ---- device-resident throughput ------------------------------------------------------------------
    for _ in range(args.warmup):
        one_step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    time.sleep(0.3)
    barrier()  # every rank enters the timed region together
    launches0 = km.launchCount()
    ker_ms, pushes, fallback = 0.0, 0, 0
    by_kind = {}  # step kernel -> [steps, kernel ms, pushes]
    t0 = time.perf_counter()
    km.timerStart()
    for _ in range(args.steps):
        one_step()
        np_alive, n_exit, ker, kind, nfall = km.stepStats()
        pushes += np_alive + n_exit
        ker_ms += ker
        rec = by_kind.setdefault(kind, [0, 0.0, 0])
        rec[0] += 1
        rec[1] += ker
        rec[2] += np_alive + n_exit
        fallback += nfall
    dev_ms = km.timerStop()
    barrier()
    wall_ms = 1e3 * (time.perf_counter() - t0)
    launches = km.launchCount() - launches0
    dev_ms = allmax(dev_ms)
    wall_ms = allmax(wall_ms)
    total_pushes = allsum(float(pushes))
    value = total_pushes / (dev_ms * 1e-3)

    # ---- end to end through the plugin API: E upload + step + deposit/moment download, host buffers ------
    for _ in range(2):
        if inject:
            inject()
        km.setFields(m)
        km.updateFields()
    barrier()
    e_pushes = 0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        if inject:
            inject()
        km.setFields(m)       # host -> device: efi, efj of this step (the Java solver's output)
        km.updateFields()
        e_pushes += km.getNp() + km.n_exited     # step + device -> host: nd/u/v/w and the mover sums (velocity-moment sums stay on the device)
    barrier()
    e2e_s = allmax(time.perf_counter() - t0)
    clocks = sampler.stop() if rank == 0 else None  # sampled over both timed regions
    e2e_value = allsum(float(e_pushes)) / e2e_s
    plane = m.ni * m.nj * 8
    h2d, d2h = 2 * plane, plane + 5 * 8  # per rank up: efi, efj; down (rank 0): nd + the mover sums (u, v, w on demand)

    # ---- multi-GPU consistency of the last step (every rank): counts, conservation, identical deposit everywhere ----
    multi_check = None
    if world > 1:
        import hashlib
        dep = km.last_deposit[0]
        np_total = int(allsum(float(km.getNp())))
        problems = []
        if int(dep[7].sum()) != np_total:
            problems.append("sum(mpc) %d != particles %d" % (int(dep[7].sum()), np_total))
        want = np_total * wl.mpw
        if abs(float(dep[0].sum()) - want) > 1e-10 * want:
            problems.append("sum(Den) %.17g != sum(mpw) %.17g" % (float(dep[0].sum()), want))
        digest = hashlib.sha1(dep.tobytes()).hexdigest()
        box = [None] * world
        dist.all_gather_object(box, digest)
        if len(set(box)) != 1:
            problems.append("deposit differs between ranks")
        allp = [None] * world
        dist.all_gather_object(allp, problems)
        flat = [q for pr in allp for q in pr]
        multi_check = "ok" if not flat else "FAILED: " + "; ".join(sorted(set(flat)))

    if rank == 0:
        peaks, peak_src = None, "fallback"
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
            peak = float(peaks["hbm_gbs"])
            peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            peak = 6650.0
        # roofline of the DOMINANT step kernel (the one that ran most of the timed steps); the other one is listed next to it
        names = {0: "k_fast_step (tiled in-place move+deposit)", 1: "streaming step (move+deposit+re-sort in one pass)", 2: "k_fast_tail (generic)"}
        dom = max(by_kind, key=lambda k: by_kind[k][0])
        d_steps, d_ms, d_pushes = by_kind[dom]
        achieved = ALGO_BYTES_PER_PUSH * float(d_pushes) / (d_ms * 1e-3) / 1e9 if d_ms > 0 else 0.0
        kernels = {names[k]: {"steps": v[0], "kernel_ms_per_step": v[1] / v[0], "GB/s": ALGO_BYTES_PER_PUSH * v[2] / (v[1] * 1e-3) / 1e9 if v[1] > 0 else 0.0}
                   for k, v in by_kind.items()}
        # dram bytes per launch of the dominant kernel from an ncu --set full capture: valid only for the kernel sources it was taken on
        traffic = None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            import hashlib
            h = hashlib.sha1()
            for fn in sorted(os.listdir(os.path.join(ROOT, "starfish_b200", "csrc"))):
                if fn.endswith(".cuh"):  # the kernels (sf_gpu.cu is host code)
                    h.update(open(os.path.join(ROOT, "starfish_b200", "csrc", fn), "rb").read())
            if tj.get("csrc_sha1") == h.hexdigest():
                traffic = tj.get(wname)
        except Exception:
            pass
        line = {
            "metric": "particle pushes/sec (move+deposit)", "value": value, "unit": "pushes/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "strong" if wname == "e" else "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl.name, "mesh_nodes": [m.ni, m.nj], "particles_per_gpu": n_local, "particles_total": int(n_local * world),
                       "l2": "particle arrays (64 B x N per GPU) exceed the 126 MB L2; no flush needed",
                       "parallelism": "particles partitioned by index, mesh replicated, NCCL allreduce of the deposit" if world > 1 else "single GPU"},
            "wall_ms_per_step": wall_ms / args.steps,
            "e2e": {"value": e2e_value, "unit": "pushes/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": 1e3 * e2e_s / args.steps},
            "gpu_launches": int(launches), "untiled_deposit_fraction": fallback / max(pushes, 1), "tile_halo": km.tileHalo()[0],
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "peak_source": peak_src, "kernel": names[dom], "kernel_ms_per_launch": d_ms / d_steps,
                         "algorithmic_bytes_per_push": ALGO_BYTES_PER_PUSH, "kernel_ms_per_step": ker_ms / args.steps,
                         "frac_all_step_kernels": (ALGO_BYTES_PER_PUSH * float(pushes) / (ker_ms * 1e-3) / 1e9 / peak) if ker_ms > 0 else 0.0,
                         "step_kernels": kernels,
                         "pushes_per_s_at_peak": peak * 1e9 / ALGO_BYTES_PER_PUSH},
            "clocks": clocks,
        }
        if multi_check is not None:
            line["multi_gpu_check"] = multi_check
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            n_sample = args.ref_particles
            v, t = cpu_reference_run(wl, n_sample, 3, 1, threads)
            line["cpu_baseline"] = {"value": v, "unit": "pushes/s", "cores": threads, "kind": "port",
                                    "sample": f"{n_sample} particles x 3 steps of the same workload, {threads} mover threads + serial deposit"}
        emit(line)
    km.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
