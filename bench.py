#!/usr/bin/env python3
"""bench.py -- cell-species-updates/s of the per-timestep solver core on an ensemble of 1-D columns.

Workload (BASELINE.json): the 4096-column EUV-flux sweep of configs[3]
(runs_mdot_vs_euv_plot/dynamic3_euv1e3_C_1s2b_rerun.par: 997 cells x 1 species, order-2 hydro + 2-band irradiation +
simple FLD), column k of the whole job has PARI_UVSTAR = 1e31 * 10^(-2 + 4k/(ncols_total-1)).  One "step" = one time
step of every column of the batch.  Weak scaling: every GPU holds --ncols columns (default 4096); ranks share nothing.

  python bench.py [--gpus N --steps K --warmup W]                  our arm (one process per GPU; torchrun for N>1)
  python bench.py --impl reference [--gpus N --steps K --warmup W]  the reference's CPU implementation (oracle/_ref)

Prints ONE JSON line (rank 0).  `value` is timed with CUDA events on the engine's stream with the state resident in HBM;
`e2e` goes through the C ABI with host buffers (H2D of the state and D2H of the result inside the timed region).
"""
import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
CASE = "c4_dynamic3_euv"
DECK_DIR, DECK_PAR = "runs_mdot_vs_euv_plot/", "dynamic3_euv1e3_C_1s2b_rerun.par"
METRIC = "cell-species-updates/sec (ensemble, FP64)"
BYTES_PER_UPDATE = 224.0          # 8*[H=18 + 2 + R=8] for configs[3]/[4]
BYTES_HYDRO, BYTES_RAD = 8.0 * 18, 8.0 * 10   # share of the hydro kernel (H) and of the radiation+tail kernel (R + e_last)


def uv_value(k, n_total):
    return 1e31 * 10.0 ** (-2.0 + 4.0 * k / max(n_total - 1, 1))


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU with NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if not self.nv:
            return
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                 "hw_power_brake_slowdown": 0x80, "sync_boost": 0x10}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for n, bit in names.items():
                    if r & bit:
                        self.reasons.add(n)
            except Exception:
                pass
            time.sleep(0.02)

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def run_reference_processes(nproc, nsteps, uv_list, workdir):
    """Run `nproc` copies of the unmodified reference (oracle/_ref/aiolos_ref_dump, one per host core), each advancing its
    own column (own PARI_UVSTAR) of configs[3] by `nsteps` time steps from t=0.  Returns (wall seconds, total steps)."""
    exe = os.path.join(ROOT, "oracle", "_ref", "aiolos_ref_dump")
    if not os.path.exists(exe):
        raise RuntimeError("oracle/_ref/aiolos_ref_dump missing: run __graft_entry__.build() in the dev container")
    procs, dirs = [], []
    for i in range(nproc):
        d = os.path.join(workdir, f"p{i}")
        shutil.copytree(os.path.join(ROOT, "oracle", "decks"), d)
        par = os.path.join(d, DECK_DIR, DECK_PAR)
        txt = open(par).read().splitlines()
        txt = [("PARI_UVSTAR             %.17g" % uv_list[i % len(uv_list)]) if ln.startswith("PARI_UVSTAR") else ln for ln in txt]
        open(par, "w").write("\n".join(txt) + "\n")
        dirs.append(d)
    t0 = time.perf_counter()
    for d in dirs:
        procs.append(subprocess.Popen([exe, "-dir", DECK_DIR, "-par", DECK_PAR, "-nsteps", str(nsteps), "-out", "bench"], cwd=d,
                                      stdin=subprocess.DEVNULL, stdout=open(os.path.join(d, "log.txt"), "w"), stderr=subprocess.STDOUT))
    for p in procs:
        p.wait()
    wall = time.perf_counter() - t0
    total = 0
    for d in dirs:
        done = [ln for ln in open(os.path.join(d, "log.txt")) if ln.startswith("Finished at time=")]
        if not done:
            raise RuntimeError("reference run failed: " + open(os.path.join(d, "log.txt")).read()[-500:])
        total += int(done[-1].split("after steps=")[1].split()[0])
    return wall, total


def reference_arm(args, rank):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    num_cells, S = 997, 1
    per_step = max(50, args.ref_steps)
    uvs = [uv_value(k * 4096 // cores, 4096) for k in range(cores)]
    with tempfile.TemporaryDirectory() as wd:
        for w in range(min(args.warmup, 1)):
            run_reference_processes(cores, per_step, uvs, os.path.join(wd, f"w{w}"))
        walls, steps = [], 0
        for k in range(args.steps):
            wall, n = run_reference_processes(cores, per_step, uvs, os.path.join(wd, f"s{k}"))
            walls.append(wall); steps += n
    total_t = sum(walls)
    value = steps * num_cells * S / total_t
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "updates/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": min(args.warmup, 1), "ms_per_step": 1e3 * total_t / max(args.steps, 1), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"configs[3] {DECK_PAR} (997 cells x 1 species), one column per host core, each "
                                   f"bench step = {per_step} time steps from t=0 of {cores} columns of the PARI_UVSTAR sweep"},
            "cpu_baseline": {"value": value, "unit": "updates/s", "cores": cores, "kind": "reference",
                             "sample": f"{cores} processes x {per_step} time steps x {args.steps} bench steps, unmodified "
                                       "reference sources + Eigen stand-in (oracle/Makefile), g++ -O3 -march=x86-64-v3"},
            "e2e": {"value": value, "unit": "updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--ncols", type=int, default=4096, help="columns per GPU (weak scaling)")
    ap.add_argument("--spinup", type=int, default=600, help="untimed steps before the warm-up: dt grows x1.1/step from "
                                                              "1e-20 s and reaches the CFL regime after ~450 steps")
    ap.add_argument("--ref-steps", type=int, default=4000, help="time steps per column per bench step of the reference arm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=100)
    ap.add_argument("--opt", action="append", default=[], metavar="KEY=VALUE", help="aio_set_option pairs (kernel variants)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        reference_arm(args, rank)
        return

    # everything but the final JSON line goes to stderr (NCCL prints its version banner on fd 1)
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    import numpy as np
    import torch
    import aiolos_b200 as A
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- aiolos_b200 has no CPU path (use --impl reference for the CPU arm)")
    if args.warmup < 3:
        args.warmup = 3
    torch.cuda.set_device(local)
    if world > 1:
        # one rank per GPU: keep the rank (and hence the pinned host buffers it first-touches for the end-to-end leg) on the
        # CPU cores / NUMA node next to its GPU; without this every rank's H2D traffic may cross the socket interconnect
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(local)
            words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
            cpus = {64 * w + b for w, m in enumerate(words) for b in range(64) if (m >> b) & 1}
            if cpus:
                os.sched_setaffinity(0, cpus)
        except Exception:      # noqa: BLE001  (affinity is an optimisation only)
            pass
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    ncols, n_total = args.ncols, args.ncols * world
    overrides = [(c, "PARI_UVSTAR", repr(uv_value(rank + c * world, n_total))) for c in range(ncols)]   # interleaved: k -> GPU k mod G
    # the reference-facing entry point: the deck's own .par/.spc/.opa files (oracle/decks holds the unmodified inputs of
    # configs[3]); opacity files are opened relative to the CWD exactly as the reference does (io.cpp:143)
    cwd = os.getcwd()
    os.chdir(os.path.join(ROOT, "oracle", "decks"))
    try:
        ens = A.Ensemble.from_files(DECK_DIR, DECK_PAR, "", ncols=ncols, overrides=overrides, device=local)
    finally:
        os.chdir(cwd)
    N, S = ens.num_cells, ens.num_species
    for kv in args.opt:
        ens.set_option(*kv.split("=", 1))

    def barrier():
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        torch.cuda.synchronize()

    ens.advance(args.spinup)
    ens.advance(args.warmup)
    steps_before = ens.total_steps
    launches_before = ens.kernel_launches
    sampler = ClockSampler(local); sampler.start()
    barrier()
    ms = ens.advance(args.steps, timed=True)           # CUDA events on the engine's stream around the K-step launch sequence
    barrier()
    sampler.stop_flag = True; sampler.join(timeout=2)
    launches = ens.kernel_launches - launches_before
    col_steps = ens.total_steps - steps_before          # columns x steps actually advanced (a crashed column stops counting)
    t = torch.tensor([ms, float(col_steps), float(launches)], dtype=torch.float64, device="cuda")
    if dist:
        tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms_all, col_steps_all, launches_all = tmax[0].item(), tsum[1].item(), tsum[2].item()
    else:
        ms_all, col_steps_all, launches_all = ms, float(col_steps), float(launches)
    value = col_steps_all * N * S / (ms_all * 1e-3)

    # ---- per-kernel device time (live, CUDA events between the launches) and roofline of the dominant kernel ----------------
    kt = ens.kernel_times(20)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_gbs = float(peaks.get("hbm_gbs", 6650.0)); peak_src = "measured" if "hbm_gbs" in peaks else "fallback"
    dom = max(("hydro", "radiation"), key=lambda k: kt[k])
    fused = kt["radiation"] == 0.0            # single-species decks run the whole step in one persistent kernel (k_step1)
    dom_bytes = (BYTES_PER_UPDATE if fused else BYTES_HYDRO if dom == "hydro" else BYTES_RAD) * ncols * N * S
    if fused:
        kt = {"step": kt["hydro"]}
        dom = "step"
    achieved = dom_bytes / (kt[dom] * 1e-3) / 1e9 if kt[dom] > 0 else 0.0
    roofline = {"bound": "hbm", "kernel": "k_step1<4>" if fused else "k_hydro2<4>" if dom == "hydro" else "k_rad3<4>", "achieved": achieved, "peak": peak_gbs,
                "unit": "GB/s", "frac": achieved / peak_gbs,
                # dram__bytes_read.sum + dram__bytes_write.sum of one k_step1<4,256> launch (16 steps, 4096 columns) from the
                # ncu --set full capture profiles/r1e_ncu_full_k_step1.txt, per time step: (267.6 + 191.9) MB / 16
                "traffic": (28.7e6 if (fused and ncols == 4096) else None), "traffic_unit": "bytes per step (all columns)",
                "algorithmic_bytes": BYTES_PER_UPDATE * ncols * N * S if fused else dom_bytes, "peak_source": peak_src,
                "kernel_ms": kt, "step_frac": BYTES_PER_UPDATE * ncols * N * S / (ms / args.steps * 1e-3) / 1e9 / peak_gbs}

    # ---- end to end through the C ABI with host buffers -----------------------------------------------------------------------
    # every e2e step: H2D of the step's input (the conserved state of all columns, from pinned host memory), one time step,
    # D2H of the step's result (per-column mass + per-column clock/status)
    u_host = torch.empty((3, ncols, S, N + 2), dtype=torch.float64).pin_memory()
    ens.download_state(u_host.numpy())
    mass = torch.empty(ncols, dtype=torch.float64).pin_memory()
    ens.upload_state(u_host.numpy()); ens.advance(1); ens.column_mass(mass.numpy())
    steps0 = ens.total_steps
    barrier()
    # pipelined: the H2D copy of step k+1's input (copy stream, second pinned buffer) overlaps step k's kernels
    u_host2 = u_host.clone().pin_memory()
    bufs = (u_host.numpy(), u_host2.numpy())
    t0 = time.perf_counter()
    ens.upload_state_async(bufs[0])
    for k in range(args.e2e_steps):
        ens.commit_upload()
        if k + 1 < args.e2e_steps:
            ens.upload_state_async(bufs[(k + 1) % 2])
        ens.advance(1)
        ens.column_mass(mass.numpy())
        st = ens.status_all()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    e2e_cs = ens.total_steps - steps0
    te = torch.tensor([e2e_s, float(e2e_cs)], dtype=torch.float64, device="cuda")
    if dist:
        a = te.clone(); dist.all_reduce(a, op=dist.ReduceOp.MAX)
        b = te.clone(); dist.all_reduce(b, op=dist.ReduceOp.SUM)
        e2e_s, e2e_cs = a[0].item(), b[1].item()
    e2e = {"value": e2e_cs * N * S / e2e_s, "unit": "updates/s", "h2d_bytes_per_step": int(u_host.numel() * 8),
           "d2h_bytes_per_step": int(ncols * 8 + ncols * (8 + 8 + 8 + 4)), "steps": args.e2e_steps}

    # ---- CPU baseline: the reference itself on this box's host cores (rank 0, N=1 only) -------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            cores = os.cpu_count() or 1
            with tempfile.TemporaryDirectory() as wd:
                wall, n = run_reference_processes(cores, 3000, [uv_value(k * 4096 // cores, 4096) for k in range(cores)], wd)
            cpu = {"value": n * N * S / wall, "unit": "updates/s", "cores": cores, "kind": "reference",
                   "sample": f"{cores} processes (one per host core) x 3000 time steps of configs[3] from t=0, unmodified reference "
                             "sources + Eigen stand-in, g++ -O3 -march=x86-64-v3"}
        except Exception as ex:           # the baseline is a reported number, never a reason to lose the bench line
            cpu = {"value": None, "unit": "updates/s", "cores": os.cpu_count(), "kind": "reference", "sample": f"failed: {ex}"}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "updates/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_all / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic",
                "config": {"workload": f"{ncols}-column/GPU PARI_UVSTAR sweep of configs[3] {DECK_PAR} ({N} cells x {S} species, "
                                       "order-2 hydro + 2-band irradiation + simple FLD), built by aio_create from the deck's own .par/.spc/.opa, "
                                       f"{args.spinup} untimed spin-up steps",
                           "columns_per_gpu": ncols, "num_cells": N, "num_species": S, "parallelism": f"columns x{world}",
                           "l2": "inputs larger than L2 (state ~%.0f MB per GPU)" % (ncols * (N + 2) * 8 * 9 / 1e6)},
                "clocks": sampler.summary(), "e2e": e2e, "gpu_launches": int(launches_all), "roofline": roofline,
                "cpu_baseline": cpu}
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        if args.opt:
            line["config"]["options"] = args.opt
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    if dist:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
