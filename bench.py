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
The same line carries
  "strong":    configs[4] as BASELINE words it -- ONE 4096-column ensemble split over the N GPUs (4096/N columns each);
  "secondary": at N=1, driver-timed lines for configs[1] (6-species implicit friction) and configs[2] (22-band coupled
               J_b/T_s block-tridiagonal solve), each with its own clocks record and an FP64 roofline against the DFMA
               rate measured live on the same GPU (aio_device_peaks).
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
STRONG_TOTAL = 4096               # configs[4]: one 4096-column ensemble sharded over 1/2/4/8 GPUs
REF_EXE = "aiolos_ref_dump_s1"    # the reference built with NUM_SPECIES fixed at compile time (=1), BASELINE.md section 3
REF_FLAGS = "g++ -O3 -march=x86-64-v3 -DNUM_SPECIES=1, unmodified reference sources + Eigen stand-in (oracle/Makefile)"
# Secondary, driver-timed cases: name -> (deck dir, par, spc, spin-up steps, timed steps, kernel slot, flop per
# cell-species-update of that kernel, what the flop figure counts)
SECONDARY = {
    "configs[1]": ("test_files/", "friction_6spc.par", "friction_6spc.spc", 100, 50, "friction", 270.0,
                   "k_friction<6>: alpha matrix + 6x6 LU + substitution + heating, ~1.6 kflop per cell / 6 species"),
    "configs[2]": ("test2/", "SEarth_opacities_1s22b.par", "irradiation.spc", 12, 20, "radiation", 57.0e3,
                   "full FLD: LU of the 23x23 block + 23 right-hand sides + back sweep per cell (SURVEY 8d: ~57 kflop)"),
}


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
    exe = os.path.join(ROOT, "oracle", "_ref", REF_EXE)
    if not os.path.exists(exe):
        raise RuntimeError(f"oracle/_ref/{REF_EXE} missing: run __graft_entry__.build() in the dev container")
    procs, dirs = [], []
    for i in range(nproc):
        d = os.path.join(workdir, f"p{i}")
        shutil.copytree(os.path.join(ROOT, "decks"), d)
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
                             "sample": f"{cores} processes x {per_step} time steps x {args.steps} bench steps; {REF_FLAGS}"},
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
    ap.add_argument("--no-secondary", action="store_true", help="skip the configs[1]/configs[2] lines (N=1 only anyway)")
    ap.add_argument("--opt2", action="append", default=[], metavar="KEY=VALUE", help="aio_set_option pairs for the secondary cases")
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

    def barrier():
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_max_sum(vals):
        t = torch.tensor(vals, dtype=torch.float64, device="cuda")
        if not dist:
            return list(vals), list(vals)
        a = t.clone(); dist.all_reduce(a, op=dist.ReduceOp.MAX)
        b = t.clone(); dist.all_reduce(b, op=dist.ReduceOp.SUM)
        return a.tolist(), b.tolist()

    def build_sweep(ncols_rank, n_total):
        """Column c of this rank is member rank + c*world of the n_total-column PARI_UVSTAR sweep (interleaved: k -> GPU k mod G).
        The reference-facing entry point: the deck's own .par/.spc/.opa files under decks/ (what a user of the reference has on
        disk); opacity files are opened relative to the CWD exactly as the reference does (io.cpp:143)."""
        ov = [(c, "PARI_UVSTAR", repr(uv_value(rank + c * world, n_total))) for c in range(ncols_rank)]
        cwd = os.getcwd()
        os.chdir(os.path.join(ROOT, "decks"))
        try:
            e = A.Ensemble.from_files(DECK_DIR, DECK_PAR, "", ncols=ncols_rank, overrides=ov, device=local)
        finally:
            os.chdir(cwd)
        for kv in args.opt:
            e.set_option(*kv.split("=", 1))
        return e

    def device_leg(ens, steps, warmup, spinup):
        """spin-up + warm-up untimed, then exactly `steps` steps between barriers, CUDA events on the engine's stream.  When
        that region is shorter than ~0.25 s (the driver's --steps 20 is 10 ms of the fused kernel) the region is repeated --
        each repeat again exactly `steps` steps between barriers -- and the MEDIAN region is the one reported; the spread goes
        into the line.  Returns (max-over-ranks ms, sum-over-ranks column-steps, sum launches, clocks of this rank, this
        rank's ms, info)."""
        ens.advance(spinup)
        ms_w = ens.advance(warmup, timed=True)
        est = max(ms_w / max(warmup, 1) * steps, 1e-3)
        mx, _ = reduce_max_sum([est])
        repeats = int(min(64, max(1, -(-250.0 // mx[0]))))
        sampler = ClockSampler(local); sampler.start()
        regions = []
        for _ in range(repeats):
            s0, l0 = ens.total_steps, ens.kernel_launches
            barrier()
            ms = ens.advance(steps, timed=True)
            barrier()
            mxr, smr = reduce_max_sum([ms, float(ens.total_steps - s0), float(ens.kernel_launches - l0)])
            regions.append((mxr[0], smr[1], smr[2], ms))
        sampler.stop_flag = True; sampler.join(timeout=2)
        order = sorted(range(repeats), key=lambda k: regions[k][0])
        med = regions[order[repeats // 2]]
        info = {"repeats": repeats, "ms_per_step_min": regions[order[0]][0] / steps, "ms_per_step_max": regions[order[-1]][0] / steps}
        return med[0], med[1], med[2], sampler.summary(), med[3], info

    def e2e_leg(ens, nsteps):
        """Every e2e step: H2D of the step's input (the conserved state of all columns, from pinned host memory), one time step,
        D2H of the step's result (per-column mass + per-column clock/status).  Pipelined: the H2D copy of step k+1's input
        (copy stream, second pinned buffer) overlaps step k's kernels."""
        nc, S_, N_ = ens.ncols, ens.num_species, ens.num_cells
        u_host = torch.empty((3, nc, S_, N_ + 2), dtype=torch.float64).pin_memory()
        ens.download_state(u_host.numpy())
        mass = torch.empty(nc, dtype=torch.float64).pin_memory()
        ens.upload_state(u_host.numpy()); ens.advance(1); ens.column_mass(mass.numpy())
        u_host2 = u_host.clone().pin_memory()
        bufs = (u_host.numpy(), u_host2.numpy())
        regions = []
        for _ in range(3):                     # three regions of `nsteps` end-to-end steps each; the median one is reported
            steps0 = ens.total_steps
            barrier()
            t0 = time.perf_counter()
            ens.upload_state_async(bufs[0])
            for k in range(nsteps):
                ens.commit_upload()
                if k + 1 < nsteps:
                    ens.upload_state_async(bufs[(k + 1) % 2])
                ens.advance(1)
                ens.column_mass(mass.numpy())
                ens.status_all()
            torch.cuda.synchronize()
            e2e_s = time.perf_counter() - t0
            mx, sm = reduce_max_sum([e2e_s, float(ens.total_steps - steps0)])
            regions.append(sm[1] * N_ * S_ / mx[0])
        regions.sort()
        return {"value": regions[1], "unit": "updates/s", "h2d_bytes_per_step": int(u_host.numel() * 8),
                "d2h_bytes_per_step": int(nc * 8 + nc * (8 + 8 + 8 + 4)), "steps": nsteps, "regions": regions}

    def h2d_ceiling(nbytes, reps=24):
        """Platform ceiling of the end-to-end leg: the same bytes as one step's upload, as plain pinned-host -> device
        cudaMemcpyAsync calls (torch copy_, one copy per buffer) on this rank's GPU while every other rank does the same and
        nothing else runs.  Each copy is timed with CUDA events; returns the best GB/s of the slowest rank."""
        h = torch.empty(nbytes // 8, dtype=torch.float64).pin_memory()
        h.zero_()
        d = torch.empty(nbytes // 8, dtype=torch.float64, device="cuda")
        for _ in range(6):
            d.copy_(h, non_blocking=True)
        torch.cuda.synchronize()
        barrier()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
        for e0, e1 in ev:
            e0.record(); d.copy_(h, non_blocking=True); e1.record()
        torch.cuda.synchronize()
        best_ms = min(e0.elapsed_time(e1) for e0, e1 in ev)
        mx, _ = reduce_max_sum([best_ms])
        del h, d
        return nbytes / (mx[0] * 1e-3) / 1e9

    ncols, n_total = args.ncols, args.ncols * world
    ens = build_sweep(ncols, n_total)
    N, S = ens.num_cells, ens.num_species
    ms_all, col_steps_all, launches_all, clocks, ms, timing = device_leg(ens, args.steps, args.warmup, args.spinup)
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
    roofline = {"bound": "hbm", "kernel": ens.kernel_name(dom), "achieved": achieved, "peak": peak_gbs,
                "unit": "GB/s", "frac": achieved / peak_gbs,
                # DRAM bytes are not observable from inside the process: the ncu --set full capture of this kernel is under
                # profiles/ (r2_*), quoted in DESIGN.md section 5, and not repeated here as a constant
                "traffic": None,
                "algorithmic_bytes": BYTES_PER_UPDATE * ncols * N * S if fused else dom_bytes, "peak_source": peak_src,
                "kernel_ms": kt, "step_frac": BYTES_PER_UPDATE * ncols * N * S / (ms / args.steps * 1e-3) / 1e9 / peak_gbs}

    # ---- end to end through the C ABI with host buffers -----------------------------------------------------------------------
    e2e = e2e_leg(ens, args.e2e_steps)
    ens.close()
    # how much of the platform's host -> device copy rate the end-to-end leg reaches (the upload of the state is what bounds it)
    try:
        gbs = h2d_ceiling(e2e["h2d_bytes_per_step"])
        e2e["h2d_ceiling_gbs_per_gpu"] = gbs
        e2e["ceiling_value"] = gbs * 1e9 / e2e["h2d_bytes_per_step"] * ncols * N * S * world     # updates/s if a step cost only its upload
        e2e["frac_of_ceiling"] = e2e["value"] / e2e["ceiling_value"]
    except Exception as ex:       # noqa: BLE001
        e2e["h2d_ceiling_gbs_per_gpu"] = None; e2e["ceiling_note"] = str(ex)

    # ---- strong scaling: configs[4] as worded -- one 4096-column ensemble over the N GPUs -------------------------------------
    if world == 1 and ncols == STRONG_TOTAL:
        strong = {"value": value, "ms_per_step": ms_all / args.steps, "columns_total": STRONG_TOTAL, "columns_per_gpu": ncols,
                  "e2e": e2e["value"], "note": "N=1: identical to the weak leg"}
    elif STRONG_TOTAL % world == 0:
        es = build_sweep(STRONG_TOTAL // world, STRONG_TOTAL)
        s_ms, s_cs, _, s_clk, _, s_info = device_leg(es, args.steps, args.warmup, args.spinup)
        s_e2e = e2e_leg(es, args.e2e_steps)
        es.close()
        strong = {"value": s_cs * N * S / (s_ms * 1e-3), "ms_per_step": s_ms / args.steps, "columns_total": STRONG_TOTAL,
                  "columns_per_gpu": STRONG_TOTAL // world, "e2e": s_e2e["value"], "clocks": s_clk, "timing": s_info}
    else:
        strong = None

    # ---- secondary driver-timed cases (N=1): configs[1] and configs[2] with an FP64 roofline --------------------------------------
    secondary = None
    if world == 1 and not args.no_secondary:
        secondary = {}
        try:
            dp = A.device_peaks(local, copy=False)
            fp64_peak, fp64_src = dp["fp64_tflops"], "measured live (aio_device_peaks: DFMA chains, CUDA events)"
        except Exception as ex:       # noqa: BLE001
            fp64_peak, fp64_src = 37.0, f"fallback (SURVEY 8d); live measurement failed: {ex}"
        for name, (ddir, par, spc, spin, nst, slot, flop, what) in SECONDARY.items():
            try:
                cwd = os.getcwd()
                os.chdir(os.path.join(ROOT, "decks"))
                try:
                    e2 = A.Ensemble.from_files(ddir, par, spc, ncols=args.ncols, device=local)
                finally:
                    os.chdir(cwd)
                for kv in args.opt2:
                    e2.set_option(*kv.split("=", 1))
                m_all, cs_all, l_all, clk, _, info2 = device_leg(e2, nst, 3, spin)
                k2 = e2.kernel_times(min(nst, 10))
                upd = args.ncols * e2.num_cells * e2.num_species
                ach = flop * upd / (k2[slot] * 1e-3) / 1e12 if k2[slot] > 0 else 0.0
                secondary[name] = {
                    "workload": f"{args.ncols} identical columns of {ddir}{par} ({e2.num_cells} cells x {e2.num_species} species, "
                                f"{e2.num_bands_out} outgoing bands), {spin} untimed spin-up steps",
                    "value": cs_all * e2.num_cells * e2.num_species / (m_all * 1e-3), "unit": "updates/s", "steps": nst,
                    "ms_per_step": m_all / nst, "gpu_launches": int(l_all), "clocks": clk, "timing": info2, "kernel_ms": k2,
                    "roofline": {"bound": "fp64", "kernel": e2.kernel_name(slot), "achieved": ach, "peak": fp64_peak,
                                 "unit": "TFLOP/s", "frac": ach / fp64_peak if fp64_peak else None, "peak_source": fp64_src,
                                 "flop_per_update": flop, "counts": what, "traffic": None},
                    "state": e2.status(0)}
                e2.close()
            except Exception as ex:       # noqa: BLE001  (a secondary line never costs the headline line)
                secondary[name] = {"failed": str(ex)}

    # ---- CPU baseline: the reference itself on this box's host cores (rank 0, N=1 only) -------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            cores = os.cpu_count() or 1
            with tempfile.TemporaryDirectory() as wd:
                wall, n = run_reference_processes(cores, 3000, [uv_value(k * 4096 // cores, 4096) for k in range(cores)], wd)
            cpu = {"value": n * N * S / wall, "unit": "updates/s", "cores": cores, "kind": "reference",
                   "sample": f"{cores} processes (one per host core) x 3000 time steps of configs[3] from t=0; {REF_FLAGS}"}
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
                "clocks": clocks, "timing": timing, "e2e": e2e, "gpu_launches": int(launches_all), "roofline": roofline,
                "cpu_baseline": cpu, "strong": strong, "secondary": secondary}
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
