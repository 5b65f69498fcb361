#!/usr/bin/env python
"""bench.py -- Navier2D time steps/s (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Our arm: one process per GPU.  A "step" is one Navier2D.update() (navier.rs:737-765) on
synthetic initial fields (set_velocity/set_temperature, no RNG).  At N=1 the workload is the
configuration the metric is quoted on that fits one GPU: confined 2048x2049, Ra=1e9.  For
N>1 the ranks run independent replicas of that workload ("replicas only", weak scaling --
DESIGN.md section 6); `--workload periodic8192` (or any periodic workload) at N>1 runs ONE
problem slab-decomposed over the Fourier modes instead (strong scaling, fused NVLink
transposes).  Timing: W>=3 warm-up steps, then exactly K steps between CUDA events on the
launching stream with a barrier + synchronize on both sides, max over ranks.  The working
set (>2 GB) is far larger than L2, so no L2 flush is needed between steps.  `e2e` is the
same metric through the C ABI with HOST buffers: every step uploads the whole state from
pinned host memory (double-buffered on a copy stream, rp_navier_stage_state) and reads
|div|_2 back.

Reference arm (--impl reference): the reference is pure Rust and cannot be built in this
image, so this times the CPU restatement (oracle/, numpy/scipy with all host threads) of
the same update() on the same configuration; each requested step is one full update, and
the number of steps actually run is bounded so the arm ends within a few minutes.
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

WORKLOADS = {
    # name: (periodic, nx, ny, ra, pr, dt, aspect, adiabatic, description)
    "confined2048": (False, 2048, 2049, 1e9, 1.0, 1e-4, 1.0, True, "Navier2D::new confined 2048x2049 Ra=1e9 Pr=1 dt=1e-4 adiabatic"),
    "periodic512": (True, 512, 513, 1e7, 1.0, 2e-3, 1.0, True, "Navier2D::new_periodic 512x513 Ra=1e7 Pr=1 dt=2e-3"),
    "confined64": (False, 64, 64, 1e5, 1.0, 0.02, 1.0, True, "Navier2D::new 64x64 Ra=1e5 Pr=1 dt=0.02 adiabatic"),
    "periodic2048": (True, 2048, 2049, 1e9, 1.0, 1e-4, 1.0, True, "Navier2D::new_periodic 2048x2049 Ra=1e9 Pr=1 dt=1e-4"),
    "periodic8192": (True, 8192, 8193, 1e10, 1.0, 2e-5, 1.0, True, "Navier2D::new_periodic 8192x8193 Ra=1e10 Pr=1 dt=2e-5"),
    "confined1024": (False, 1024, 1025, 1e8, 1.0, 2e-4, 1.0, True, "Navier2D::new confined 1024x1025 Ra=1e8 Pr=1 dt=2e-4 adiabatic"),
}
METRIC = "Navier2D time steps/sec at Nx(N+1) (device-timed)"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_oracle(wl, eig=None):
    import oracle as O
    periodic, nx, ny, ra, pr, dt, aspect, adiabatic, _ = wl
    if periodic:
        o = O.Navier2D.new_periodic(nx, ny, ra, pr, dt, aspect, banded=True)
    else:
        o = O.Navier2D.new(nx, ny, ra, pr, dt, aspect, adiabatic, banded=True, eig_data=eig)
    o.set_velocity(0.2, 1.0, 1.0)
    o.set_temperature(0.2, 1.0, 1.0)
    return o


def time_oracle(wl, want_steps, warmup, budget_s, eig=None):
    """Time the CPU restatement; returns (steps_per_s, steps_run, description)."""
    o = make_oracle(wl, eig)
    t0 = time.perf_counter()
    o.update()
    t1 = time.perf_counter() - t0
    nwarm = max(0, min(warmup - 1, int(0.15 * budget_s / max(t1, 1e-9))))
    for _ in range(nwarm):
        o.update()
    nrun = max(1, min(want_steps, int(budget_s / max(t1, 1e-9))))
    t0 = time.perf_counter()
    for _ in range(nrun):
        o.update()
    dt = time.perf_counter() - t0
    return nrun / dt, nrun, "%d full update() steps of the numpy/scipy restatement (of %d requested; %d warm-up)" % (nrun, want_steps, nwarm + 1)


def run_reference(args, wl, rank):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    val, nrun, sample = time_oracle(wl, args.steps, args.warmup, budget_s=150.0)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "steps/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 / val, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic (set_velocity(0.2,1,1)+set_temperature(0.2,1,1), no RNG)",
        "config": {"workload": wl[8], "note": "CPU restatement of rustpde (oracle/), not the rustpde binary: no Rust toolchain in this image"},
        "cpu_baseline": {"value": val, "unit": "steps/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_ours(args, wl, rank, world, local_rank):
    import numpy as np
    import torch

    import rustpde_b200 as R
    from rustpde_b200 import _ffi

    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lib = _ffi.product_lib(local_rank)
    periodic, nx, ny, ra, pr, dt, aspect, adiabatic, desc = wl
    t_setup = time.perf_counter()
    nav = R.Navier2D.new_periodic(nx, ny, ra, pr, dt, aspect, lib=lib) if periodic else R.Navier2D.new(nx, ny, ra, pr, dt, aspect, adiabatic, lib=lib)
    nav.set_velocity(0.2, 1.0, 1.0)
    nav.set_temperature(0.2, 1.0, 1.0)
    t_setup = time.perf_counter() - t_setup
    W = max(3, args.warmup)
    slab = None
    if periodic and (world > 1 or args.parallel == "slab") and args.parallel != "replicas":
        # strong scaling: ONE problem, slab-decomposed over the Fourier modes kx, NCCL all-to-all transposes
        from rustpde_b200.slab import Navier2DSlab
        slab = Navier2DSlab(nav, transport=args.transport)
    stepper = slab if slab is not None else nav
    stepper.update(W)
    stepper.sync()

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    stepper.update(args.steps)
    e1.record()
    barrier()
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1)
    if dist is not None:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    steps_per_s = (1 if slab is not None else world) * args.steps / (ms * 1e-3)
    launches = nav.launches_per_step() if slab is None else 9  # our kernels per slab step (+ torch pack copies, NCCL)
    if slab is not None:
        slab.gather_state()
    div = nav.div_norm()
    if slab is not None:
        if rank == 0:
            line = {
                "metric": METRIC, "value": steps_per_s, "unit": "steps/s", "n_gpus": world, "steps": args.steps, "warmup": W,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic (set_velocity(0.2,1,1)+set_temperature(0.2,1,1), no RNG)",
                "config": {"workload": desc, "parallelism": "kx slabs x%d, 9 transposes per step, %s" % (world, "fused into the kernels over NVLink peer memory (CUDA IPC)" if slab.transport == "p2p" else "NCCL all_to_all"),
                           "l2": "working set >> 126 MB L2, no flush needed", "cuda_graph": False, "setup_s": round(t_setup, 2),
                           "div_norm_after": div, "nvlink_egress_bytes_per_rank_per_step": slab.bytes_exchanged_per_step},
                "clocks": clocks, "e2e": None, "gpu_launches": launches * args.steps, "roofline": None, "cpu_baseline": None,
            }
            print(json.dumps(line), flush=True)
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- e2e: HOST buffers through the C ABI; H2D of the state + D2H of the step's metric inside the timed region
    fields = [nav.temp, nav.ux, nav.uy, nav.pres[0]]
    pinned = []
    for f in fields:
        a = f.vhat
        t = torch.empty(a.view(np.float64).size, dtype=torch.float64).pin_memory()
        t.numpy()[:] = a.view(np.float64).reshape(-1)
        pinned.append(t)
    h2d = sum(t.numel() * 8 for t in pinned)
    ne2e = max(3, min(args.steps, 20))

    def e2e_step_serial():
        for f, t in zip(fields, pinned):
            lib.call("rp_field_upload_vhat", f._h, _ffi.C.cast(t.data_ptr(), _ffi.c_double_p), t.numel())
        nav.update(1)
        return nav.div_norm()  # reference: integrate() calls exit() -> |div| on the host every step (lib.rs:182)

    def stage():
        nav.stage_state(*[(t.data_ptr(), t.numel()) for t in pinned])

    def e2e_step():
        # double-buffered: the state uploaded during the previous step is moved into place, update(1) is queued,
        # the upload of the next step's inputs is queued on the copy stream (it overlaps this step's kernels),
        # then |div|_2 of this step comes back to the host
        nav.commit_staged()
        nav.update(1)
        stage()
        return nav.div_norm()

    def timed(step):
        step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(ne2e):
            step()
        torch.cuda.synchronize()
        te = time.perf_counter() - t0
        if dist is not None:
            t = torch.tensor([te], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            te = float(t.item())
        return te

    te_serial = timed(e2e_step_serial)
    stage()
    te = timed(e2e_step)
    e2e = {"value": world * ne2e / te, "unit": "steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 8,
           "steps": ne2e, "serial_value": world * ne2e / te_serial,
           "note": "per step: temp/ux/uy/pres vhat from pinned host memory (rp_navier_stage_state on a copy stream, "
                   "overlapping the previous step's kernels; rp_navier_commit_staged), update(1), |div|_2 back to the host; "
                   "serial_value = the same without the overlap (rp_field_upload_vhat x4, update, |div|)"}

    # ---- roofline of the dominant kernel (per-launch CUDA-event times of eagerly launched steps)
    roof = None
    if rank == 0:
        prof = nav.profile(5)
        tot = sum(o["ms"] for o in prof)
        # group launches of the same kernel program
        groups = {}
        for o in prof:
            g = groups.setdefault(o["name"], {"ms": 0.0, "bytes": 0.0, "flops": 0.0, "n": 0})
            g["ms"] += o["ms"]
            g["bytes"] += o["bytes"]
            g["flops"] += o["flops"]
            g["n"] += 1
        top = max(groups.items(), key=lambda kv: kv[1]["ms"])
        name, g = top
        hbm_peak, src = measured_peaks()
        traffic = None
        tj = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tj):
            try:
                # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` capture
                traffic = json.load(open(tj)).get(args.workload, {}).get(name)
            except Exception:
                traffic = None
        if g["flops"] > 0:
            # FP64 tensor (DMMA) bound: measure the denominator here with cuBLAS DGEMM
            a = torch.randn(4096, 4096, device="cuda", dtype=torch.float64)
            bmat = torch.randn(4096, 4096, device="cuda", dtype=torch.float64)
            torch.matmul(a, bmat)
            torch.cuda.synchronize()
            best = 1e9
            for _ in range(3):
                s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s0.record()
                torch.matmul(a, bmat)
                s1.record()
                torch.cuda.synchronize()
                best = min(best, s0.elapsed_time(s1))
            peak = 2 * 4096 ** 3 / (best * 1e-3) / 1e12
            ach = g["flops"] / (g["ms"] * 1e-3) / 1e12
            roof = {"bound": "tensor", "kernel": name, "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                    "traffic": traffic, "peak_source": "cuBLAS DGEMM 4096^3 measured in this run (f64)",
                    "share_of_step": g["ms"] / tot}
        else:
            ach = g["bytes"] / (g["ms"] * 1e-3) / 1e9
            roof = {"bound": "hbm", "kernel": name, "launches_per_step": g["n"], "achieved": ach, "peak": hbm_peak, "unit": "GB/s",
                    "frac": ach / hbm_peak, "traffic": traffic, "peak_source": src, "share_of_step": g["ms"] / tot,
                    "algorithmic_bytes_per_launch": g["bytes"] / g["n"]}
        roof["per_kernel"] = [{"kernel": k, "launches": v["n"], "ms": round(v["ms"], 4),
                               "GBps": round(v["bytes"] / (v["ms"] * 1e-3) / 1e9, 1) if v["ms"] > 0 else None,
                               "TFLOPs": round(v["flops"] / (v["ms"] * 1e-3) / 1e12, 2) if v["flops"] > 0 else None}
                              for k, v in sorted(groups.items(), key=lambda kv: -kv[1]["ms"])]

    # ---- CPU baseline beside it (rank 0, N = 1): bounded sample of the same workload
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        eig = None if periodic else nav.export_eig()
        if eig is not None:
            lam = eig[0].copy()
            if abs(lam[0] + 1e-10) < 1e-10:
                lam = lam + 1e-10  # un-shift: the oracle's Poisson applies poisson.rs:80-83 itself
            eig = (lam, eig[1], eig[2])
        val, nrun, sample = time_oracle(wl, 4, 1, budget_s=20.0, eig=eig)
        cpu = {"value": val, "unit": "steps/s", "cores": os.cpu_count() or 1, "kind": "port", "sample": sample}

    if rank == 0:
        line = {
            "metric": METRIC, "value": steps_per_s, "unit": "steps/s", "n_gpus": world, "steps": args.steps, "warmup": W,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic (set_velocity(0.2,1,1)+set_temperature(0.2,1,1), no RNG)",
            "config": {"workload": desc, "parallelism": "replicas x%d" % world if world > 1 else "1 GPU",
                       "l2": "working set > 2 GB >> 126 MB L2, no flush needed", "cuda_graph": True,
                       "setup_s": round(t_setup, 2), "div_norm_after": div},
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches * args.steps, "roofline": roof, "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="confined2048", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--transport", default="p2p", choices=["p2p", "collective"],
                    help="slab mode: fused transposes over NVLink peer memory (p2p) or NCCL all_to_all (collective)")
    ap.add_argument("--parallel", default="auto", choices=["auto", "replicas", "slab"],
                    help="N > 1: independent replicas (confined path) or one slab-decomposed problem (periodic path)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, wl, rank)
        return
    if world == 1 and args.gpus > 1:
        # launched without torchrun: re-exec under torch.distributed.run
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", "29531", os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    run_ours(args, wl, rank, world, local_rank)


if __name__ == "__main__":
    main()
