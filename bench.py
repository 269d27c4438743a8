#!/usr/bin/env python
"""bench.py — EulerBeam3D residual+Jacobian element-assemblies/s (BASELINE.json metric) on N B200s.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--nele E] [--ox 0|2]
  torchrun … bench.py --gpus N …      (one rank per GPU; RANK/LOCAL_RANK/WORLD_SIZE/MASTER_* from the env)

Workload (config.workload): BASELINE.json configs[2] — the inspect/PerformanceEulerBeam3D-style synthetic chain of 10M
EulerBeam3D elements per GPU, one SweepX `assemble!{:iter}` per step = element kernels + deterministic
reduction into Lλ and the CSC nzval.  `value`: state resident in HBM.  `e2e`: through the C-ABI call
mb_sweepx_assemble with pinned HOST buffers (H2D state, D2H Lλ + nzval inside the timed region).
Inputs/outputs per step (≈1.3 GB geometry+maps read, ≈20 GB written/re-read) exceed the 126 MB L2: no L2 flush needed.

`--impl reference` times the reference algorithm's CPU restatement (oracle/, "port": Julia is not in this image) on the
host cores, on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "EulerBeam3D residual+Jacobian element-assemblies/s"
UNIT = "element-assemblies/s"

# Executed FP64 work of the beam kernel per element, this formulation (forward-over-reverse, 12 directional lanes), counted
# by ncu (profiles/): DFMA·2 + DMUL + DADD per element.  See DESIGN.md §roofline.
FLOP_PER_ELEMENT = {0: None, 2: None}   # filled from profiles/flops.json when present
# algorithmic HBM bytes per element-assembly of the element kernel: geometry 128 + dof idx 48 + X gather 96·ND
# + element tangent 1152 + residual 96 (write)
def alg_bytes_per_element(OX):
    return 128 + 48 + 96 * (OX + 1) + 1152 + 96


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--nele", type=float, default=1e7)
    ap.add_argument("--ox", type=int, default=0)
    ap.add_argument("--cpu-sample", type=int, default=8000)
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush(); self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f.read().strip().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if sm:
            out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        return out


def cpu_port_rate(mb, OX, nsample, nthreads):
    """oracle (literal restatement of the reference algorithm) on `nsample` elements of the same chain/state."""
    from oracle import elements as OE, pattern as OP
    eleobj, idx, ndof = mb.synthetic.chain(nsample, dynamic=OX > 0)
    X = mb.synthetic.state(ndof, nder=OX + 1)
    nm = mb.synthetic.newmark_coefficients(OX, 0.3)
    dis = [dict(X=idx, U=np.zeros((nsample, 0), np.int64), A=np.zeros((nsample, 0), np.int64))]
    a1, a2, cp, rv = OP.prepare_sweepx(dis, ndof, 0, 0)
    L = np.zeros(ndof); nz = np.zeros(len(rv))
    a1t, a2t = np.ascontiguousarray(a1[0].T), np.ascontiguousarray(a2[0].T)
    t = time.perf_counter()
    if nthreads == 1:
        OE.sweepx_assemble_beams(eleobj, idx, a1t, a2t, OX, "iter", X, np.ones(12), nm, L, nz)
    else:
        OE.sweepx_assemble_beams_mt(eleobj, idx, a1t, a2t, OX, X, np.ones(12), nm, L, nz, nthreads)
    dt = time.perf_counter() - t
    return nsample / dt, dt


def run_reference(args, rank, world):
    """reference arm: the reference algorithm's CPU port on the host cores (rank 0 only)."""
    if rank != 0:
        return
    import muscade_b200 as mb
    from oracle import elements as OE
    OX = args.ox
    cores = OE.max_threads()
    nsample = max(1000, int(args.cpu_sample) * max(1, cores // 2))
    for _ in range(max(1, min(args.warmup, 1))):
        cpu_port_rate(mb, OX, 500, cores)
    rates, times = [], []
    for _ in range(args.steps):
        r, dt = cpu_port_rate(mb, OX, nsample, cores)
        rates.append(r); times.append(dt)
    value = nsample * len(times) / sum(times)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": workload_config(args, OX, nsample, note="bounded sample of the same chain/state; element loop spread over all host threads "
                                      "(the reference's own loop src/Assemble.jl:479 is serial), serial scatter"),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": "%d elements x %d steps" % (nsample, args.steps)},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def workload_config(args, OX, nele, note=None):
    c = {"workload": "inspect/PerformanceEulerBeam3D-style synthetic chain, %d EulerBeam3D elements per GPU, SweepX{%d} assemble!{:iter} "
                     "(residual + 12x12 tangent per element, scatter into Llambda and CSC nzval)" % (nele, OX),
         "elements_per_gpu": int(nele), "OX": OX, "mission": "iter", "l2": "inputs/outputs larger than L2, no flush"}
    if note:
        c["note"] = note
    return c


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    import muscade_b200 as mb
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    OX = args.ox
    N = int(args.nele)
    eng = mb.Engine(local)
    fp64_peak = eng.fp64_tflops()
    eleobj, idx, ndof = mb.synthetic.chain(N, dynamic=OX > 0)
    eng.add_eulerbeam3d(eleobj, idx, np.ones(12))
    del eleobj, idx
    nnz = eng.sweepx_prepare(ndof)
    X = mb.synthetic.state(ndof, nder=OX + 1, seed=0x5EED + 7 * rank)
    nm = mb.synthetic.newmark_coefficients(OX, 0.3)
    # upload state once (resident in HBM for `value`)
    Lh = np.empty(ndof); nzh = np.empty(nnz)
    eng.sweepx_assemble(OX, "iter", X, nm, Llambda=Lh, nzval=nzh)

    def barrier():
        eng.sync()
        if dist is not None:
            dist.barrier()

    # ---- device-resident timing: W warm-up, K timed steps, CUDA events on the engine's stream via mb_sweepx_time_dev
    for _ in range(args.warmup):
        eng.sweepx_assemble_dev(OX, "iter", nm)
    barrier()
    launches0 = eng.launch_count()
    sampler = ClockSampler(local) if rank == 0 else None
    t0 = time.perf_counter()
    el_ms, ga_ms = eng.time_dev(OX, "iter", nm, reps=args.steps)      # events bracket every launch set; returns averages
    eng.sync()
    wall = time.perf_counter() - t0
    launches = eng.launch_count() - launches0
    clocks = sampler.stop() if sampler else None
    step_ms = el_ms + ga_ms
    if dist is not None:
        import torch
        tt = torch.tensor([step_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        step_ms = float(tt.item())
    value = world * N / (step_ms * 1e-3)

    # ---- end to end through the C ABI with pinned host buffers
    e2e = None
    if not args.no_e2e:
        for a in X + [Lh, nzh]:
            eng.pin(a)
        for _ in range(2):
            eng.sweepx_assemble(OX, "iter", X, nm, Llambda=Lh, nzval=nzh)
        barrier()
        ts = []
        for _ in range(max(3, min(args.steps, 5))):
            t1 = time.perf_counter()
            eng.sweepx_assemble(OX, "iter", X, nm, Llambda=Lh, nzval=nzh)
            ts.append(time.perf_counter() - t1)
        e2e_ms = 1e3 * float(np.mean(ts))
        if dist is not None:
            import torch
            tt = torch.tensor([e2e_ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            e2e_ms = float(tt.item())
        e2e = {"value": world * N / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms,
               "h2d_bytes_per_step": int(8 * ndof * (OX + 1)), "d2h_bytes_per_step": int(8 * (ndof + nnz)),
               "note": "mb_sweepx_assemble: pinned host state in, Llambda and nzval out; PCIe-bound (D2H of the CSC values)"}
        for a in X + [Lh, nzh]:
            eng.unpin(a)

    if rank == 0:
        flops = None
        try:
            with open(os.path.join(ROOT, "profiles", "flops.json")) as f:
                flops = json.load(f).get("flop_per_element", {}).get(str(OX))
        except Exception:
            pass
        roof = {"bound": "fp64", "kernel": "beam_kernel<ND=%d>" % (OX + 1), "unit": "TFLOP/s", "peak": fp64_peak,
                "peak_source": "measured live by mb_measure_fp64_tflops (DFMA loop); MEASURED_PEAKS.json has no FP64 figure",
                "kernel_ms": el_ms, "kernel_share_of_step": el_ms / (el_ms + ga_ms), "traffic": None}
        if flops:
            roof["flop_per_element"] = flops
            roof["achieved"] = N * flops / (el_ms * 1e-3) / 1e12
            roof["frac"] = roof["achieved"] / fp64_peak
        else:
            roof["achieved"] = None; roof["frac"] = None
        hbm_peak = None
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                hbm_peak = json.load(f)["hbm_gbs"]
        except Exception:
            hbm_peak = 6650.0
        roof["hbm"] = {"alg_bytes_per_element": alg_bytes_per_element(OX), "achieved_gbs": N * alg_bytes_per_element(OX) / (el_ms * 1e-3) / 1e9,
                       "peak_gbs": hbm_peak, "frac": N * alg_bytes_per_element(OX) / (el_ms * 1e-3) / 1e9 / hbm_peak}
        cpu_rate, cpu_dt = cpu_port_rate(mb, OX, args.cpu_sample, 1)
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": workload_config(args, OX, N), "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
                "roofline": roof,
                "cpu_baseline": {"value": cpu_rate, "unit": UNIT, "cores": 1, "kind": "port",
                                 "sample": "%d elements of the same chain, oracle literal restatement, 1 thread (%.1f s)" % (args.cpu_sample, cpu_dt)},
                "breakdown_ms": {"element_kernels": el_ms, "segmented_reduction": ga_ms, "wall_timed_region_s": wall}}
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    eng.close()


if __name__ == "__main__":
    main()
