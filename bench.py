#!/usr/bin/env python
"""bench.py — EulerBeam3D residual+Jacobian element-assemblies/s (BASELINE.json metric) on N B200s.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--nele E] [--ox 0|1|2] [--workload sweepx|directxua]
  torchrun … bench.py --gpus N …      (one rank per GPU; RANK/LOCAL_RANK/WORLD_SIZE/MASTER_* from the env)

Default workload (config.workload): BASELINE.json configs[2] — the inspect/PerformanceEulerBeam3D-style synthetic chain of 10M
EulerBeam3D elements per GPU, one SweepX `assemble!{:iter}` per step = element kernels + deterministic
reduction into Lλ and the CSC nzval.  `value`: state resident in HBM.  `e2e`: through the C-ABI call mb_sweepx_assemble with
pinned HOST buffers (H2D state, D2H Lλ + nzval inside the timed region).
N>1 (weak scaling): the chain has N·E elements, rank r owns elements [r·E,(r+1)·E); every step ends with the interface exchange
(78 doubles per cut, NCCL send/recv to the right neighbour INSIDE the shim: mb_comm_init / mb_iface_exchange — torch.distributed (gloo) only hands the
128-byte NCCL id round at start-up, nothing of it is on the timed path).
Every default line also carries a `directxua` block: BASELINE.json configs[3] (1e5 elements x 2000 time steps, time-sharded, strong scaling) measured in the
same run, so that the driver's 1→8 scaling record holds the DirectXUA curve too (--no-directxua skips it).
Inputs/outputs per step (≈1.8 GB read, ≈20 GB written/re-read per GPU) exceed the 126 MB L2: no L2 flush needed.

`--workload directxua` (BASELINE.json configs[3]): DirectXUA{2,0,0} assemblebig! of a 100k-element Udof beam chain over 2000 time steps, the
steps sharded over the ranks (strong scaling).  The materialised FP64 Lvv of 2000 steps (2.9e11 non-zeros) does not fit on 8 x 180 GB, so every rank
streams its steps through a window of 10 steps (1.4e9 non-zeros: 11.5 GB of values + 11.5 GB of row indices, handed to the solver window by window): mb_direct_rebase slides the
window without rebuilding the structure and only the new steps are evaluated.  `--workload directxua_weak`: 16 steps per GPU, one window, halo
L2[Λ,X] blocks of 2 steps exchanged with each neighbour over NCCL.

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


def alg_bytes_per_element(OX):
    """algorithmic HBM bytes per element-assembly of the element kernel: geometry line 128 + dof indices 48 +
    state gather 96·(OX+1) + element tangent 1152 + residual 96 (both written once)"""
    return 128 + 48 + 96 * (OX + 1) + 1152 + 96


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--workload", default="sweepx")
    ap.add_argument("--nele", type=float, default=None)
    ap.add_argument("--ox", type=int, default=0)
    ap.add_argument("--steps-per-gpu", type=int, default=16)      # directxua_weak
    ap.add_argument("--nstep", type=int, default=2000)            # directxua: time steps of the whole problem (BASELINE.json configs[3])
    ap.add_argument("--window", type=int, default=10)             # directxua: time steps per Lvv window
    ap.add_argument("--cpu-sample", type=int, default=8000)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--gauged", action="store_true")              # directxua: the beams carry 5 strain gauges each and a quadratic strain cost (ElementCost accelerator, windowed path)
    ap.add_argument("--no-directxua", action="store_true")        # skip the configs[3] block of the default line
    ap.add_argument("--dx-passes", type=int, default=2)           # timed passes of the configs[3] block (after 1 warm-up pass)
    ap.add_argument("--scr-nstep", type=int, default=4000)        # scr: time steps of BASELINE.json configs[4]
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

    def count(self):
        """samples written so far"""
        try:
            with open(self.f.name) as f:
                return sum(1 for line in f if line.count(",") >= 8)
        except OSError:
            return 0

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


def ensure_samples(sampler, comm, load):
    """nvidia-smi samples every 100 ms and needs a moment to start: when fewer than two samples fell inside a short timed region, repeat the same launches UNTIMED under the
    sampler (all ranks together) until it has them"""
    for _ in range(8):
        if comm.max(1. if (sampler is not None and sampler.count() < 2) else 0.) == 0.:
            break
        load()


def cpu_port_rate(mb, OX, nsample, nthreads):
    """oracle (literal restatement of the reference algorithm; the -O3 build with static-size duals, bit-identical to the checker build) on `nsample`
    elements of the same chain/state."""
    from oracle import elements as OE, pattern as OP
    eleobj, idx, ndof = mb.synthetic.chain(nsample, dynamic=OX > 0)
    X = mb.synthetic.state(ndof, nder=OX + 1)
    nm = mb.synthetic.newmark_coefficients(OX, 0.3)
    dis = [dict(X=idx, U=np.zeros((nsample, 0), np.int64), A=np.zeros((nsample, 0), np.int64))]
    a1, a2, cp, rv = OP.prepare_sweepx(dis, ndof, 0, 0)
    L = np.zeros(ndof); nz = np.zeros(len(rv))
    a1t, a2t = np.ascontiguousarray(a1[0].T), np.ascontiguousarray(a2[0].T)
    t = time.perf_counter()
    OE.sweepx_assemble_beams_mt(eleobj, idx, a1t, a2t, OX, X, np.ones(12), nm, L, nz, nthreads, static_duals=True)
    dt = time.perf_counter() - t
    return nsample / dt, dt


def workload_config(args, OX, nele, world, note=None):
    c = {"workload": "inspect/PerformanceEulerBeam3D-style synthetic chain, %d EulerBeam3D elements per GPU, SweepX{%d} assemble!{:iter} "
                     "(residual + 12x12 tangent per element, scatter into Llambda and CSC nzval)" % (nele, OX),
         "elements_per_gpu": int(nele), "OX": OX, "mission": "iter", "l2": "inputs/outputs larger than L2, no flush",
         "sharding": "none" if world == 1 else "element ranges, %d ranks, interface rows exchanged over NCCL (78 doubles per cut)" % world}
    if note:
        c["note"] = note
    return c


def run_reference(args, rank, world):
    """reference arm: the reference algorithm's CPU port on the host cores (rank 0 only)."""
    if rank != 0:
        return
    import muscade_b200 as mb
    from oracle import elements as OE
    if args.workload.startswith("directxua"):
        nsamp = max(50, args.cpu_sample // 4)
        cpu_port_rate_direct(mb, 50)
        times = [cpu_port_rate_direct(mb, nsamp)[1] for _ in range(args.steps)]
        value = nsamp * len(times) / sum(times)
        unit = "element-step assemblies/s"
        print(json.dumps({"impl": "reference", "metric": "EulerBeam3D residual+Jacobian element-step assemblies/s (DirectXUA{2,0,0} assemblebig!)", "value": value,
                          "unit": unit, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times),
                          "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                          "config": {"workload": "DirectXUA{2,0,0}, bounded sample: one time step of %d EulerBeam3D{Udof} elements per step, oracle port, 1 thread" % nsamp},
                          "cpu_baseline": {"value": value, "unit": unit, "cores": 1, "kind": "port", "sample": "%d element-steps x %d" % (nsamp, args.steps)},
                          "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}), flush=True)
        return
    OX = args.ox
    # all host cores this process may use; torchrun exports OMP_NUM_THREADS=1 for N>1, which must not shrink the reference arm (the other ranks idle)
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    if OE.max_threads() == 1 and os.environ.get("OMP_NUM_THREADS", "") not in ("", "1"):
        cores = 1                                   # built without OpenMP
    nsample = max(1000, int(args.cpu_sample) * max(1, cores) * 2)
    for _ in range(max(1, min(args.warmup, 1))):
        cpu_port_rate(mb, OX, 500, cores)
    times = []
    for _ in range(args.steps):
        _, dt = cpu_port_rate(mb, OX, nsample, cores)
        times.append(dt)
    value = nsample * len(times) / sum(times)
    N = int(args.nele or 1e7)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": workload_config(args, OX, N, 1, note="bounded sample (%d elements per step) of the same chain/state; element loop spread over all host "
                                      "threads (the reference's own loop src/Assemble.jl:479 is serial), serial scatter" % nsample),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": "%d elements x %d steps, C++ port of the algorithm (not Muscade/Julia), -O3, static-size duals (12 partials)" % (nsample, args.steps)},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def load_flops(OX):
    try:
        with open(os.path.join(ROOT, "profiles", "flops.json")) as f:
            return json.load(f).get("flop_per_element", {}).get(str(OX))
    except Exception:
        return None


def hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f)["hbm_gbs"], "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def load_falg(key):
    """frozen algorithmic flop counts (tools/opcount.cpp → profiles/falg.json, table in BASELINE.md §5)"""
    try:
        with open(os.path.join(ROOT, "profiles", "falg.json")) as f:
            return json.load(f)[key]
    except Exception:
        return None


class Comm:
    """One NCCL communicator per process, owned by a tiny engine that lives for the whole run (the shim's mb_comm_*); the 128-byte id travels over the
    gloo group torchrun's environment describes — start-up plumbing only."""

    def __init__(self, mb, rank, world, local, dist):
        self.rank, self.world, self.eng = rank, world, None
        if world > 1:
            import torch
            self.eng = mb.Engine(local)
            uid = torch.from_numpy(mb.Engine.comm_unique_id() if rank == 0 else np.zeros(128, np.uint8))
            dist.broadcast(uid, src=0)
            self.eng.comm_init(uid.numpy(), rank, world)

    def attach(self, eng):
        if self.eng is not None:
            eng.comm_init_from(self.eng)

    def max(self, x):
        return float(self.eng.comm_allreduce([x], "max")[0]) if self.eng is not None else float(x)

    def barrier(self):
        if self.eng is not None:
            self.eng.comm_barrier()

    def close(self):
        if self.eng is not None:
            self.eng.close(); self.eng = None


def run_sweepx(args, rank, world, local, comm, OX=None, N=None, steps=None, block=False):
    """block: a secondary workload carried inside the default line (no e2e / CPU legs): BASELINE.json configs[1] = SweepX{2} Newmark on 1e6 elements"""
    import muscade_b200 as mb
    OX = args.ox if OX is None else OX
    N = int(args.nele or 1e7) if N is None else int(N)
    steps = args.steps if steps is None else steps
    eng = mb.Engine(local)
    fp64_peak = eng.fp64_tflops()
    comm.attach(eng)
    eleobj, idx, ndof, dof0 = mb.sharding.chain_shard(N, rank, world, dynamic=OX > 0)
    ityp = eng.add_eulerbeam3d(eleobj, idx, np.ones(12))
    del eleobj, idx
    nnz = eng.sweepx_prepare(ndof)
    X = mb.synthetic.state(ndof, nder=OX + 1, offset=dof0)
    nm = mb.synthetic.newmark_coefficients(OX, 0.3)
    Lh = np.empty(ndof); nzh = np.empty(nnz)
    eng.sweepx_assemble(OX, "iter", X, nm, Llambda=Lh, nzval=nzh)       # uploads the state once: resident in HBM for `value`
    if world > 1:
        a2f = eng.sweepx_asm_range(ityp, 0, 1)[1][0]; a2l = eng.sweepx_asm_range(ityp, N - 1, N)[1][0]
        snz, sv, rnz, rvv = mb.sharding.interface_indices(a2l, a2f, ndof, rank, world)
        eng.iface_setup(snz, sv, rnz, rvv)

    def barrier():
        eng.sync(); comm.barrier()

    # a step = mb_sweepx_assemble_dev (+ mb_iface_exchange on a shard), exactly as a solver issues it; mb_sweepx_time_step_dev runs `reps` of them
    # between two CUDA events on the engine's stream
    # nvidia-smi samples every 100 ms and needs a moment to start, the timed region of 10 steps lasts 0.2 s: the sampler runs from the warm-up on (the same steps), and when
    # fewer than two samples fell inside, the same steps are repeated UNTIMED under it (all ranks together: a step may hold an exchange) until it has them
    sampler = ClockSampler(local) if rank == 0 else None
    eng.time_step_dev(OX, "iter", nm, reps=max(1, args.warmup))
    barrier()
    launches0 = eng.launch_count()
    t0 = time.perf_counter()
    step_ms = comm.max(eng.time_step_dev(OX, "iter", nm, reps=steps))
    barrier()
    wall = time.perf_counter() - t0
    launches = eng.launch_count() - launches0
    ensure_samples(sampler, comm, lambda: (eng.time_step_dev(OX, "iter", nm, reps=steps), barrier()))
    clocks = sampler.stop() if sampler else None
    if clocks is not None:
        clocks["window"] = "warm-up and timed steps (the same launches), plus untimed repeats of them when the region is shorter than two sampling periods"
    el_ms, ga_ms = eng.time_dev(OX, "iter", nm, reps=max(3, min(args.steps, 5)))      # same launch sets with an event between element kernels and reduction
    barrier()
    value = world * N / (step_ms * 1e-3)

    e2e = None
    if not args.no_e2e and not block:
        for a in X + [Lh, nzh]:
            eng.pin(a)

        def timed(fn, reps):
            for _ in range(2):
                fn()
            barrier()
            ts = []
            for _ in range(reps):
                t1 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t1)
            return comm.max(1e3 * float(np.mean(ts)))
        reps = max(3, min(args.steps, 5))
        # (1) what a solver with a DEVICE sparse solver does (north_star: nzval handed over behind mb_get_device_ptrs): host state in, Lλ out, nzval stays in HBM
        res_ms = timed(lambda: eng.sweepx_assemble(OX, "iter", X, nm, Llambda=Lh, nzval_on_device=True), reps)
        # (2) everything back to the host (a host sparse solver): + 8·nnz bytes of CSC values over PCIe, chunk-pipelined with the element kernels
        all_ms = timed(lambda: eng.sweepx_assemble(OX, "iter", X, nm, Llambda=Lh, nzval=nzh), reps)
        # the ceiling of (1): the same bytes (state in, Llambda out) copied both ways at once with no kernel in between, every rank at the same time
        barrier()
        copy_ms = comm.max(eng.host_copy_ms(X[0], Lh, reps=3)) * (OX + 1)
        e2e = {"value": world * N / (res_ms * 1e-3), "unit": UNIT, "ms_per_step": res_ms,
               "h2d_bytes_per_step": int(8 * ndof * (OX + 1)), "d2h_bytes_per_step": int(8 * ndof),
               "host_copy_only_ms": copy_ms,
               "host_copy_note": "H2D of the state and D2H of Llambda alone, concurrently, all ranks at once (mb_measure_host_copy_ms): what the host's PCIe / memory gives "
                                 "the %d GPU(s); the pipelined call cannot be faster than max(this, ms_per_step of the device-resident step)" % world,
               "note": "mb_sweepx_assemble through the C ABI with pinned HOST buffers, per rank: state.X in, Llambda out; the CSC values stay in HBM for the device "
                       "solver (mb_get_device_ptrs), as north_star specifies",
               "host_csc": {"value": world * N / (all_ms * 1e-3), "unit": UNIT, "ms_per_step": all_ms, "h2d_bytes_per_step": int(8 * ndof * (OX + 1)),
                            "d2h_bytes_per_step": int(8 * (ndof + nnz)),
                            "note": "the same call with nzval copied to the host as well (host sparse solver): PCIe-bound on the D2H of the CSC values"}}
        for a in X + [Lh, nzh]:
            eng.unpin(a)

    line = None
    if rank == 0:
        flops = load_flops(OX)
        falg = load_falg({0: "static", 1: "newmark_iter", 2: "newmark_iter"}[OX])
        kern = "beam_static_ap_kernel" if OX == 0 else "beam_cot_kernel<%d> + beam_kernel_sd<%d,split>" % (OX + 1, OX + 1)
        roof = {"bound": "fp64", "kernel": kern, "unit": "TFLOP/s", "peak": fp64_peak,
                "peak_source": "measured live by mb_measure_fp64_tflops (DFMA loop); MEASURED_PEAKS.json has no FP64 figure",
                "kernel_ms": el_ms, "kernel_share_of_step": el_ms / (el_ms + ga_ms), "kernel_ms_source": "CUDA events around the kernel's launches on the engine's stream",
                "traffic": None, "achieved": None, "frac": None}
        if flops:
            roof["flop_per_element"] = flops["flop"]
            if flops.get("dram_bytes_per_element"):      # measured DRAM bytes of one launch of the element kernel(s) (ncu --set full, profiles/)
                roof["traffic"] = int(N * flops["dram_bytes_per_element"]); roof["traffic_unit"] = "bytes per launch (dram read+write, ncu)"
            roof["achieved"] = N * flops["flop"] / (el_ms * 1e-3) / 1e12
            roof["frac"] = roof["achieved"] / fp64_peak
            roof["frac_executed"] = roof["frac"]
            roof["flop_source"] = "executed FP64 flops per element counted by ncu (profiles/flops.json): every lane of an element repeats the value sweep"
            roof["fp64_inst_per_element"] = flops.get("fp64_inst")
            if flops.get("fp64_inst"):
                # issue-slot view of the same pipe: DMUL/DADD occupy a DFMA slot but count one flop
                roof["fp64_pipe_frac"] = N * flops["fp64_inst"] * 2 / (el_ms * 1e-3) / 1e12 / fp64_peak
        if falg:
            # the frozen ALGORITHMIC count: forward-over-reverse mathematics per element with the value sweep counted once
            fa = falg["algorithmic"]["flop"]
            roof["alg_flop_per_element"] = fa
            roof["achieved_alg"] = N * fa / (el_ms * 1e-3) / 1e12
            roof["frac_alg"] = roof["achieved_alg"] / fp64_peak
            roof["alg_flop_source"] = "tools/opcount.cpp on the product's math headers (profiles/falg.json); +-*/sqrt = 1, fma = 2, libm call = 40"
        hp, hsrc = hbm_peak()
        ab = alg_bytes_per_element(OX)
        roof["hbm"] = {"alg_bytes_per_element": ab, "achieved_gbs": N * ab / (el_ms * 1e-3) / 1e9, "peak_gbs": hp, "peak_source": hsrc,
                       "frac": N * ab / (el_ms * 1e-3) / 1e9 / hp}
        cpu_n = args.cpu_sample * 5                       # ≈ 10 s of one host core with the static-dual build
        cpu_rate, cpu_dt = (None, 0.) if block else cpu_port_rate(mb, OX, cpu_n, 1)
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": args.warmup,
                "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": workload_config(args, OX, N, world), "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
                "roofline": roof,
                "cpu_baseline": {"value": cpu_rate, "unit": UNIT, "cores": 1, "kind": "port",
                                 "sample": "%d elements of the same chain, C++ port of the reference algorithm (not Muscade/Julia: no Julia in the image), -O3 with "
                                           "static-size duals, 1 thread as the reference's serial element loop (%.1f s)" % (cpu_n, cpu_dt)},
                "breakdown_ms": {"element_kernels": el_ms, "segmented_reduction": ga_ms, "serial_sum": el_ms + ga_ms, "wall_timed_region_s": wall,
                                 "note": "value is from K whole steps between two events; element_kernels / segmented_reduction are the same launches "
                                         "with an event in between (second pass)"}}
        if block:
            for k in ("cpu_baseline", "e2e", "higher_is_better", "vs_baseline", "dtype", "data", "clocks"):
                line.pop(k, None)
    eng.close()
    return line


def directxua_model(mb, N):
    model = mb.Model()
    nod = mb.addnode(model, np.arange(N + 1)[:, None] * np.array([.8, .6, 0.])[None, :])
    unod = mb.addnode(model, np.zeros((N, 0)))
    mat = mb.BeamCrossSection(EA=10., EI2=3., EI3=3., GJ=4., mu=1., iota1=1., Ca2=169.6, Ca3=169.6, Cq2=235.2, Cq3=235.2)
    mb.addelement(model, mb.EulerBeam3D, np.stack([nod[:-1], nod[1:], unod], axis=1), mat=mat, Udof=True)
    return model


def gauged_chain_model(mb, N):
    """the same chain with ElementCost{StrainGaugeOnEulerBeam3D} on every beam: 5 gauges (test/TestBeamElementStrainGauge.jl:10-11), quadratic strain cost"""
    P5 = np.array([[0., .5, 0.], [0., 0, .5], [0., -.5, 0.], [0., 0, -.5], [0., .5, 0.]]).T
    D5 = np.array([[1., 0., 0.], [1., 0., 0.], [1., 0., 0.], [1., 0., 0.], [1 / np.sqrt(2), 0, 1 / np.sqrt(2)]]).T
    model = mb.Model("gauged")
    nod = mb.addnode(model, np.arange(N + 1)[:, None] * np.array([.8, .6, 0.])[None, :])
    unod = mb.addnode(model, np.zeros((N, 0)))
    cost = mb.QuadraticGaugeCost(15e-6, lambda t: np.array([np.cos(t), 0., -np.cos(t), 0., np.cos(t) / 2]) * 0.001)
    mat = mb.BeamCrossSection(EA=10., EI2=3., EI3=3., GJ=4., mu=1., iota1=1., Ca2=169.6, Ca3=169.6, Cq2=235.2, Cq3=235.2)
    mb.addelement(model, mb.ElementCost, np.stack([nod[:-1], nod[1:], unod], axis=1), req=("ε",), cost=cost, ElementType=mb.StrainGaugeOnEulerBeam3D,
                  elementkwargs=dict(P=P5, D=D5, elementkwargs=dict(mat=mat, Udof=True)))
    return model, cost


def cpu_port_rate_direct(mb, nsample):
    """oracle (literal reference algorithm, DirectXUA.jl:85-120 with Np = 39 partials) on one time step of `nsample` elements, 1 thread"""
    from oracle import elements as OE, pattern as OP
    model = directxua_model(mb, nsample)
    st0 = mb.initialize(model); dis = st0.dis
    nX, nU, nA = model.getndof(("X", "U", "A"))
    P = OP.prepare_direct([dict(X=d.X, U=d.U, A=d.A) for d in dis.dis], nX, nU, nA, 2, 0, 0)
    X = [mb.synthetic.uniform_pm1(10 + d, nX) * (0.05 if d == 0 else 0.1) for d in range(3)]; U = mb.synthetic.uniform_pm1(99, nU)
    t = time.perf_counter()
    OE.direct_assemble_step_beams(model.ele[0].eleobj, dis.dis[0].X, dis.dis[0].U, 2, 0, X, [U], dis.dis[0].scaleX, dis.dis[0].scaleU, P, 0)
    dt = time.perf_counter() - t
    return nsample / dt, dt


def run_directxua(args, rank, world, local, comm, steps=None, warmup=None, block=False):
    """BASELINE.json configs[3]: DirectXUA{2,0,0} assemblebig! of N Udof beams over `nstep` time steps; rank r owns steps [r·nstep/world,(r+1)·nstep/world)
    (strong scaling) and streams them through a sliding window of `window` steps.  One bench step = one pass over all time steps."""
    import torch
    import muscade_b200 as mb
    OX, OU = 2, 0
    steps = args.steps if steps is None else steps
    warmup = args.warmup if warmup is None else warmup
    N = int(1e5 if block or not args.nele else args.nele)
    nstep = args.nstep
    L, H, Wn, windows, interior = mb.sharding.directxua_windows(nstep, rank, world, args.window)
    torch.cuda.set_device(local)
    stream = torch.cuda.current_stream().cuda_stream
    gauged = bool(getattr(args, "gauged", False)) and not block
    gcost = None
    if gauged:
        model, gcost = gauged_chain_model(mb, N)
    else:
        model = directxua_model(mb, N)
    st0 = mb.initialize(model)
    nX, nU = model.getndof("X"), model.getndof("U")
    dtm = 0.1
    engines = {}

    def make(lo, hi):
        e = mb.directxua.prepare(OX, OU, model, st0.dis, nstep, dtm, lo, hi, device=local)
        e.set_stream(stream)
        return e
    for w in windows:
        if w not in interior:
            engines[w] = make(*w)                      # contains the first or the last step: one-sided stencils, its own structure
    e_int = make(*interior[0]) if interior else None
    # synthetic states: a bank of B states resident in HBM, step s reads bank[s % B]; a pinned host copy feeds the e2e pass
    B = 32
    hostX = [[mb.synthetic.uniform_pm1(10 + 3 * b + d, nX) * (0.05 if d == 0 else 0.1) for d in range(3)] for b in range(B)]
    hostU = [mb.synthetic.uniform_pm1(99 + b, nU) for b in range(B)]
    devX = [[torch.from_numpy(x).cuda() for x in xs] for xs in hostX]
    devU = [torch.from_numpy(u).cuda() for u in hostU]
    nset = [0]

    if gauged:                                         # Λ of every step and the measured strains of that step (5 numbers) go in with the state
        import ctypes as C
        hostL = [mb.synthetic.uniform_pm1(200 + b, nX) for b in range(B)]
        devL = [torch.from_numpy(x).cuda() for x in hostL]

    def set_dev(e, s):
        e.set_state_dev(s, [x.data_ptr() for x in devX[s % B]], devU[s % B].data_ptr()); nset[0] += 1
        if gauged:
            mb._lib.check(e.h, e.L.mb_direct_set_lambda(e.h, int(s), C.c_void_p(devL[s % B].data_ptr())))
            e.set_gauge_measurements(s, 1, gcost.measured(dtm * s))

    def set_host(e, s):
        e.set_state(s, hostX[s % B], hostU[s % B]); nset[0] += 1
        if gauged:
            e.set_lambda(s, hostL[s % B]); e.set_gauge_measurements(s, 1, gcost.measured(dtm * s))

    def one_pass(setter, Lv_host=None):
        """every window of this rank once: new states in, new steps evaluated, owned Lvv columns and Lv rows built"""
        kept = (0, 0)
        for w in windows:
            if w in engines:
                e = engines[w]; kept_w = (0, 0)
            else:
                e = e_int
                if (e.lo, e.hi) != w:
                    old = e.stored_range()
                    e.rebase(w[0])
                    new = e.stored_range()
                    kept_w = (max(old[0], new[0]), min(old[1], new[1])) if e_int_valid[0] else (0, 0)
                else:
                    kept_w = (0, 0)
                e_int_valid[0] = True
            a, b = e.stored_range()
            todo = [s for s in range(a, b) if not (kept_w[0] <= s < kept_w[1])]
            for s in todo:
                setter(e, s)
            if todo:
                e.direct_assemble(eval_range=(todo[0], todo[-1] + 1), build_big=False)      # the new steps are one contiguous run
            e.direct_assemble(eval_range=(w[0], w[0]), build_big=True, Lv=Lv_host)

    e_int_valid = [False]

    def barrier():
        torch.cuda.synchronize()
        comm.barrier()

    for _ in range(warmup):
        e_int_valid[0] = False
        one_pass(set_dev)
    barrier()
    all_eng = list(engines.values()) + ([e_int] if e_int else [])
    launches0 = sum(e.launch_count() for e in all_eng)
    sampler = ClockSampler(local) if rank == 0 else None
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(steps):
        e_int_valid[0] = False                         # every pass starts from scratch: nothing kept from the previous pass
        one_pass(set_dev)
    ev1.record(); torch.cuda.synchronize()
    step_ms = comm.max(ev0.elapsed_time(ev1) / steps)
    barrier()
    launches = sum(e.launch_count() for e in all_eng) - launches0
    clocks = sampler.stop() if sampler else None
    # kernel breakdown on one window (CUDA events inside the engine)
    et = e_int or all_eng[0]
    a_ms, b_ms = et.direct_time(reps=2)
    el_ms = et.last_element_ms
    # e2e: the same pass through the C ABI with HOST buffers — states from pinned host memory, Lv rows back to the host; Lvv stays in HBM for the
    # device solver (mb_get_device_ptrs), as it would in production: 29 GB per window over PCIe would only measure the bus
    e2e = None
    if not args.no_e2e and not block:
        Lvh = np.empty(et.ncol)
        for arrs in hostX:
            for x in arrs: et.pin(x)
        for u in hostU: et.pin(u)
        if gauged:
            for x in hostL: et.pin(x)
        et.pin(Lvh)
        e_int_valid[0] = False; one_pass(set_host, Lvh)
        barrier()
        n0 = nset[0]
        t1 = time.perf_counter()
        e_int_valid[0] = False; one_pass(set_host, Lvh)
        torch.cuda.synchronize()
        e2e_ms = 1e3 * (time.perf_counter() - t1)
        nset_pass = nset[0] - n0
        e2e_ms = comm.max(e2e_ms)
        e2e = {"value": N * nstep / (e2e_ms * 1e-3), "unit": "element-step assemblies/s", "ms_per_step": e2e_ms,
               "h2d_bytes_per_step": int(nset_pass * 8 * (3 * nX + nU + ((nX + 5) if gauged else 0))), "d2h_bytes_per_step": int(len(windows) * 8 * et.ncol),
               "note": "per rank: mb_direct_set_state from pinned host memory for every newly stored step, mb_direct_assemble with a host Lv; "
                       "Lvv stays in HBM (device-pointer hand-off to the solver)"}
    line = None
    if rank == 0:
        flops = load_flops("directxua")
        falg = load_falg("directxua_200_udof")
        nown = et.hi - et.lo
        roof = {"bound": "fp64", "kernel": "beam_direct_cot_kernel<3,false> + <3,true> + beam_direct_b0_kernel<3> + beam_direct_lin_kernel<3>", "unit": "TFLOP/s",
                "peak": et.fp64_tflops(), "peak_source": "measured live by mb_measure_fp64_tflops (DFMA loop); MEASURED_PEAKS.json has no FP64 figure",
                "kernel_ms": el_ms, "kernel_ms_source": "CUDA events around the element kernels of one window (%d steps, 4 launches each) on the engine's stream" % nown,
                "kernel_share_of_step": el_ms / (a_ms + b_ms), "traffic": None, "achieved": None, "frac": None}
        if flops:
            roof["flop_per_element"] = flops["flop"]; roof["fp64_inst_per_element"] = flops.get("fp64_inst")
            roof["achieved"] = N * nown * flops["flop"] / (el_ms * 1e-3) / 1e12
            roof["frac"] = roof["achieved"] / roof["peak"]
            if flops.get("fp64_inst"):
                roof["fp64_pipe_frac"] = N * nown * flops["fp64_inst"] * 2 / (el_ms * 1e-3) / 1e12 / roof["peak"]
            roof["frac_executed"] = roof["frac"]
            if flops.get("dram_bytes_per_element"):
                roof["traffic"] = int(N * flops["dram_bytes_per_element"]); roof["traffic_unit"] = "bytes per time step (dram read+write of the four launches, ncu)"
        if falg:
            roof["alg_flop_per_element"] = falg["algorithmic"]["flop"]
            roof["achieved_alg"] = N * nown * falg["algorithmic"]["flop"] / (el_ms * 1e-3) / 1e12
            roof["frac_alg"] = roof["achieved_alg"] / roof["peak"]
        nsamp = max(50, args.cpu_sample // 4)
        cpu_rate, cpu_dt = (None, 0.) if block else cpu_port_rate_direct(mb, nsamp)
        line = {"metric": "EulerBeam3D residual+Jacobian element-step assemblies/s (DirectXUA{2,0,0} assemblebig!)", "value": N * nstep / (step_ms * 1e-3),
                "unit": "element-step assemblies/s", "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": step_ms, "ms_per_pass": step_ms,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": "DirectXUA{2,0,0} load identification, %d EulerBeam3D{Udof} elements x %d time steps (BASELINE.json configs[3]), time steps "
                                       "sharded over %d GPU(s), each rank streaming its %d steps through a sliding window of %d steps (Lvv columns of the window "
                                       "built on the device)" % (N, nstep, world, H - L, Wn),
                           "elements": N, "nstep": nstep, "steps_per_gpu": H - L, "window": Wn, "lvv_nnz_per_window": int(et.nnzbig),
                           "states": "bank of %d synthetic states resident in HBM, step s reads bank[s mod %d]" % (B, B),
                           "l2": "per-step blocks and Lvv columns far larger than L2, no flush",
                           "halo": "the 2+2 halo steps at a rank's edges are evaluated locally (no data-path collective)"},
                "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roof,
                "cpu_baseline": {"value": cpu_rate, "unit": "element-step assemblies/s", "cores": 1, "kind": "port",
                                 "sample": "one time step of %d elements, oracle literal restatement (39 partials), 1 thread (%.1f s)" % (nsamp, cpu_dt)},
                "breakdown_ms": {"window_elements_and_step_blocks": a_ms, "window_lvv_build": b_ms, "window_element_kernels": el_ms, "windows_per_rank": len(windows)}}
        if gauged:
            line["metric"] = "element-step assemblies/s (DirectXUA{2,0,0} assemblebig!, ElementCost{StrainGaugeOnEulerBeam3D} on every beam, windowed path)"
            line["config"]["workload"] = line["config"]["workload"].replace("load identification,", "load identification from strain gauges (5 gauges per beam, quadratic strain cost: the "
                                                                            "ElementCost accelerator in the windowed path, mb_direct_set_gauge_cost),")
            line["config"]["states"] += "; Λ from a bank of the same size, measured strains set per step"
            line["roofline"]["note"] = "flop counts are those of the plain beam kernels (the 3 strain-gauge launches per step are not counted): kernel_ms covers all 6 element launches per step"
            line["cpu_baseline"] = None
        if block:
            for k in ("cpu_baseline", "e2e", "higher_is_better", "vs_baseline", "dtype", "data", "clocks"):
                line.pop(k, None)
            line["steps_per_gpu"] = H - L
    for e in all_eng:
        e.close()
    del devX, devU
    torch.cuda.empty_cache()
    return line


def run_directxua_weak(args, rank, world, local, comm):
    """DirectXUA{2,0,0} assemblebig! with the time steps sharded over the ranks (weak: steps_per_gpu each), one resident window per rank and the halo
    blocks exchanged over NCCL inside the shim (mb_direct_halo_exchange) — the data-path collective of SURVEY.md §8e, kept as a bench of its own."""
    import muscade_b200 as mb
    OX, OU = 2, 0
    N = int(args.nele or 1e5)
    S = args.steps_per_gpu
    nstep = S * world
    lo, hi = rank * S, (rank + 1) * S
    model = directxua_model(mb, N)
    st0 = mb.initialize(model)
    nX, nU = model.getndof("X"), model.getndof("U")
    eng = mb.directxua.prepare(OX, OU, model, st0.dis, nstep, 0.1, lo, hi, device=local)
    comm.attach(eng)
    for s in range(lo, hi):                                # own steps only: the halo blocks come from the neighbours
        eng.set_state(s, [mb.synthetic.uniform_pm1(10 + 3 * s + d, nX) * (0.05 if d == 0 else 0.1) for d in range(3)], mb.synthetic.uniform_pm1(99 + s, nU))

    def step_dev():
        eng.direct_assemble(eval_range=(lo, hi), build_big=False)
        if world > 1:
            eng.halo_exchange()
        eng.direct_assemble(eval_range=(lo, lo), build_big=True)

    def barrier():
        eng.sync(); comm.barrier()

    for _ in range(args.warmup):
        step_dev()
    barrier()
    launches0 = eng.launch_count()
    t0 = time.perf_counter()
    if world == 1:
        a, b = eng.direct_time(reps=args.steps)
        step_ms = a + b
    else:
        for _ in range(args.steps):
            step_dev()
        eng.sync()
        step_ms = comm.max(1e3 * (time.perf_counter() - t0) / args.steps)      # host clock around device-synchronised steps (the engine owns its stream)
        a, b = eng.direct_time(reps=1)
    barrier()
    launches = eng.launch_count() - launches0
    line = None
    if rank == 0:
        line = {"metric": "EulerBeam3D residual+Jacobian element-step assemblies/s (DirectXUA{2,0,0} assemblebig!)", "value": world * N * S / (step_ms * 1e-3),
                "unit": "element-step assemblies/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_ms,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": "DirectXUA{2,0,0} load identification, %d EulerBeam3D{Udof} elements x %d time steps per GPU, time-sharded; "
                                       "Lvv columns of the owned steps built on the device" % (N, S),
                           "elements": N, "steps_per_gpu": S, "nstep": nstep, "lvv_nnz_per_gpu": int(eng.nnzbig),
                           "halo": "L2[Lambda,X] of 2 steps per neighbour over NCCL send/recv inside the shim (mb_direct_halo_exchange)" if world > 1 else "none"},
                "gpu_launches": int(launches), "breakdown_ms": {"elements_and_step_blocks": a, "lvv_build": b}}
    eng.close()
    return line


def run_scr(args, rank, world, local, comm, block=False):
    """BASELINE.json configs[4]: DirectXUA{2,0,0} assemblebig! of the SCR riser (examples/DynamicBeamAnalysis.jl:62-135: 100 EulerBeam3D{Udof} + 61 SoilContact on
    the device, Hold / DofConstraint / DofLoad / U-costs host-evaluated) over `--scr-nstep` time steps (4000), time-sharded over the ranks with locally evaluated
    halos.  The whole series fits one handle per rank (≈ 5·10⁴ Lvv non-zeros per step); the steps share launch sets (step batching).  One bench step = one pass."""
    import muscade_b200 as mb
    OX, OU, nstep, dt, t0, su = 2, 0, args.scr_nstep, 0.1, 0., 50.
    gauged = bool(getattr(args, "gauged", False)) and not block
    gcost = mb.QuadraticGaugeCost(2e-5, lambda t: 1e-4 * np.cos(0.5 * t) * np.array([1., 0.5, -1., -0.5])) if gauged else None
    model, node_lists, weights = mb.examples.scr_riser(mb, udof=True, gauge_cost=gcost)
    unodes = np.concatenate([et.nodID[:, 2] for et in model.ele if et.ElType.__name__ in ("EulerBeam3D", "ElementCost")])
    for f in ("t1", "t2", "t3"):
        mb.addelement(model, mb.SingleDofCost, unodes[:, None], clas="U", field=f, cost=lambda u, t: 0.5 * (u / su) ** 2)
    mb.setscale(model, scale=dict(X=dict(t1=2., t2=2., t3=2.), U=dict(t1=30., t2=30., t3=30.)), Λscale=1e3)
    st0 = mb.initialize(model); dis = st0.dis
    nX, nU = model.getndof("X"), model.getndof("U")
    nel_dev = sum(et.nele for et in model.ele if et.ElType.kind in ("eulerbeam3d", "soilcontact", "bar3d", "elementcost"))
    if nstep % world or nstep // world < 6:
        raise ValueError("the number of time steps must be a multiple of the number of ranks, at least 6 each")
    lo, hi = rank * nstep // world, (rank + 1) * nstep // world
    import torch
    torch.cuda.set_device(local)
    eng = mb.directxua.prepare(OX, OU, model, dis, nstep, dt, lo, hi, device=local, t0=t0)
    eng.set_stream(torch.cuda.current_stream().cuda_stream)          # CUDA events of the timed region are recorded on this stream
    a, b = eng.stored_range()
    B = 16                                            # bank of synthetic states (amplitudes of the parity test), step s reads bank[s mod B]
    bankX = [[mb.synthetic.uniform_pm1(10 + 3 * k + d, nX) * (0.2 if d == 0 else 0.3) for d in range(3)] for k in range(B)]
    bankU = [20. * mb.synthetic.uniform_pm1(99 + k, nU) for k in range(B)]
    bankL = [mb.synthetic.uniform_pm1(500 + k, nX) for k in range(B)]

    def upload():
        for s in range(a, b):
            eng.set_state(s, bankX[s % B], bankU[s % B])
    upload()
    for s in range(a, b):                             # Λ and the host-evaluated types: the host's share of an iteration, set once (not device work)
        eng.set_lambda(s, bankL[s % B])
    if gauged:
        eng.set_gauge_times(t0 + dt * np.arange(nstep))
    for s in range(a, min(b, a + B)):                 # host contributions of the first B steps, reused round-robin through the bank (same states ⇒ same values up to t)
        eng.set_host_cost(s, *mb.directxua.host_costs(eng, s, bankX[s % B][0], bankU[s % B], t0 + s * dt)[:4])
        mb.directxua.host_elements(eng, s, bankX[s % B], bankL[s % B], t0 + s * dt, model.scaleΛ)

    def barrier():
        eng.sync(); torch.cuda.synchronize(); comm.barrier()

    steps = args.dx_passes if block else args.steps
    warm = 1 if block else args.warmup
    for _ in range(warm):
        eng.direct_assemble()
    barrier()
    launches0 = eng.launch_count()
    sampler = ClockSampler(local) if rank == 0 and not block else None
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(steps):
        eng.direct_assemble()                         # every stored step (owned + the 2+2 halo steps) evaluated, owned Lvv columns / Lv rows built
    ev1.record(); torch.cuda.synchronize()
    step_ms = comm.max(ev0.elapsed_time(ev1) / steps)
    a_ms, b_ms = eng.direct_time(reps=1)              # breakdown of the owned steps (CUDA events inside the engine)
    barrier()
    launches = eng.launch_count() - launches0
    ensure_samples(sampler, comm, lambda: ([eng.direct_assemble() for _ in range(steps)], barrier()))
    clocks = sampler.stop() if sampler else None
    el_ms = eng.last_element_ms
    e2e = None
    if not args.no_e2e and not block:
        Lvh = np.empty(eng.ncol)
        for arrs in bankX:
            for x in arrs: eng.pin(x)
        for u in bankU: eng.pin(u)
        eng.pin(Lvh)

        def one():
            upload(); eng.direct_assemble(Lv=Lvh)
        one(); barrier()
        t1 = time.perf_counter(); one(); eng.sync()
        e2e_ms = comm.max(1e3 * (time.perf_counter() - t1))
        e2e = {"value": nel_dev * nstep / (e2e_ms * 1e-3), "unit": "element-step assemblies/s", "ms_per_step": e2e_ms,
               "h2d_bytes_per_step": int((b - a) * 8 * (3 * nX + nU)), "d2h_bytes_per_step": int(8 * eng.ncol),
               "note": "per rank: mb_direct_set_state for every stored step from pinned host memory, mb_direct_assemble with a host Lv; Lvv stays in HBM"}
    line = None
    if rank == 0:
        nown = hi - lo
        line = {"metric": "element-step assemblies/s (DirectXUA{2,0,0} assemblebig!, SCR riser: EulerBeam3D{Udof} + SoilContact)", "value": nel_dev * nstep / (step_ms * 1e-3),
                "unit": "element-step assemblies/s", "n_gpus": world, "steps": steps, "warmup": warm, "ms_per_step": step_ms, "ms_per_pass": step_ms,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": "DirectXUA{2,0,0} XU-optimisation set-up of the SCR riser (BASELINE.json configs[4]; IA = 0): %d device elements per step (100 EulerBeam3D{Udof} in "
                                       "5 types + 61 SoilContact), 107 host-evaluated elements (Hold, DofConstraint, DofLoad) and 300 U-costs merged, x %d time steps, "
                                       "%d per GPU, step batching" % (nel_dev, nstep, nown),
                           "device_elements_per_step": nel_dev, "nstep": nstep, "steps_per_gpu": nown, "lvv_nnz_per_gpu": int(eng.nnzbig), "ndofX": nX, "ndofU": nU,
                           "l2": "per-step blocks and Lvv (%.1f GB per GPU) larger than L2" % (16e-9 * eng.nnzbig)},
                "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
                "breakdown_ms": {"elements_and_step_blocks": a_ms, "lvv_build": b_ms, "element_kernels": el_ms}}
        if gauged:
            line["metric"] = "element-step assemblies/s (DirectXUA{2,0,0} assemblebig!, SCR riser with ElementCost{StrainGaugeOnEulerBeam3D} on every beam + SoilContact)"
            line["config"]["workload"] = line["config"]["workload"].replace("(100 EulerBeam3D{Udof} in", "(100 EulerBeam3D{Udof}, each with 4 strain gauges and a quadratic strain cost — "
                                                                            "the ElementCost accelerator in the windowed path, mb_direct_set_gauge_cost —, in")
        if block:
            for k in ("e2e", "higher_is_better", "vs_baseline", "dtype", "data", "clocks"):
                line.pop(k, None)
    eng.close()
    return line


def run_gauges(args, rank, world, local, comm):
    """The GENERAL DirectXUA form (mb_xua_*) with DEVICE element types and the ElementCost accelerator: N strain-gauged EulerBeam3D{Udof} (ElementCost{StrainGaugeOnEulerBeam3D},
    quadratic strain cost, 5 gauges per element) x 6 time steps, DirectXUA{2,0,0}.  One bench step = one assemblebig! pass: per step the first-order beam kernels, the packet and
    strain-gauge kernels, the segmented reductions into out and the weighted additions into Lvv / Lv.  Replicas only (the general form is not sharded): every rank runs the same problem."""
    import ctypes as C
    import muscade_b200 as mb
    from muscade_b200 import xua
    from muscade_b200._lib import check
    N = int(args.nele) if args.nele else 100000
    OX, nstep, dt = 2, 6, 0.1
    P5 = np.array([[0., .5, 0.], [0., 0, .5], [0., -.5, 0.], [0., 0, -.5], [0., .5, 0.]]).T          # test/TestBeamElementStrainGauge.jl:10-11
    D5 = np.array([[1., 0., 0.], [1., 0., 0.], [1., 0., 0.], [1., 0., 0.], [1 / np.sqrt(2), 0, 1 / np.sqrt(2)]]).T
    m = mb.Model("gauged")
    nod = mb.addnode(m, np.stack([np.arange(N + 1, dtype=float), np.zeros(N + 1), np.zeros(N + 1)], axis=1))
    un = mb.addnode(m, np.zeros((N, 3)))
    cost = mb.QuadraticGaugeCost(15e-6, lambda t: np.array([np.cos(t), 0., -np.cos(t), 0., np.cos(t) / 2]) * 0.001)
    mb.addelement(m, mb.ElementCost, np.stack([nod[:-1], nod[1:], un], axis=1), req=("ε",), cost=cost, ElementType=mb.StrainGaugeOnEulerBeam3D,
                  elementkwargs=dict(P=P5, D=D5, elementkwargs=dict(mat=mb.BeamCrossSection(EA=1e4, EI2=300., EI3=300., GJ=400., mu=1., iota1=1.), orient2=(0., 1., 0.), Udof=True)))
    s0 = mb.initialize(m); dis = s0.dis
    nX, nU, _ = m.getndof(("X", "U", "A"))
    eng = xua.XUAEngine(local)
    nbig, nnz = eng.prepare(m, dis, OX, 0, 0, [nstep], [dt])
    st = s0.with_orders(1, OX + 1, 1)
    for k in range(nstep):
        s = mb.State(dt * k, [mb.synthetic.uniform_pm1(7 + k, nX)], [0.01 * mb.synthetic.uniform_pm1(20 + 3 * k + d, nX) for d in range(OX + 1)],
                     [0.1 * mb.synthetic.uniform_pm1(90 + k, nU)], st.A, None, m, dis)
        eng.put_state(1, k + 1, s)
    em = np.ascontiguousarray(cost.measured(0.))
    check(eng.h, eng.L.mb_xua_set_gauge_measurements(eng.h, 1, em.ctypes.data_as(C.c_void_p), 0))
    ms = np.zeros(1, np.float32)
    check(eng.h, eng.L.mb_xua_time_device_pass(eng.h, max(1, args.warmup), ms))
    comm.barrier()
    launches0 = eng.launch_count()
    sampler = ClockSampler(local) if rank == 0 else None
    check(eng.h, eng.L.mb_xua_time_device_pass(eng.h, args.steps, ms))      # CUDA events on the engine's stream around `steps` whole passes (after one more warm-up pass)
    step_ms = comm.max(float(ms[0]))
    launches = (eng.launch_count() - launches0) * args.steps // (args.steps + 1)
    ensure_samples(sampler, comm, lambda: check(eng.h, eng.L.mb_xua_time_device_pass(eng.h, args.steps, ms)))
    clocks = sampler.stop() if sampler else None
    line = None
    if rank == 0:
        Np = 12 + 12 * (OX + 1) + 3
        npd = 12 * (OX + 1) + 3
        nxx = 108 * N + 36                                                                   # non-zeros of the X-X class pattern of the chain (= Λ-X, X-Λ); Λ-U / U-Λ: 36 N
        out_nz = (2 * (OX + 1) + 1) * nxx + 2 * 36 * N                                       # live blocks of out: L2[Λ,X_d], L2[X_d,Λ], L2[X₀,X₀], L2[Λ,U], L2[U,Λ]
        add_nz = (2 * 6 + 1) * nxx + 2 * 36 * N                                              # block entries added into Lvv per step: 1+2+3 stencil points per derivative order
        bytes_step = N * 8 * 12 * (npd + 1) * 2 + N * 8 * (Np + 12 + 144) * 2 + out_nz * 16 + add_nz * 24      # (R,dR), ∇L / cost terms, out written + read, Lvv read + written + Lvvasm
        line = {"metric": "element-step assemblies/s (DirectXUA{2,0,0} assemblebig!, general form, ElementCost{StrainGaugeOnEulerBeam3D} on the device)",
                "value": N * nstep / (step_ms * 1e-3), "unit": "element-step assemblies/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": step_ms, "ms_per_pass": step_ms, "higher_is_better": True, "scaling": "replicas", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": "general DirectXUA{2,0,0} form (mb_xua_*): %d strain-gauged EulerBeam3D{Udof} (5 gauges each, quadratic strain cost: the ElementCost accelerator on "
                                       "the device) x %d time steps; no host-evaluated type, states resident in HBM" % (N, nstep),
                           "elements": N, "nstep": nstep, "lvv_size": int(nbig), "lvv_nnz": int(nnz), "l2": "element outputs, out blocks (%.1f GB per step) and Lvv (%.1f GB) larger than L2" % (8e-9 * out_nz, 8e-9 * nnz)},
                "clocks": clocks, "gpu_launches": int(launches),
                "roofline": {"bound": "hbm", "unit": "GB/s", "achieved": bytes_step * nstep / (step_ms * 1e-3) / 1e9, "peak": hbm_peak()[0], "peak_source": hbm_peak()[1],
                             "frac": bytes_step * nstep / (step_ms * 1e-3) / 1e9 / hbm_peak()[0], "traffic": None,
                             "bytes_per_step": int(bytes_step),
                             "note": "whole pass against the HBM peak (MEASURED_PEAKS.json); algorithmic bytes: element outputs (R, dR, gradient and cost terms) written and read, "
                                     "the live blocks of out written by the segmented reductions and read by the weighted adds, Lvv values read + written with their Lvvasm "
                                     "positions.  Device element types are reduced from the outputs of their kernels: no dense %d x %d packet is formed" % (Np, Np)}}
    eng.close()
    return line


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
        dist.init_process_group("gloo")              # start-up plumbing only (the NCCL id); the data path and the timing reductions are NCCL inside the shim
    comm = Comm(mb, rank, world, local, dist)
    if args.workload == "directxua":
        line = run_directxua(args, rank, world, local, comm)
    elif args.workload == "directxua_weak":
        line = run_directxua_weak(args, rank, world, local, comm)
    elif args.workload == "scr":
        line = run_scr(args, rank, world, local, comm)
    elif args.workload == "gauges":
        line = run_gauges(args, rank, world, local, comm)
    else:
        line = run_sweepx(args, rank, world, local, comm)
        if not args.no_directxua and args.nele is None and args.ox == 0:
            # BASELINE.json configs[3] in the same run (strong scaling over the ranks): 1 warm-up + dx_passes timed passes over all 2000 time steps
            dx = run_directxua(args, rank, world, local, comm, steps=args.dx_passes, warmup=1, block=True)
            # BASELINE.json configs[1]: SweepX{2} Newmark-β assemble!{:iter} on 1e6 elements per GPU (the per-step assembly of a 1000-step run)
            nw = run_sweepx(args, rank, world, local, comm, OX=2, N=1e6, steps=50, block=True)
            # BASELINE.json configs[4]: DirectXUA on the SCR riser, 4000 time steps (few elements, many steps: step batching)
            sc = run_scr(args, rank, world, local, comm, block=True)
            if line is not None:
                line["directxua"] = dx
                line["newmark"] = nw
                line["scr"] = sc
    if rank == 0 and line is not None:
        line["comm"] = {"backend": "NCCL inside libmuscade_b200.so (mb_comm_init, dlopen libnccl.so.2)" if world > 1 else "none", "ranks": world,
                        "nccl_version": comm.eng.comm_info()[2] if comm.eng is not None else None,
                        "torch_distributed_on_timed_path": False}
        print(json.dumps(line), flush=True)
    comm.barrier()
    comm.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
