#!/usr/bin/env python
"""bench.py — EulerBeam3D residual+Jacobian element-assemblies/s (BASELINE.json metric) on N B200s.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--nele E] [--ox 0|1|2] [--workload sweepx|directxua]
  torchrun … bench.py --gpus N …      (one rank per GPU; RANK/LOCAL_RANK/WORLD_SIZE/MASTER_* from the env)

Default workload (config.workload): BASELINE.json configs[2] — the inspect/PerformanceEulerBeam3D-style synthetic chain of 10M
EulerBeam3D elements per GPU, one SweepX `assemble!{:iter}` per step = element kernels + deterministic
reduction into Lλ and the CSC nzval.  `value`: state resident in HBM.  `e2e`: through the C-ABI call mb_sweepx_assemble with
pinned HOST buffers (H2D state, D2H Lλ + nzval inside the timed region).
N>1 (weak scaling): the chain has N·E elements, rank r owns elements [r·E,(r+1)·E); every step ends with the interface exchange
(78 doubles per cut, NCCL send/recv to the right neighbour) — sharding.py.
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
    nsample = max(1000, int(args.cpu_sample) * max(1, cores // 2))
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
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": "%d elements x %d steps" % (nsample, args.steps)},
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


def run_sweepx(args, rank, world, local, dist):
    import muscade_b200 as mb
    OX = args.ox
    N = int(args.nele or 1e7)
    eng = mb.Engine(local)
    fp64_peak = eng.fp64_tflops()
    torch = None
    if world > 1:
        import torch
        eng.set_stream(torch.cuda.current_stream().cuda_stream)
    eleobj, idx, ndof, dof0 = mb.sharding.chain_shard(N, rank, world, dynamic=OX > 0)
    ityp = eng.add_eulerbeam3d(eleobj, idx, np.ones(12))
    del eleobj, idx
    nnz = eng.sweepx_prepare(ndof)
    X = mb.synthetic.state(ndof, nder=OX + 1, offset=dof0)
    nm = mb.synthetic.newmark_coefficients(OX, 0.3)
    Lh = np.empty(ndof); nzh = np.empty(nnz)
    eng.sweepx_assemble(OX, "iter", X, nm, Llambda=Lh, nzval=nzh)       # uploads the state once: resident in HBM for `value`
    sendbuf = recvbuf = None
    if world > 1:
        a2f = eng.sweepx_asm_range(ityp, 0, 1)[1][0]; a2l = eng.sweepx_asm_range(ityp, N - 1, N)[1][0]
        snz, sv, rnz, rvv = mb.sharding.interface_indices(a2l, a2f, ndof, rank, world)
        eng.iface_setup(snz, sv, rnz, rvv)
        sendbuf = torch.zeros(max(1, len(snz) + len(sv)), dtype=torch.float64, device="cuda")
        recvbuf = torch.zeros(max(1, len(rnz) + len(rvv)), dtype=torch.float64, device="cuda")

    def step_dev():
        eng.sweepx_assemble_dev(OX, "iter", nm)
        if world > 1:
            eng.iface_pack(sendbuf.data_ptr())
            mb.sharding.exchange_neighbours(dist, sendbuf, recvbuf, rank, world)
            eng.iface_unpack_add(recvbuf.data_ptr())

    def barrier():
        eng.sync()
        if dist is not None:
            torch.cuda.synchronize(); dist.barrier()

    for _ in range(args.warmup):
        step_dev()
    barrier()
    launches0 = eng.launch_count()
    sampler = ClockSampler(local) if rank == 0 else None
    t0 = time.perf_counter()
    if world == 1:
        # K whole steps exactly as a solver issues them (mb_sweepx_assemble_dev), CUDA events on the engine's stream; then the same launch sets
        # with an event between element kernels and reduction for the per-kernel breakdown / roofline
        step_ms = eng.time_step_dev(OX, "iter", nm, reps=args.steps)
        el_ms, ga_ms = eng.time_dev(OX, "iter", nm, reps=max(3, min(args.steps, 5)))
    else:
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(args.steps):
            step_dev()
        ev1.record()
        torch.cuda.synchronize()
        step_ms = ev0.elapsed_time(ev1) / args.steps
        el_ms, ga_ms = eng.time_dev(OX, "iter", nm, reps=2)
        tt = torch.tensor([step_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        step_ms = float(tt.item())
    barrier()
    wall = time.perf_counter() - t0
    launches = eng.launch_count() - launches0
    clocks = sampler.stop() if sampler else None
    value = world * N / (step_ms * 1e-3)

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
        if world > 1:
            tt = torch.tensor([e2e_ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            e2e_ms = float(tt.item())
        e2e = {"value": world * N / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms,
               "h2d_bytes_per_step": int(8 * ndof * (OX + 1)), "d2h_bytes_per_step": int(8 * (ndof + nnz)),
               "note": "mb_sweepx_assemble: pinned host state in, Llambda and nzval out (per rank); PCIe-bound (D2H of the CSC values)"}
        for a in X + [Lh, nzh]:
            eng.unpin(a)

    if rank == 0:
        flops = load_flops(OX)
        roof = {"bound": "fp64", "kernel": "beam_static_sym_kernel" if OX == 0 else "beam_cot_kernel<%d> + beam_kernel_sd<%d,split>" % (OX + 1, OX + 1), "unit": "TFLOP/s", "peak": fp64_peak,
                "peak_source": "measured live by mb_measure_fp64_tflops (DFMA loop); MEASURED_PEAKS.json has no FP64 figure",
                "kernel_ms": el_ms, "kernel_share_of_step": el_ms / (el_ms + ga_ms), "kernel_ms_source": "CUDA events around the kernel's launches on the engine's stream", "traffic": None, "achieved": None, "frac": None}
        if flops:
            roof["flop_per_element"] = flops["flop"]
            if flops.get("dram_bytes_per_element"):      # measured DRAM bytes of one launch of the element kernel(s) (ncu --set full, profiles/)
                roof["traffic"] = int(N * flops["dram_bytes_per_element"]); roof["traffic_unit"] = "bytes per launch (dram read+write, ncu)"
            roof["achieved"] = N * flops["flop"] / (el_ms * 1e-3) / 1e12
            roof["frac"] = roof["achieved"] / fp64_peak
            roof["fp64_inst_per_element"] = flops.get("fp64_inst")
            if flops.get("fp64_inst"):
                # issue-slot view of the same pipe: DMUL/DADD occupy a DFMA slot but count one flop
                roof["fp64_pipe_frac"] = N * flops["fp64_inst"] * 2 / (el_ms * 1e-3) / 1e12 / fp64_peak
        # SURVEY.md §8d's provisional algorithmic count (structure-exploiting evaluation of the same mathematics, before any kernel existed): the
        # forward-over-reverse kernels need a fifth of it, so `achieved`/`frac` above use the EXECUTED flops; this block only places the kernel against the
        # ceiling the survey derived from its own figure (FP64 peak / F_alg elements/s; its 60 % target was 2.2e8 el/s)
        f_alg = {0: 1.0e5, 1: 1.5e5, 2: 1.5e5}[OX]
        roof["survey_f_alg"] = {"flop_per_element": f_alg, "ceiling_elements_per_s": fp64_peak * 1e12 / f_alg,
                                "kernel_elements_per_s": N / (el_ms * 1e-3), "kernel_over_ceiling": N / (el_ms * 1e-3) / (fp64_peak * 1e12 / f_alg)}
        hp, hsrc = hbm_peak()
        ab = alg_bytes_per_element(OX)
        roof["hbm"] = {"alg_bytes_per_element": ab, "achieved_gbs": N * ab / (el_ms * 1e-3) / 1e9, "peak_gbs": hp, "peak_source": hsrc,
                       "frac": N * ab / (el_ms * 1e-3) / 1e9 / hp}
        cpu_rate, cpu_dt = cpu_port_rate(mb, OX, args.cpu_sample, 1)
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": workload_config(args, OX, N, world), "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
                "roofline": roof,
                "cpu_baseline": {"value": cpu_rate, "unit": UNIT, "cores": 1, "kind": "port",
                                 "sample": "%d elements of the same chain, oracle literal restatement, 1 thread (%.1f s)" % (args.cpu_sample, cpu_dt)},
                "breakdown_ms": {"element_kernels": el_ms, "segmented_reduction": ga_ms, "serial_sum": el_ms + ga_ms, "wall_timed_region_s": wall,
                                 "note": "value is from K whole steps between two events; element_kernels / segmented_reduction are the same launches "
                                         "with an event in between (second pass)"}}
        print(json.dumps(line), flush=True)
    eng.close()


def directxua_model(mb, N):
    model = mb.Model()
    nod = mb.addnode(model, np.arange(N + 1)[:, None] * np.array([.8, .6, 0.])[None, :])
    unod = mb.addnode(model, np.zeros((N, 0)))
    mat = mb.BeamCrossSection(EA=10., EI2=3., EI3=3., GJ=4., mu=1., iota1=1., Ca2=169.6, Ca3=169.6, Cq2=235.2, Cq3=235.2)
    mb.addelement(model, mb.EulerBeam3D, np.stack([nod[:-1], nod[1:], unod], axis=1), mat=mat, Udof=True)
    return model


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


def run_directxua(args, rank, world, local, dist):
    """BASELINE.json configs[3]: DirectXUA{2,0,0} assemblebig! of N Udof beams over `nstep` time steps; rank r owns steps [r·nstep/world,(r+1)·nstep/world)
    (strong scaling) and streams them through a sliding window of `window` steps.  One bench step = one pass over all time steps."""
    import torch
    import muscade_b200 as mb
    OX, OU = 2, 0
    N = int(args.nele or 1e5)
    nstep = args.nstep
    L, H, Wn, windows, interior = mb.sharding.directxua_windows(nstep, rank, world, args.window)
    torch.cuda.set_device(local)
    stream = torch.cuda.current_stream().cuda_stream
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

    def set_dev(e, s):
        e.set_state_dev(s, [x.data_ptr() for x in devX[s % B]], devU[s % B].data_ptr()); nset[0] += 1

    def set_host(e, s):
        e.set_state(s, hostX[s % B], hostU[s % B]); nset[0] += 1

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
        if dist is not None:
            dist.barrier()

    for _ in range(args.warmup):
        e_int_valid[0] = False
        one_pass(set_dev)
    barrier()
    all_eng = list(engines.values()) + ([e_int] if e_int else [])
    launches0 = sum(e.launch_count() for e in all_eng)
    sampler = ClockSampler(local) if rank == 0 else None
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        e_int_valid[0] = False                         # every pass starts from scratch: nothing kept from the previous pass
        one_pass(set_dev)
    ev1.record(); torch.cuda.synchronize()
    step_ms = ev0.elapsed_time(ev1) / args.steps
    if dist is not None:
        tt = torch.tensor([step_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        step_ms = float(tt.item())
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
    if not args.no_e2e:
        Lvh = np.empty(et.ncol)
        for arrs in hostX:
            for x in arrs: et.pin(x)
        for u in hostU: et.pin(u)
        et.pin(Lvh)
        e_int_valid[0] = False; one_pass(set_host, Lvh)
        barrier()
        n0 = nset[0]
        t1 = time.perf_counter()
        e_int_valid[0] = False; one_pass(set_host, Lvh)
        torch.cuda.synchronize()
        e2e_ms = 1e3 * (time.perf_counter() - t1)
        nset_pass = nset[0] - n0
        if dist is not None:
            tt = torch.tensor([e2e_ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            e2e_ms = float(tt.item())
        e2e = {"value": N * nstep / (e2e_ms * 1e-3), "unit": "element-step assemblies/s", "ms_per_step": e2e_ms,
               "h2d_bytes_per_step": int(nset_pass * 8 * (3 * nX + nU)), "d2h_bytes_per_step": int(len(windows) * 8 * et.ncol),
               "note": "per rank: mb_direct_set_state from pinned host memory for every newly stored step, mb_direct_assemble with a host Lv; "
                       "Lvv stays in HBM (device-pointer hand-off to the solver)"}
    if rank == 0:
        flops = load_flops("directxua")
        nown = et.hi - et.lo
        roof = {"bound": "fp64", "kernel": "beam_direct_cot_kernel<3> + beam_direct_b0_kernel<3> + beam_direct_lin_kernel<3>", "unit": "TFLOP/s",
                "peak": et.fp64_tflops(), "peak_source": "measured live by mb_measure_fp64_tflops (DFMA loop); MEASURED_PEAKS.json has no FP64 figure",
                "kernel_ms": el_ms, "kernel_ms_source": "CUDA events around the element kernels of one window (%d steps, 3 launches each) on the engine's stream" % nown,
                "kernel_share_of_step": el_ms / (a_ms + b_ms), "traffic": None, "achieved": None, "frac": None}
        if flops:
            roof["flop_per_element"] = flops["flop"]; roof["fp64_inst_per_element"] = flops.get("fp64_inst")
            roof["achieved"] = N * nown * flops["flop"] / (el_ms * 1e-3) / 1e12
            roof["frac"] = roof["achieved"] / roof["peak"]
            if flops.get("fp64_inst"):
                roof["fp64_pipe_frac"] = N * nown * flops["fp64_inst"] * 2 / (el_ms * 1e-3) / 1e12 / roof["peak"]
            if flops.get("dram_bytes_per_element"):
                roof["traffic"] = int(N * flops["dram_bytes_per_element"]); roof["traffic_unit"] = "bytes per time step (dram read+write of the three launches, ncu)"
        nsamp = max(50, args.cpu_sample // 4)
        cpu_rate, cpu_dt = cpu_port_rate_direct(mb, nsamp)
        line = {"metric": "EulerBeam3D residual+Jacobian element-step assemblies/s (DirectXUA{2,0,0} assemblebig!)", "value": N * nstep / (step_ms * 1e-3),
                "unit": "element-step assemblies/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_ms,
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
        print(json.dumps(line), flush=True)
    for e in all_eng:
        e.close()


def run_directxua_weak(args, rank, world, local, dist):
    """DirectXUA{2,0,0} assemblebig! with the time steps sharded over the ranks (weak: steps_per_gpu each)."""
    import muscade_b200 as mb
    OX, OU = 2, 0
    N = int(args.nele or 1e5)
    S = args.steps_per_gpu
    nstep = S * world
    lo, hi = rank * S, (rank + 1) * S
    torch = None
    model = mb.Model()
    nod = mb.addnode(model, np.arange(N + 1)[:, None] * np.array([.8, .6, 0.])[None, :])
    unod = mb.addnode(model, np.zeros((N, 0)))
    mat = mb.BeamCrossSection(EA=10., EI2=3., EI3=3., GJ=4., mu=1., iota1=1., Ca2=169.6, Ca3=169.6, Cq2=235.2, Cq3=235.2)
    mb.addelement(model, mb.EulerBeam3D, np.stack([nod[:-1], nod[1:], unod], axis=1), mat=mat, Udof=True)
    st0 = mb.initialize(model)
    nX, nU = model.getndof("X"), model.getndof("U")
    eng = mb.directxua.prepare(OX, OU, model, st0.dis, nstep, 0.1, lo, hi, device=local)
    if world > 1:
        import torch
        eng.set_stream(torch.cuda.current_stream().cuda_stream)
    for s in range(max(0, lo - 2), min(nstep, hi + 2)):
        eng.set_state(s, [mb.synthetic.uniform_pm1(10 + 3 * s + d, nX) * (0.05 if d == 0 else 0.1) for d in range(3)], mb.synthetic.uniform_pm1(99 + s, nU))
    plan = mb.sharding.directxua_halo_plan(nstep, lo, hi)

    def view(steps):
        p, n = eng.step_ptrs(steps[0])[0]
        return torch.as_tensor(mb.sharding.CudaView(p, n * len(steps)), device="cuda")

    bufs = {}
    if world > 1:
        for k in ("send_left", "send_right", "recv_left", "recv_right"):
            if plan[k]:
                bufs[k] = view(plan[k])

    def step_dev():
        eng.direct_assemble(eval_range=(lo, hi), build_big=False)
        if world > 1:
            ops = []
            if "send_right" in bufs: ops.append(dist.P2POp(dist.isend, bufs["send_right"], rank + 1))
            if "send_left" in bufs: ops.append(dist.P2POp(dist.isend, bufs["send_left"], rank - 1))
            if "recv_left" in bufs: ops.append(dist.P2POp(dist.irecv, bufs["recv_left"], rank - 1))
            if "recv_right" in bufs: ops.append(dist.P2POp(dist.irecv, bufs["recv_right"], rank + 1))
            for r in dist.batch_isend_irecv(ops):
                r.wait()
        eng.direct_assemble(eval_range=(lo, lo), build_big=True)

    def barrier():
        eng.sync()
        if dist is not None:
            torch.cuda.synchronize(); dist.barrier()

    for _ in range(args.warmup):
        step_dev()
    barrier()
    launches0 = eng.launch_count()
    if world == 1:
        a, b = eng.direct_time(reps=args.steps)
        step_ms = a + b
    else:
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(args.steps):
            step_dev()
        ev1.record(); torch.cuda.synchronize()
        step_ms = ev0.elapsed_time(ev1) / args.steps
        tt = torch.tensor([step_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        step_ms = float(tt.item())
        a, b = eng.direct_time(reps=1)
    barrier()
    launches = eng.launch_count() - launches0
    if rank == 0:
        line = {"metric": "EulerBeam3D residual+Jacobian element-step assemblies/s (DirectXUA{2,0,0} assemblebig!)", "value": world * N * S / (step_ms * 1e-3),
                "unit": "element-step assemblies/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_ms,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": "DirectXUA{2,0,0} load identification, %d EulerBeam3D{Udof} elements x %d time steps per GPU, time-sharded; "
                                       "Lvv columns of the owned steps built on the device" % (N, S),
                           "elements": N, "steps_per_gpu": S, "nstep": nstep, "lvv_nnz_per_gpu": int(eng.nnzbig),
                           "halo": "L2[Lambda,X] of 2 steps per neighbour over NCCL send/recv" if world > 1 else "none"},
                "gpu_launches": int(launches), "breakdown_ms": {"elements_and_step_blocks": a, "lvv_build": b}}
        print(json.dumps(line), flush=True)
    eng.close()


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if args.workload == "directxua":
        run_directxua(args, rank, world, local, dist)
    elif args.workload == "directxua_weak":
        run_directxua_weak(args, rank, world, local, dist)
    else:
        run_sweepx(args, rank, world, local, dist)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
