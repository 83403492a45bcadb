#!/usr/bin/env python
"""Benchmark of the tfQMR hot path on B200 (driver contract: see the task statement).

Workload (BASELINE.json configs[2], the largest single-GPU configuration the metric is quoted on):
    synthetic 3-D 27-point block-stencil BSR operator, 32x32 complex fp32 blocks, 32^3 = 32768 block rows,
    64 right-hand-side columns (2 block columns of 32), unit right-hand sides, sigma = 8, tolerance 1e-3 (see DEFAULT_TOL).
A "step" is one complete tfQMR solve of that system through the C-ABI of libtfQMRgpu.so.

  value : whole-job solve throughput in GFLOP/s, flops counted with the reference's own convention
          (tfqmrgpu_bsrsv_getInfo.flops_performed), A and B already resident in HBM;
  e2e   : same metric through setMatrix(A) + setMatrix(B) + solve + getMatrix(X) with pinned HOST buffers,
          host<->device copies inside the timed region;
  roofline : the block-sparse product kernel (the dominant kernel), algorithmic bytes per launch over the
          average launch duration measured with CUDA events on the solver's stream inside the timed steps;
  cpu_baseline : the unmodified reference CPU build (oracle/_ref) - or the oracle port when that is not
          available - on a bounded sample (6^3 block rows, same blocks / RHS / tolerance), N=1 rank 0 only.
          The same baseline leg also times, when oracle/_ref holds it, the reference's own CUDA kernels (unmodified sources
          compiled for sm_100) on the full workload on this GPU and reports them as `reference_gpu` - a reported baseline
          like cpu_baseline, after the timed region, never part of the product path (`--no-cpu` skips both).

N > 1 (torchrun): every rank solves its own 64 right-hand-side columns of ONE 64*N-column problem with A replicated
(RHS block-column sharding); the ranks keep the reference's global iteration / probe rule by all-gathering three numbers
per rank twice per iteration (NCCL); value = all ranks' flops / max time (weak scaling).  `strong`: one problem of
--strong-rhs columns split over the ranks.  A reaches the GPUs in N ranges over N PCIe links + NCCL broadcasts.
`--impl reference` times the reference's CPU path on the bounded sample instead (rank 0 only).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "tfqmr_solve_throughput"
UNIT = "GFLOP/s"
SAMPLE_N = 6
# fp32 tfQMR on this system cannot go below a relative residual of ~7e-5 (measured on B200: 7.7e-5 with the SIMT product and
# with the reference's own CUDA kernels, 6.4e-5 with the tensor-core product), and some right-hand-side shards stall just
# above 1e-4.  The benchmark tolerance therefore is 1e-3: every shard converges in 7 iterations to ~1.8e-4.
DEFAULT_TOL = 1e-3


def workload_name(n, lm, ln, ncols, prec, tol=1e-3, sigma=8.0):
    return (f"stencil27 n={n}^3={n**3} block rows, {lm}x{ln} complex {'fp64' if prec == 'z' else 'fp32'} blocks, "
            f"{ncols*ln} RHS columns per GPU, sigma={sigma:g}, tol={tol:g}")


class Quiet:
    """Silence C-level stdout chatter of the reference library while it runs."""
    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        self.null = os.open(os.devnull, os.O_WRONLY)
        os.dup2(self.null, 1)

    def __exit__(self, *a):
        os.dup2(self.saved, 1)
        os.close(self.saved); os.close(self.null)


def cpu_reference_sample(lm, ln, ncols, prec, tol, maxit, repeats=1):
    """The reference's own CPU implementation (oracle/_ref/libtfqmr_ref_cpu.so, HAS_NO_CUDA build of the
    unmodified sources) on the bounded sample; falls back to the oracle port when that .so is absent."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import orclib as O
    from tfqmrgpu_b200 import problems as P
    dt = np.float64 if prec == "z" else np.float32
    d = P.stencil27(SAMPLE_N, lm, ln, ncols, sigma=8.0, dtype=dt)
    ref = O.ref_cpu()
    times, flops, its = [], 0.0, 0
    kind = "reference" if ref is not None else "port"
    for _ in range(repeats):
        if ref is not None:
            with Quiet():
                r = ref.solve(d["mb"], lm, ln, d["rpA"], d["ciA"], d["valA"].reshape(-1), d["rpX"], d["ciX"],
                              d["rpB"], d["ciB"], d["valB"].reshape(-1), tol, maxit, prec)
            times.append(r["t_solve"]); flops = r["flops"]; its = r["iterations"]
        else:
            op = O.OraclePlan(d["mb"], d["rpA"], d["ciA"], d["rpX"], d["ciX"], d["rpB"], d["ciB"])
            A_int = O.import_blocks(d["valA"].reshape(-1), d["nnzbA"], lm, lm, var="A")
            B_int = O.import_blocks(d["valB"].reshape(-1), d["nnzbB"], lm, ln, var="B")
            v3 = O.v3_glibc(op.nnzbX*2*lm*ln)
            t0 = time.perf_counter()
            o = O.solve(op, lm, ln, A_int, B_int, v3, tol, maxit)
            times.append(time.perf_counter() - t0); flops = o["flops"]; its = o["iterations"]
    return dict(kind=kind, times=times, flops=flops, iterations=its,
                sample=f"stencil27 n={SAMPLE_N}^3={SAMPLE_N**3} block rows (full size: 32^3), same {lm}x{ln} blocks, "
                       f"{ncols*ln} RHS, tol {tol:g}: one complete solve ({its} iterations, {flops*1e-9:.1f} GFLOP)")


def reference_gpu_same_box(sp, lm, ln, prec, tol, maxit):
    """Informational: the reference's OWN CUDA kernels (unmodified sources, nvcc -arch=sm_100, oracle/_ref) on this GPU,
    same full-size workload, one warm-up + one timed solve.  Not the contract's reference arm (that is the CPU path)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import orclib as O
    ref = O.ref_gpu()
    if ref is None:
        return None
    vA = sp.valA_host.numpy().reshape(-1); vB = sp.valB.reshape(-1)
    out = None
    for _ in range(2):
        with Quiet():
            out = ref.solve(sp.mb, lm, ln, sp.rpA, sp.ciA, vA, sp.rpX, sp.ciX, sp.rpB, sp.ciB, vB, tol, maxit, prec)
    return {"value": out["flops"]/out["t_solve"]*1e-9, "unit": UNIT, "ms_per_solve": 1e3*out["t_solve"], "iterations": out["iterations"],
            "createPlan_ms": 1e3*out.get("t_plan", float("nan")),
            "status": int(out["status"]), "residual": out["residuum"],
            "what": "unmodified reference CUDA kernels (gemmNxNf etc.) recompiled for sm_100, same GPU, same workload"}


class ClockSampler(threading.Thread):
    """nvidia-smi clocks and throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.proc = index, [], None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        self.join(timeout=2)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for nme, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                continue
        if not sm:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=[], samples=0)
        return dict(sm_mhz=float(np.median(sm)), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    lm, ln, ncols, prec, tol, maxit = 32, 32, 2, "c", DEFAULT_TOL, 100
    cpu_reference_sample(lm, ln, ncols, prec, tol, maxit, repeats=args.warmup)
    r = cpu_reference_sample(lm, ln, ncols, prec, tol, maxit, repeats=args.steps)
    t = float(np.sum(r["times"]))
    value = r["flops"]*args.steps/t*1e-9
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3*t/args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args.n, lm, ln, ncols, prec, tol),
                   "note": "reference CPU path (HAS_NO_CUDA build, serial solver) timed on a bounded sample of the workload"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": r["kind"], "sample": r["sample"],
                         "host_cores_available": os.cpu_count()},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def _make_plan(torch, api, sp, lm, ln, prec, dev, stream=None, shard=None):
    """Plan + caller-owned workspace (a torch tensor, so that windows of it can be NCCL buffers).  shard = (dist, rank, world,
    ncols_global): register the exchange that keeps the reference's GLOBAL iteration / probe rule across the ranks."""
    from tfqmrgpu_b200 import sharded
    h = api.Handle(stream if stream is not None else torch.cuda.current_stream(dev).cuda_stream)
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    pl = api.BsrsvPlan(h, sp.mb, sp.rpA, sp.ciA, sp.rpX, sp.ciX, sp.rpB, sp.ciB)
    torch.cuda.synchronize(dev)
    pl.create_plan_ms = 1e3*(time.perf_counter() - t0)          # plan analysis on the device (createPlan), host -> host
    keep = None
    if shard is not None:
        dist, rank, world, ncols_global = shard[:4]
        group = shard[4] if len(shard) > 4 else None
        pl.set_shard_hints(0, ncols_global)
        keep = sharded.NcclExchange(pl, dist, rank, world, ncols_global*ln, dev, group=group)
    t0 = time.perf_counter()
    nbytes = pl.buffer_size_for(lm, ln, prec)
    torch.cuda.synchronize(dev)
    pl.configure_ms = 1e3*(time.perf_counter() - t0)            # tiles, units, CTA schedule (bufferSize)
    ws_t = torch.empty(nbytes + 256, dtype=torch.uint8, device=dev)
    ws_ptr = (ws_t.data_ptr() + 255) & ~255
    pl.set_buffer(ws_ptr, keep_alive=(ws_t, keep))
    base = ws_ptr - ws_t.data_ptr()
    return h, pl, ws_t, base, nbytes


def run_ours(args):
    import torch
    import torch.distributed as dist
    from tfqmrgpu_b200 import api, synthetic, _lib as L

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n, lm, ln, ncols, prec, tol, maxit = args.n, args.lm, args.ln, args.ncols, args.precision, args.tol, 100
    dt = np.float64 if prec == "z" else np.float32
    es = 8 if prec == "z" else 4

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def reduce_max(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX); return float(t.item())

    def reduce_sum(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev); dist.all_reduce(t, op=dist.ReduceOp.SUM); return float(t.item())

    # this rank's shard: block columns [rank*ncols, (rank+1)*ncols) of ONE world*ncols-column problem, A replicated; the ranks
    # keep the reference's global iteration / probe rule through the exchange (3 numbers per rank, twice per iteration)
    ncols_global = world*ncols
    sp = synthetic.Stencil27(n, lm, ln, ncols, sigma=args.sigma, dtype=dt, device=dev, col0=rank*ncols, ncols_global=ncols_global,
                             with_values=False)
    # the exchange of the convergence monitors gets its own communicator: the broadcasts that distribute the NEXT system's operator
    # (default group, other stream) are issued before a solve and would otherwise be ahead of its all-gathers in the same queue
    ex_group = dist.new_group(backend="nccl") if world > 1 else None
    shard = (dist, rank, world, ncols_global, ex_group) if world > 1 else None
    h, pl, ws_t, base, nbytes = _make_plan(torch, api, sp, lm, ln, prec, dev, shard=shard)
    info = pl.plan_info()
    create_plan_ms, configure_ms = pl.create_plan_ms, pl.configure_ms
    # every rank uploads ITS row range of A over its own PCIe link (setMatrixPart) and the ranks exchange the converted ranges
    # device to device; the host values of that range only are generated (and pinned) here
    parts = [pl.matrix_part_info(r, world) for r in range(world)]
    mine = parts[rank]
    valA_part = sp.values_of(mine["block0"], mine["block0"] + mine["nblocks"])
    valB = torch.from_numpy(sp.valB).pin_memory()
    x_host = torch.empty(sp.nnzbX*lm*ln*2, dtype=torch.float64 if prec == "z" else torch.float32).pin_memory()
    x_np = x_host.numpy()

    def upload(pl=pl, ws_t=ws_t, base=base, stream=None, after=None):
        """A (this rank's row range, then the exchange of the converted ranges) and B of one system, asynchronously on the plan's
        stream.  after: an event of the previous upload - two host->device copies in flight on two streams SHARE the PCIe link
        and both finish late (measured: two 7.25 GB uploads enqueued together both end after 272 ms instead of 131 and 262), so
        a pipelined caller chains its uploads.  Returns the event that marks this upload done."""
        s_up = stream if stream is not None else torch.cuda.current_stream(dev)
        if after is not None:
            s_up.wait_event(after)
        pl.set_matrix_part(valA_part.data_ptr(), rank, world)
        if world > 1:
            with torch.cuda.stream(stream if stream is not None else torch.cuda.current_stream(dev)):
                for r, q in enumerate(parts):      # N broadcasts of 1/N each = an all-gather of ranges of (slightly) different sizes
                    if q["length"]:
                        dist.broadcast(ws_t[base + q["off"]:base + q["off"] + q["length"]], src=r)
                    if q["scale_length"]:
                        dist.broadcast(ws_t[base + q["scale_off"]:base + q["scale_off"] + q["scale_length"]], src=r)
        pl.set_matrix("B", None, "n", raw_ptr=valB.data_ptr())
        done = torch.cuda.Event()
        done.record(s_up)
        return done

    upload()
    statuses = []
    for _ in range(max(args.warmup, 3)):
        statuses.append(pl.solve(tol, maxit))
    # ---- device-resident metric: K solves, inputs already in HBM; the production path (iteration bodies as CUDA graphs on one
    #      GPU, exchange hook on several), no profiling events ----------------------------------------------------------------
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start(); time.sleep(0.3)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    flops = launches = iters = 0
    barrier()
    e0.record()
    for _ in range(args.steps):
        statuses.append(pl.solve(tol, maxit))
        st_ = pl.solve_stats()
        flops += pl.info()["flops"]; launches += st_["launches"]; iters += pl.info()["iterations"]
    e1.record()
    barrier()
    ms = reduce_max(e0.elapsed_time(e1))
    total_flops = reduce_sum(flops)
    worst_status = int(reduce_max(float(max(statuses[-args.steps:]))))
    iters_max = reduce_max(iters/args.steps); iters_min = -reduce_max(-iters/args.steps)
    last = pl.info()
    # ---- the same K solves once more with CUDA events around every block-sparse product (the roofline's launch durations) -----
    pl.set_profiling(True)
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    spmm_ms = spmm_n = 0
    barrier()
    p0.record()
    for _ in range(args.steps):
        pl.solve(tol, maxit)
        prof = pl.solve_profile()
        spmm_ms += prof["spmm_ms"]; spmm_n += prof["spmm_launches"]
    p1.record()
    barrier()
    clocks = sampler.stop() if sampler else None
    ms_profiled = reduce_max(p0.elapsed_time(p1))
    pl.set_profiling(False)

    # ---- end to end through the C-ABI with host buffers -------------------------------------------------------
    h2d = reduce_sum(float(valA_part.numel()*valA_part.element_size() + valB.numel()*valB.element_size()))   # all ranks together
    d2h = reduce_sum(float(x_host.numel()*x_host.element_size()))
    # Double-buffered, as a caller with a stream of systems would run it: a second plan with its own workspace and stream
    # takes the upload of step k+1 (asynchronous C-ABI calls: pinned H2D, conversion on its stream) while step k is being
    # solved.  Every step's A, B go host -> device and its X device -> host inside the timed region.
    # If the second workspace cannot be allocated on any rank (a smaller GPU), every rank falls back to one plan: same
    # per-step copies, nothing overlapped.
    lanes = [(pl, ws_t, base, None, x_np)]
    pl2 = h2 = None
    create_plan_warm_ms = configure_warm_ms = None
    try:
        if os.environ.get("TFQMRGPU_BENCH_NO_PIPELINE"):
            raise RuntimeError("disabled by TFQMRGPU_BENCH_NO_PIPELINE")
        st2 = torch.cuda.Stream(dev)
        h2, pl2, ws2_t, base2, nbytes2 = _make_plan(torch, api, sp, lm, ln, prec, dev, stream=st2.cuda_stream, shard=shard)
        assert nbytes2 == nbytes
        create_plan_warm_ms, configure_warm_ms = pl2.create_plan_ms, pl2.configure_ms       # second analysis of the same problem in this process
        x2_host = torch.empty_like(x_host).pin_memory()
        lanes.append((pl2, ws2_t, base2, st2, x2_host.numpy()))
    except Exception as exc:                                  # noqa: BLE001 - any allocation failure means "no second lane"
        print(f"bench: second plan for the double-buffered e2e leg not available ({exc!r}); e2e runs unpipelined", file=sys.stderr)
    if reduce_max(float(len(lanes) < 2)) > 0:                 # all ranks together (the upload holds collectives)
        lanes = lanes[:1]
    pipelined = len(lanes) > 1
    if pipelined:
        upload(*lanes[1][:4]); pl2.solve(tol, maxit)        # untimed warm-up of the second plan (graph capture, first touch)
    e2e_flops = 0
    timeline = [] if os.environ.get("TFQMRGPU_BENCH_TIMELINE") else None      # dev: when did every upload finish on the device?

    def mark(lane_stream, what):
        if timeline is not None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record(lane_stream if lane_stream is not None else torch.cuda.current_stream(dev))
            timeline.append((what, ev, 1e3*(time.perf_counter() - t0)))
    barrier()
    base_ev = torch.cuda.Event(enable_timing=True); base_ev.record()
    t0 = time.perf_counter()
    last_up = upload(*lanes[0][:4]); mark(lanes[0][3], "upload 0 done")
    for k in range(args.steps):
        if pipelined and k + 1 < args.steps:
            last_up = upload(*lanes[(k + 1) % 2][:4], after=last_up); mark(lanes[(k + 1) % 2][3], f"upload {k + 1} done")
        elif not pipelined and k > 0:
            upload(*lanes[0][:4])
        cur, _, _, lane_stream, xout = lanes[k % len(lanes)]
        cur.solve(tol, maxit); mark(lane_stream, f"solve {k} done")
        cur.get_matrix("X", "n", L.LAYOUT_RIRIRIRI, out=xout); mark(lane_stream, f"download {k} done")
        e2e_flops += cur.info()["flops"]
    torch.cuda.synchronize(dev)
    e2e_s = reduce_max(time.perf_counter() - t0)
    if timeline is not None and rank == 0:
        for what, ev, host_ms in timeline:
            print(f"timeline: {what:18s} device {base_ev.elapsed_time(ev):8.1f} ms   (host enqueued at {host_ms:8.1f} ms)", file=sys.stderr)
    e2e_total = reduce_sum(e2e_flops)
    # one unpipelined step for comparison (upload -> solve -> download, nothing overlapped)
    barrier()
    t1 = time.perf_counter()
    upload(); pl.solve(tol, maxit); pl.get_matrix("X", "n", L.LAYOUT_RIRIRIRI, out=x_np)
    torch.cuda.synchronize(dev)
    e2e_serial_s = reduce_max(time.perf_counter() - t1)
    if pl2 is not None:
        pl2.close()
    if h2 is not None:
        h2.close()
    lanes = None

    # ---- roofline of the block-sparse product ------------------------------------------------------------------
    nPairs, nnzbX = info["nPairs"], info["nnzbX"]
    spmm_bytes = sp.nnzbA*2*lm*lm*es + 2*nnzbX*2*lm*ln*es + 8*nPairs + 4*(nnzbX + 1)   # SURVEY.md 8d formula
    spmm_flops = nPairs*8*lm*lm*ln
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "spmm_traffic.json"))).get("dram_bytes_per_launch")
    except Exception:
        pass
    spmm_avg_ms = spmm_ms/max(spmm_n, 1)
    achieved = spmm_bytes/(spmm_avg_ms*1e-3)*1e-9 if spmm_n else 0.0
    roofline = {"bound": "hbm", "kernel": "spmm (block-sparse product Y=A*X)", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved/peak, "traffic": traffic if (prec == "c" and (n, lm, ln, ncols) == (32, 32, 32, 2)) else None,
                "peak_source": "measured" if peaks else "fallback",
                "launches_timed": int(spmm_n), "avg_launch_ms": spmm_avg_ms, "algorithmic_bytes_per_launch": spmm_bytes,
                "gflops_per_launch": spmm_flops*1e-9, "achieved_tflops": spmm_flops/(spmm_avg_ms*1e-3)*1e-12 if spmm_n else 0.0,
                "share_of_step": spmm_ms/max(p0.elapsed_time(p1), 1e-9),
                "timed_region": f"{args.steps} solves with CUDA events around every product (kernel-by-kernel launches), "
                                f"{ms_profiled/args.steps:.2f} ms per solve; the headline region runs without them"}

    if prec == "z" and spmm_n:
        # complex fp64 (BASELINE configs 4/5): 64 flop/B and more, the product is bound by the FP64 (DMMA) pipe, not by HBM
        tf = spmm_flops/(spmm_avg_ms*1e-3)*1e-12
        fp64 = measured_fp64_peak()
        roofline.update({"bound": "tensor", "achieved": tf, "peak": fp64["tflops"], "unit": "TFLOP/s", "frac": tf/fp64["tflops"],
                         "peak_source": fp64["source"], "hbm_gbs_algorithmic": achieved})

    # ---- strong scaling: ONE problem of --strong-rhs right-hand-side columns split over the ranks -------------------------------
    strong = None
    if args.strong_rhs > 0 and args.strong_rhs % (ln*world) == 0:
        pl.close(); h.close(); pl = h = None
        ws_t = None; valA_keep = valA_part
        torch.cuda.empty_cache()
        sc_global = args.strong_rhs//ln
        sc = sc_global//world
        sp2 = synthetic.Stencil27(n, lm, ln, sc, sigma=args.sigma, dtype=dt, device=dev, col0=rank*sc, ncols_global=sc_global, with_values=False)
        hs, ps, wss, bases, nbs = _make_plan(torch, api, sp2, lm, ln, prec, dev, shard=(dist, rank, world, sc_global, ex_group) if world > 1 else None)
        parts2 = [ps.matrix_part_info(r, world) for r in range(world)]
        assert parts2[rank]["block0"] == mine["block0"] and parts2[rank]["nblocks"] == mine["nblocks"]
        ps.set_matrix_part(valA_keep.data_ptr(), rank, world)
        if world > 1:
            for r, q in enumerate(parts2):
                if q["length"]:
                    dist.broadcast(wss[bases + q["off"]:bases + q["off"] + q["length"]], src=r)
                if q["scale_length"]:
                    dist.broadcast(wss[bases + q["scale_off"]:bases + q["scale_off"] + q["scale_length"]], src=r)
        ps.set_matrix("B", sp2.valB)
        for _ in range(2):
            ps.solve(tol, maxit)
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sflops = 0
        barrier()
        s0.record()
        for _ in range(args.steps):
            sst = ps.solve(tol, maxit); sflops += ps.info()["flops"]
        s1.record()
        barrier()
        sms = reduce_max(s0.elapsed_time(s1))
        sinfo = ps.info()
        strong = {"rhs_columns_total": args.strong_rhs, "rhs_columns_per_gpu": sc*ln, "ms_per_solve": sms/args.steps,
                  "value": reduce_sum(sflops)/(sms*1e-3)*1e-9, "unit": UNIT, "iterations": sinfo["iterations"], "status": int(reduce_max(float(sst))),
                  "residual_reached": sinfo["residuum"], "workspace_bytes_per_gpu": nbs,
                  "what": "one problem, right-hand-side block columns split over the GPUs, A replicated, global iteration rule"}
        ps.close(); hs.close()

    line = None
    if rank == 0:
        line = {
            "metric": METRIC, "value": total_flops/(ms*1e-3)*1e-9, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms/args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64" if prec == "z" else "f32", "data": "synthetic",
            "config": {"workload": workload_name(n, lm, ln, ncols, prec, tol, args.sigma),
                       "parallelism": f"one {ncols_global*ln}-column problem, rhs block columns sharded x{world}, A replicated, "
                                      f"global iteration rule ({'NCCL all-gather of 3 numbers per rank, twice per iteration' if world > 1 else 'single GPU'})",
                       "l2": "working set 12 GB per GPU >> 126 MB L2, no flush needed",
                       "iterations_per_solve": iters/args.steps, "iterations_per_solve_min_max_over_ranks": [iters_min, iters_max],
                       "residual_reached": last["residuum"], "status": int(statuses[-1]), "worst_status_over_ranks": worst_status,
                       "workspace_bytes": nbytes, "nnzbA": sp.nnzbA, "nnzbX": nnzbX, "nPairs": nPairs,
                       "createPlan_ms": {"first_call_in_process": create_plan_ms, "repeat": create_plan_warm_ms,
                                         "what": "device analysis incl. upload of the index arrays; first call also loads the CUDA modules"},
                       "bufferSize_ms": {"first_call_in_process": configure_ms, "repeat": configure_warm_ms},
                       "product_kernel": "tcgen05 fp16-pair" if info["use_tc"] else ("dmma" if info["use_dmma"] else "simt")},
            "e2e": {"value": e2e_total/e2e_s*1e-9, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "a_distribution": (f"every rank uploads 1/{world} of A over its own PCIe link, converted ranges exchanged with NCCL broadcasts over NVLink"
                                       if world > 1 else "H2D"),
                    "ms_per_step": 1e3*e2e_s/args.steps, "pipelining": "2 plans: upload of step k+1 (chained after upload k) overlaps solve of step k" if pipelined else "none (second workspace not available)",
                    "ms_single_step_unpipelined": 1e3*e2e_serial_s},
            "gpu_launches": int(launches),
            "roofline": roofline,
            "clocks": clocks,
        }
        if strong is not None:
            line["strong"] = strong
    if pl is not None:
        pl.close(); h.close()
    if rank == 0 and world == 1 and not args.no_cpu:
        try:
            r = cpu_reference_sample(lm, ln, ncols, prec, tol, maxit, repeats=1)
            line["cpu_baseline"] = {"value": r["flops"]/r["times"][0]*1e-9, "unit": UNIT, "cores": 1, "kind": r["kind"],
                                    "sample": r["sample"], "host_cores_available": os.cpu_count()}
        except Exception as e:      # the GPU measurement above must still be reported
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 1, "kind": "unavailable", "sample": repr(e)[:200]}
        if not args.no_ref_gpu:
            try:
                spf = synthetic.Stencil27(n, lm, ln, ncols, sigma=args.sigma, dtype=dt, device=dev)
                line["reference_gpu"] = reference_gpu_same_box(spf, lm, ln, prec, tol, maxit)
            except Exception as e:  # informational leg only
                line["reference_gpu"] = {"error": repr(e)[:200]}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def measured_fp64_peak():
    """FP64 rate of this GPU from tools/fp64_rate (DFMA and DMMA.8x8x4 issue-rate microbenchmark, built by the library's Makefile);
    the datasheet value when the tool is missing."""
    exe = os.path.join(ROOT, "tools", "fp64_rate")
    try:
        out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
        vals = dict(l.split("=") for l in out.stdout.split() if "=" in l)
        t = max(float(vals["dfma_tflops"]), float(vals["dmma_tflops"]))
        return {"tflops": t, "source": f"measured (tools/fp64_rate: DFMA {float(vals['dfma_tflops']):.1f}, DMMA {float(vals['dmma_tflops']):.1f} TFLOP/s)"}
    except Exception:
        return {"tflops": 37.0, "source": "nominal B200 fp64 (tools/fp64_rate not available)"}


# ---- the other BASELINE configurations, same JSON schema (run by `--config N`; the headline stays config 3) ---------------------
def _ref_bin(name):
    path = os.path.join(ROOT, "oracle", "_ref", name)
    return path if os.path.exists(path) else None


def run_config1(args):
    """BASELINE configs[0]: the reference's multiplication plan test/multiplication/plan_unordered.14-287-16 (golden copy under
    tests/golden), complex fp32 16x16 blocks, bare product Y = A*X (`bench_tfqmrgpu multi`): a step = one product."""
    import re
    import tempfile
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import orclib as O
    from tfqmrgpu_b200 import api, problems as P, formats as F, _lib as L
    g = np.load(os.path.join(ROOT, "tests", "golden", "plan_unordered.npz"))
    starts, pairs = g["starts"], g["pairs"]
    nY, nA, nX = [int(v) for v in g["nnz"]]
    mb, rpA, ciA, rpX, ciX = P.bsr_from_multiplication_plan(starts, pairs, nA)
    lm = ln = 16
    h = api.Handle(); pl = api.BsrsvPlan(h, mb, rpA, ciA, rpX, ciX, rpX, ciX)
    pl.buffer_size_for(lm, ln, "c"); pl.set_buffer()
    A = O.fill_cos_sin(nA, lm, lm, np.float32); X = O.fill_cos_sin(nX, lm, ln, np.float32)
    pl.set_matrix("A", A, "t", L.LAYOUT_RRRRIIII); pl.set_matrix("X", X, "n", L.LAYOUT_RRRRIIII)
    info = pl.plan_info()
    used_a = len(np.unique(pairs[:, 0]))
    nbytes = used_a*2*lm*lm*4 + 2*nX*2*lm*ln*4 + 8*len(pairs) + 4*(nY + 1)      # 45.14 MB
    flops = len(pairs)*8*lm*lm*ln
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stream_ptr = torch.cuda.current_stream().cuda_stream
    assert stream_ptr == h.get_stream() or h.get_stream() == 0
    pl.multiply(max(args.warmup, 3)); torch.cuda.synchronize()
    sampler = ClockSampler(0); sampler.start(); time.sleep(0.2)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")     # > 126 MB L2: every timed product starts cold
    cold = []
    for _ in range(args.steps):
        flush.zero_(); e0.record(); pl.multiply(1); e1.record(); torch.cuda.synchronize(); cold.append(e0.elapsed_time(e1))
    e0.record(); pl.multiply(200); e1.record(); torch.cuda.synchronize()
    warm_ms = e0.elapsed_time(e1)/200
    clocks = sampler.stop()
    cold_ms = float(np.median(cold))
    Y = pl.get_vector("Y", "n", L.LAYOUT_RRRRIIII).reshape(nX, 2, lm, ln)
    # end to end: X host -> device, product, Y device -> host (A resident like in the harness)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        pl.set_matrix("X", X, "n", L.LAYOUT_RRRRIIII); pl.multiply(1); pl.get_vector("Y", "n", L.LAYOUT_RRRRIIII)
    e2e_ms = 1e3*(time.perf_counter() - t0)/args.steps
    pl.close(); h.close()
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    line = {"metric": "bsr_spmm_time", "value": 1e3*cold_ms, "unit": "us", "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": cold_ms, "higher_is_better": False, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "plan_unordered.14-287-16: 4490 Y blocks, 50526 pairs, complex fp32 16x16 blocks, cos/sin fill (bench_tfqmrgpu multi)",
                       "l2": "working set 45 MB < 126 MB L2: flushed before every timed product (value), warm figure beside it",
                       "warm_us": 1e3*warm_ms, "warm_tflops": flops/warm_ms*1e-9, "units": info["nUnits"], "entries": info["nEntries"],
                       "product_kernel": "tcgen05 fp16-pair" if info["use_tc"] else "simt"},
            "e2e": {"value": 1e3*e2e_ms, "unit": "us", "h2d_bytes_per_step": int(X.nbytes), "d2h_bytes_per_step": int(Y.nbytes)},
            "gpu_launches": 3*args.steps,
            "roofline": {"bound": "hbm", "kernel": "spmm (block-sparse product Y=A*X)", "achieved": nbytes/cold_ms*1e-6, "peak": peak, "unit": "GB/s",
                         "frac": nbytes/cold_ms*1e-6/peak, "traffic": None, "peak_source": "measured" if peaks else "fallback",
                         "avg_launch_ms": cold_ms, "algorithmic_bytes_per_launch": nbytes, "achieved_tflops": flops/cold_ms*1e-9,
                         "note": "148 persistent CTAs share 1663 units of 6-14 entries (2 KB A blocks): the launch is latency-bound "
                                 "(pipeline fill + one epilogue per unit), not HBM-bound; the X operand is converted with the upload of X (setMatrix), the timed launch is the product alone"},
            "clocks": clocks}
    if not args.no_cpu:
        t0 = time.perf_counter()
        Yo = O.multiply(A, X, starts, pairs.reshape(-1), lm, ln, nthreads=os.cpu_count())
        tcpu = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": 1e6*tcpu, "unit": "us", "cores": os.cpu_count(), "kind": "port",
                                "sample": "the whole product once: restatement of the harness's OpenMP check loop (bench_tfqmrgpu.cu:358-404)"}
        line["config"]["maxdev_vs_cpu"] = float(np.abs(Y - Yo).max())
        exe = _ref_bin("bench_tfqmrgpu_ref")
        if exe and not args.no_ref_gpu:      # the reference's own kernel (gemmNxNf, sm_100 build) on this GPU: `multi <plan> f 100 1 16 16`
            with tempfile.TemporaryDirectory() as td:
                plan = os.path.join(td, "plan_unordered.14-287-16")
                F.write_multiplication_plan(plan, starts, pairs, nA, nX)
                out = subprocess.run([exe, "multi", plan, "f", "100", "1", "16", "16"], capture_output=True, text=True, timeout=600)
            m = re.search(r"# GPU performance \(lm,ln,tune\)=\( *16, *16,\d\) is +([0-9.]+) G[fF]lop/sec", out.stdout)
            line["reference_gpu"] = ({"value": flops/float(m.group(1))*1e-3, "unit": "us", "gflops": float(m.group(1)),
                                      "what": "unmodified bench_tfqmrgpu multi (reference kernels, sm_100 build), warm, same GPU"}
                                     if m else {"error": out.stdout[-300:]})
    print(json.dumps(line), flush=True)
    return 0


def run_config2(args):
    """BASELINE configs[1]: full tfQMR solve of FD_problem.xml (generate_FD_example defaults), complex fp64, 1 GPU: a step = one solve."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import orclib as O
    from tfqmrgpu_b200 import api, problems as P
    prob = P.read_xml(os.path.join(ROOT, "tests", "golden", "FD_problem.xml"))
    vA = P.interleave(prob.A.val, np.float64); vB = P.interleave(prob.B.val, np.float64)
    h = api.Handle()
    pl = api.BsrsvPlan(h, prob.mb, prob.A.rowptr, prob.A.colind, prob.X.rowptr, prob.X.colind, prob.B.rowptr, prob.B.colind)
    pl.buffer_size_for(prob.lm, prob.ln, "z"); pl.set_buffer()
    pl.set_matrix("A", vA, "t"); pl.set_matrix("B", vB, "t")
    for _ in range(max(args.warmup, 3)):
        st = pl.solve(prob.tolerance, 2000)
    sampler = ClockSampler(0); sampler.start(); time.sleep(0.2)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(args.steps):
        st = pl.solve(prob.tolerance, 2000)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)/args.steps
    clocks = sampler.stop()
    info = pl.info(); stats = pl.solve_stats()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        pl.set_matrix("A", vA, "t"); pl.set_matrix("B", vB, "t"); pl.solve(prob.tolerance, 2000); X = pl.get_matrix("X")
    e2e_ms = 1e3*(time.perf_counter() - t0)/args.steps
    pl.close(); h.close()
    line = {"metric": "tfqmr_solve_time", "value": ms, "unit": "ms", "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms, "higher_is_better": False, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "FD_problem.xml (generate_FD_example defaults): 171 block rows of 8x8, 1557 A blocks, 1 RHS block column, complex fp64, tol 1e-9",
                       "l2": "working set 3 MB, L2-resident by nature (launch-latency-bound; no flush: a flush would time the flush)",
                       "iterations": info["iterations"], "status": int(st), "residual": info["residuum"], "gflops": info["flops"]/ms*1e-6},
            "e2e": {"value": e2e_ms, "unit": "ms", "h2d_bytes_per_step": int(vA.nbytes + vB.nbytes), "d2h_bytes_per_step": int(X.nbytes)},
            "gpu_launches": int(stats["launches"]),
            "roofline": {"bound": "hbm", "kernel": "whole solve (latency-bound: ~10 MB per iteration)", "achieved": None, "peak": None, "unit": "GB/s",
                         "frac": None, "traffic": None, "note": "the roofline fraction is not meaningful for this 3 MB problem (SURVEY.md 8d): "
                                                                 "time = iterations x (kernel launches of one CUDA graph)"},
            "clocks": clocks}
    if not args.no_cpu:
        for name, ref in (("cpu_baseline", O.ref_cpu()), ("reference_gpu", None if args.no_ref_gpu else O.ref_gpu())):
            if ref is None:
                continue
            tr = []
            for _ in range(3):
                with Quiet():
                    r = ref.solve(prob.mb, prob.lm, prob.ln, prob.A.rowptr, prob.A.colind, vA, prob.X.rowptr, prob.X.colind,
                                  prob.B.rowptr, prob.B.colind, vB, prob.tolerance, 2000, "z", transA="t", trans_b="t")
                tr.append(r["t_solve"])
            entry = {"value": 1e3*float(np.median(tr)), "unit": "ms", "iterations": r["iterations"], "status": int(r["status"])}
            if name == "cpu_baseline":
                entry.update({"cores": 1, "kind": "reference", "sample": "the whole solve (serial reference CPU build), median of 3"})
            else:
                entry["what"] = "unmodified reference CUDA build (sm_100), same GPU, median of 3"
            line[name] = entry
    print(json.dumps(line), flush=True)
    return 0


def run_config5(args):
    """BASELINE configs[4]: sweep of block size x right-hand-side count x precision on the 27-point block stencil; the cases are
    dealt round robin to the ranks (independent problems, no collective), rank 0 prints all rows.  fp64 at tol 1e-9 (sigma 1);
    fp32 cannot reach 1e-9 (its floor is ~1e-4): tol 1e-3 (sigma 8; 1e-2 beyond 64 right-hand sides), stated per row."""
    import torch
    import torch.distributed as dist
    from tfqmrgpu_b200 import api, synthetic
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    sizes = [(4, 4), (4, 5), (4, 8), (4, 32), (8, 8), (8, 9), (8, 10), (8, 32), (8, 64), (16, 16), (16, 32), (16, 64), (32, 32), (32, 64), (64, 64)]
    n = args.sweep_n
    cases = [(lm, ln, rhs, prec) for (lm, ln) in sizes for rhs in args.sweep_rhs for prec in ("c", "z") if rhs >= ln or rhs == min(args.sweep_rhs)]
    rows = []
    t_all0 = time.perf_counter()
    for k, (lm, ln, rhs, prec) in enumerate(cases):
        if k % world != rank:
            continue
        ncols = max(1, rhs//ln)
        dt, sigma, tol, es = (np.float32, 8.0, 1e-3, 4) if prec == "c" else (np.float64, 1.0, 1e-9, 8)
        if prec == "c" and rhs > 64:
            # fp32 with hundreds of right-hand sides: the reference's stopping rule wants ALL of them below the tolerance at the same
            # probe, and the fp32 recurrences of columns that are already done drift (floor 3-4e-3 here; the oracle, i.e. the
            # reference's algorithm, stalls the same way): tol 1e-2, stated per row
            tol = 1e-2
        if 8.5*n**3*ncols*2*lm*ln*es + 27*n**3*2*lm*lm*es > 150e9:
            continue
        sp = synthetic.Stencil27(n, lm, ln, ncols, sigma=sigma, dtype=dt, device=dev)
        h = api.Handle(); pl = api.BsrsvPlan(h, sp.mb, sp.rpA, sp.ciA, sp.rpX, sp.ciX, sp.rpB, sp.ciB)
        pl.buffer_size_for(lm, ln, prec); pl.set_buffer()
        pl.set_matrix("A", None, "n", raw_ptr=sp.valA_host.data_ptr()); pl.set_matrix("B", sp.valB)
        info = pl.plan_info()
        maxit = 200 if prec == "z" else 50
        for _ in range(2):
            st = pl.solve(tol, maxit)
        pl.set_profiling(True); pl.solve(tol, maxit); prof = pl.solve_profile(); pl.set_profiling(False)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); st = pl.solve(tol, maxit); e1.record(); torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1); res = pl.info()
        sp_ms = prof["spmm_ms"]/max(prof["spmm_launches"], 1)
        nP = info["nPairs"]
        rows.append(dict(lm=lm, ln=ln, rhs=ncols*ln, prec=prec, tol=tol, status=int(st), iterations=res["iterations"], residual=res["residuum"],
                         solve_ms=ms, gflops=res["flops"]/ms*1e-6, spmm_us=1e3*sp_ms, spmm_gflops=nP*8*lm*lm*ln/sp_ms*1e-6 if sp_ms else 0,
                         kernel="dmma" if info["use_dmma"] else ("tcgen05" if info["use_tc"] else ("simt-small" if info.get("use_small") else "simt")), rank=rank))
        if int(st) != 0 and prec == "c":
            # the reference's stopping rule stalled (see the note of this record): the same solve with the opt-in early freeze
            pl.set_early_freeze(True)
            pl.solve(tol, maxit)
            f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            f0.record(); fst = pl.solve(tol, maxit); f1.record(); torch.cuda.synchronize(dev)
            fres = pl.info()
            rows[-1]["with_early_freeze"] = dict(status=int(fst), iterations=fres["iterations"], residual=fres["residuum"], solve_ms=f0.elapsed_time(f1))
        pl.close(); h.close(); del sp
        torch.cuda.empty_cache()
    torch.cuda.synchronize(dev)
    wall = time.perf_counter() - t_all0
    if world > 1:
        gathered = [None]*world
        dist.all_gather_object(gathered, rows)
        rows = [r for part in gathered for r in part]
        t = torch.tensor([wall], dtype=torch.float64, device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX); wall = float(t.item())
    if rank == 0:
        rows.sort(key=lambda r: (r["lm"], r["ln"], r["rhs"], r["prec"]))
        total_flops = sum(r["gflops"]*r["solve_ms"]*1e-3 for r in rows)       # GFLOP of one timed solve per case
        busy = sum(r["solve_ms"] for r in rows)*1e-3
        line = {"metric": "tfqmr_sweep_throughput", "value": total_flops/max(busy/world, 1e-9), "unit": "GFLOP/s", "n_gpus": world, "steps": 1, "warmup": 3,
                "ms_per_step": 1e3*busy/world, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32+f64", "data": "synthetic",
                "config": {"workload": f"sweep: 15 block sizes x RHS {list(args.sweep_rhs)} x (fp32 tol 1e-3, 1e-2 beyond 64 RHS | fp64 tol 1e-9) on stencil27 n={n}^3, "
                                       f"{len(rows)} cases dealt round robin to {world} GPU(s)", "wall_s_incl_setup": wall,
                           "all_converged": all(r["status"] == 0 for r in rows),
                           "not_converged": [f"{r['lm']}x{r['ln']} {r['prec']} {r['rhs']} RHS: residual {r['residual']:.1e}" for r in rows if r["status"] != 0],
                           "note": "fp32 rows that end with status 9 (50 iterations) stall in the oracle - the reference's algorithm on the CPU - with the same "
                                   "shadow vector as well (tests/tools/dev_sweep_diag.py): the residual floor of fp32 tfQMR, not a kernel; "
                                   "with_early_freeze = the same solve with tfqmrgpux_bsrsv_setEarlyFreeze (opt-in, not reference behaviour)",
                           "rows": rows},
                "e2e": None, "gpu_launches": None, "roofline": None, "clocks": None}
        if not args.no_cpu and world >= 1:
            try:
                r = cpu_reference_sample(8, 8, 8, "z", 1e-9, 200, repeats=1)
                line["cpu_baseline"] = {"value": r["flops"]/r["times"][0]*1e-9, "unit": UNIT, "cores": 1, "kind": r["kind"],
                                        "sample": r["sample"].replace("32x32", "8x8") + " (one case of the sweep)"}
            except Exception as e:
                line["cpu_baseline"] = {"value": None, "kind": "unavailable", "sample": repr(e)[:200]}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=32, help="grid edge (block rows = n^3); 32 is the BASELINE configuration")
    ap.add_argument("--lm", type=int, default=32)
    ap.add_argument("--ln", type=int, default=32)
    ap.add_argument("--ncols", type=int, default=2, help="block columns of X per GPU")
    ap.add_argument("--precision", default="c", choices=["c", "z"])
    ap.add_argument("--tol", type=float, default=DEFAULT_TOL)
    ap.add_argument("--sigma", type=float, default=8.0, help="diagonal shift of the stencil operator (8: fp32 config; 1: fp64 configs)")
    ap.add_argument("--strong-rhs", type=int, default=512, help="right-hand-side columns of the strong-scaling leg (one problem split "
                    "over the GPUs; 0: skip).  512 columns of the fp32 stencil need 46 GB on one GPU")
    ap.add_argument("--config", type=int, default=3, choices=[1, 2, 3, 4, 5], help="BASELINE.json configuration (1-based); 3 is the headline")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline (and reference_gpu) legs")
    ap.add_argument("--no-ref-gpu", action="store_true", help="skip the informational same-box run of the reference's CUDA kernels")
    ap.add_argument("--sweep-n", type=int, default=12, help="config 5: grid edge of the sweep's stencil")
    ap.add_argument("--sweep-rhs", type=int, nargs="+", default=[64, 512], help="config 5: right-hand-side counts (BASELINE: 8 ... 4096)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.config == 1:
        return run_config1(args)
    if args.config == 2:
        return run_config2(args)
    if args.config == 5:
        return run_config5(args)
    if args.config == 4:
        # BASELINE configs[3]: the stencil in complex fp64 with 1024 right-hand-side columns (32 block columns of 32; read as in
        # SURVEY.md 8d) sharded over the GPUs: strong scaling by construction (the per-GPU share shrinks with N)
        world = int(os.environ.get("WORLD_SIZE", "1"))
        args.precision, args.sigma, args.tol, args.strong_rhs = "z", 1.0, 1e-9, 0
        args.ncols = max(1, 32//world)
        os.environ.setdefault("TFQMRGPU_BENCH_NO_PIPELINE", "1")      # one workspace of 8 fp64 vectors + A per GPU is the budget
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
