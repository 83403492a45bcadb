# Diagnostic (torchrun, one rank per GPU): where does the time of a column-sharded solve go?  Per rank, config 3 shard:
#   indep  - no exchange (every shard decides for itself, iteration bodies as CUDA graphs)
#   null   - exchange registered, hook does nothing (kernel-by-kernel path, local decisions)
#   nccl   - the real exchange (all-gather of 4 doubles per rank, twice per iteration)
# plus the raw latency of that all-gather.  Rank 0 writes one line per mode with every rank's ms per solve.
import os, sys, time, json
import numpy as np
import torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tfqmrgpu_b200 import api, synthetic, sharded

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n, lm, ln, ncols, prec, tol, maxit = 32, 32, 32, 2, "c", 1e-3, 100
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
ncols_global = world*ncols
sp = synthetic.Stencil27(n, lm, ln, ncols, sigma=8.0, dtype=np.float32, device=dev, col0=rank*ncols, ncols_global=ncols_global, with_values=False)
out = []


def gather(x):
    t = torch.tensor([x], dtype=torch.float64, device=dev)
    g = torch.zeros(world, dtype=torch.float64, device=dev)
    dist.all_gather_into_tensor(g, t)
    return [round(v, 3) for v in g.tolist()]


class NullExchange:
    def __init__(self, plan):
        self.slots = torch.zeros(2*2*world*4, dtype=torch.float64, device=dev)
        self.calls = 0
        def hook(ptr, count, stream):
            self.calls += 1
            return 0
        plan.set_shard_exchange(rank, world, ncols_global*ln, self.slots.data_ptr(), hook)


class TimedNccl(sharded.NcclExchange):
    pass


valA_part = None
for mode in ("indep", "null", "nccl"):
    h = api.Handle(torch.cuda.current_stream(dev).cuda_stream)
    pl = api.BsrsvPlan(h, sp.mb, sp.rpA, sp.ciA, sp.rpX, sp.ciX, sp.rpB, sp.ciB)
    keep = None
    if mode != "indep":
        pl.set_shard_hints(0, ncols_global)
        keep = NullExchange(pl) if mode == "null" else sharded.NcclExchange(pl, dist, rank, world, ncols_global*ln, dev)
    nbytes = pl.buffer_size_for(lm, ln, prec)
    ws_t = torch.empty(nbytes + 256, dtype=torch.uint8, device=dev)
    ws_ptr = (ws_t.data_ptr() + 255) & ~255
    pl.set_buffer(ws_ptr, keep_alive=(ws_t, keep))
    base = ws_ptr - ws_t.data_ptr()
    parts = [pl.matrix_part_info(r, world) for r in range(world)]
    mine = parts[rank]
    if valA_part is None:
        valA_part = sp.values_of(mine["block0"], mine["block0"] + mine["nblocks"])
    pl.set_matrix_part(valA_part.data_ptr(), rank, world)
    for r, q in enumerate(parts):
        if q["length"]:
            dist.broadcast(ws_t[base + q["off"]:base + q["off"] + q["length"]], src=r)
        if q["scale_length"]:
            dist.broadcast(ws_t[base + q["scale_off"]:base + q["scale_off"] + q["scale_length"]], src=r)
    pl.set_matrix("B", sp.valB)
    for _ in range(3):
        pl.solve(tol, maxit)
    torch.cuda.synchronize(dev); dist.barrier(); torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    its = 0
    for _ in range(steps):
        pl.solve(tol, maxit); its += pl.info()["iterations"]
    e1.record()
    torch.cuda.synchronize(dev)
    host_ms = 1e3*(time.perf_counter() - t0)/steps
    hook_ms = 1e3*getattr(keep, "cpu_s", 0.0)/max(getattr(keep, "calls", 0), 1)
    ms = e0.elapsed_time(e1)/steps
    # per-rank product time
    pl.set_profiling(True)
    pl.solve(tol, maxit)
    prof = pl.solve_profile()
    pl.set_profiling(False)
    rec = dict(mode=mode, ms_per_solve=gather(ms), host_ms=gather(host_ms), iterations=gather(its/steps),
               hook_cpu_ms_per_call=gather(hook_ms), spmm_avg_ms=gather(prof["spmm_ms"]/max(prof["spmm_launches"], 1)), profiled_solve_ms=gather(prof.get("solve_ms", 0.0)))
    out.append(rec)
    pl.close(); h.close(); del ws_t, keep
    torch.cuda.empty_cache()

# raw all-gather latency on 4 doubles per rank
slots = torch.zeros(world*4, dtype=torch.float64, device=dev); minev = torch.ones(4, dtype=torch.float64, device=dev)
for _ in range(20):
    dist.all_gather_into_tensor(slots, minev)
torch.cuda.synchronize(dev); dist.barrier(); torch.cuda.synchronize(dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
e0.record()
for _ in range(200):
    dist.all_gather_into_tensor(slots, minev)
e1.record()
cpu_us = 1e6*(time.perf_counter() - t0)/200
torch.cuda.synchronize(dev)
out.append(dict(mode="raw all_gather", gpu_us_per_call=gather(1e3*e0.elapsed_time(e1)/200), cpu_enqueue_us_per_call=gather(cpu_us)))
dist.barrier()
if 0 == rank:
    with open(sys.argv[1], "w") as f:
        for rec in out:
            f.write(json.dumps(rec) + "\n")
dist.destroy_process_group()
