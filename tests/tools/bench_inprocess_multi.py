# One process, N devices behind the unchanged C-ABI (tfqmrgpux_bsrsv_setDevices): config 3 with 64 right-hand sides per GPU as ONE problem.
# Prints one JSON line: upload (setMatrix A + B), solve, download (getMatrix X) in ms, iterations, aggregate GFLOP/s.
import os, sys, time, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from tfqmrgpu_b200 import api, synthetic, _lib as L

ndev = int(sys.argv[1]) if len(sys.argv) > 1 else torch.cuda.device_count()
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
n, lm, ln, tol = 32, 32, 32, 1e-3
ncols = 2*ndev
torch.cuda.set_device(0)
sp = synthetic.Stencil27(n, lm, ln, ncols, sigma=8.0, dtype=np.float32, device="cuda:0")
h = api.Handle()
pl = api.BsrsvPlan(h, sp.mb, sp.rpA, sp.ciA, sp.rpX, sp.ciX, sp.rpB, sp.ciB)
pl.set_devices(ndev)
nbytes = pl.buffer_size_for(lm, ln, "c"); pl.set_buffer()
x_host = torch.empty(sp.nnzbX*lm*ln*2, dtype=torch.float32).pin_memory()
valB = torch.from_numpy(sp.valB).pin_memory()
rec = dict(devices=ndev, rhs=ncols*ln, workspace_bytes_home=nbytes, upload_ms=[], solve_ms=[], download_ms=[])
for k in range(steps + 1):
    t0 = time.perf_counter()
    pl.set_matrix("A", None, "n", raw_ptr=sp.valA_host.data_ptr()); pl.set_matrix("B", None, "n", raw_ptr=valB.data_ptr())
    for d in range(ndev):
        torch.cuda.synchronize(d)
    t1 = time.perf_counter()
    st = pl.solve(tol, 100)
    t2 = time.perf_counter()
    pl.get_matrix("X", "n", L.LAYOUT_RIRIRIRI, out=x_host.numpy())
    t3 = time.perf_counter()
    if k > 0:      # the first round warms up (module load, first touch)
        rec["upload_ms"].append(1e3*(t1 - t0)); rec["solve_ms"].append(1e3*(t2 - t1)); rec["download_ms"].append(1e3*(t3 - t2))
info = pl.info()
rec.update(status=int(st), iterations=info["iterations"], residual=info["residuum"], flops=info["flops"],
           gflops=info["flops"]/(np.median(rec["solve_ms"])*1e-3)*1e-9,
           what="one process, one host thread, N devices behind tfqmrgpu_bsrsv_* (tfqmrgpux_bsrsv_setDevices); host wall-clock times")
print(json.dumps(rec), flush=True)
pl.close(); h.close()
