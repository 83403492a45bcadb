# Diagnostic: host->device copy rate from pinned memory alone and while a config-3 solve loop keeps the GPU's HBM busy.
import os, sys, time, threading
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tfqmrgpu_b200 import api, synthetic

dev = torch.device("cuda", 0); torch.cuda.set_device(0)
sp = synthetic.Stencil27(32, 32, 32, 2, sigma=8.0, dtype=np.float32, device=dev)
h = api.Handle(); pl = api.BsrsvPlan(h, sp.mb, sp.rpA, sp.ciA, sp.rpX, sp.ciX, sp.rpB, sp.ciB)
pl.buffer_size_for(32, 32, "c"); pl.set_buffer()
pl.set_matrix("A", None, "n", raw_ptr=sp.valA_host.data_ptr()); pl.set_matrix("B", sp.valB)
pl.solve(1e-3, 100)
src = torch.empty(1 << 30, dtype=torch.uint8).pin_memory(); dst = torch.empty(1 << 30, dtype=torch.uint8, device=dev)
side = torch.cuda.Stream(dev)

def copy_rate(n=4):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(side):
        e0.record()
        for _ in range(n):
            dst.copy_(src, non_blocking=True)
        e1.record()
    e1.synchronize()
    return n*src.numel()/(e0.elapsed_time(e1)*1e-3)*1e-9

print("H2D alone: %.1f GB/s" % copy_rate(), flush=True)
stop = False
def solver():
    while not stop:
        pl.solve(1e-3, 100)
t = threading.Thread(target=solver); t.start()
time.sleep(0.2)
print("H2D during solves: %.1f GB/s" % copy_rate(), flush=True)
stop = True; t.join()
print("H2D alone again: %.1f GB/s" % copy_rate(), flush=True)
pl.close(); h.close()
