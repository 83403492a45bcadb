# dev-only: time the bare product for one block size on the sweep's stencil (N^3 block rows, 64 RHS columns):
#   LM=4 LN=4 PREC=c python scripts/dev_small_time.py        (TFQMRGPU_SMALL=0: ring kernel)
import sys, os, numpy as np, torch
sys.path.insert(0, '.')
from tfqmrgpu_b200 import api, synthetic, _lib as L
n = int(os.environ.get('N', '12')); lm = int(os.environ.get('LM', '4')); ln = int(os.environ.get('LN', '4')); prec = os.environ.get('PREC', 'c')
dt = np.float32 if prec == 'c' else np.float64
sp = synthetic.Stencil27(n, lm, ln, max(1, 64//ln), sigma=8.0, dtype=dt, device='cuda')
h = api.Handle(); pl = api.BsrsvPlan(h, sp.mb, sp.rpA, sp.ciA, sp.rpX, sp.ciX, sp.rpB, sp.ciB)
pl.buffer_size_for(lm, ln, prec); pl.set_buffer()
pl.set_matrix('A', None, 'n', raw_ptr=sp.valA_host.data_ptr()); pl.set_matrix('B', sp.valB)
X = np.random.default_rng(0).uniform(-1, 1, sp.nnzbX*2*lm*ln).astype(dt)
pl.set_matrix('X', X, 'n', L.LAYOUT_RRRRIIII)
pl.multiply(3); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); pl.multiply(20); e1.record(); torch.cuda.synchronize()
print(f'{lm}x{ln} {prec} n={n} gmax={pl.plan_info()["gmax"]} units={pl.plan_info()["nUnits"]}: {1e3*e0.elapsed_time(e1)/20:.1f} us per product', flush=True)
