# dev-only: attainable residual / iterations for several tolerances (env: N, LM, LN, NCOL, SIGMA)
import sys, os, numpy as np, torch
sys.path.insert(0, '.')
from tfqmrgpu_b200 import api, synthetic, _lib as L
n = int(os.environ.get('N', '32')); lm = int(os.environ.get('LM', '32')); ln = int(os.environ.get('LN', '32'))
ncol = int(os.environ.get('NCOL', str(max(1, 64//ln)))); sigma = float(os.environ.get('SIGMA', '8'))
sp = synthetic.Stencil27(n, lm, ln, ncol, sigma=sigma, dtype=np.float32, device='cuda')
h = api.Handle(); pl = api.BsrsvPlan(h, sp.mb, sp.rpA, sp.ciA, sp.rpX, sp.ciX, sp.rpB, sp.ciB)
pl.buffer_size_for(lm, ln, 'c'); pl.set_buffer()
pl.set_matrix('A', None, 'n', raw_ptr=sp.valA_host.data_ptr()); pl.set_matrix('B', sp.valB)
print('n', n, 'block', lm, ln, 'use_tc', pl.plan_info()['use_tc'])
for tol in (1e-2, 3e-3, 1e-3, 3e-4, 1e-4, 1e-5):
    st = pl.solve(tol, 60); i = pl.info(); s = pl.solve_stats()
    print(f'  tol {tol:g}: status {st} it {i["iterations"]} res {i["residuum"]:.3e} probes {s["probes"]}')
