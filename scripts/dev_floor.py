# dev-only: attainable residual / iterations of config 3 for several tolerances
import sys, os, numpy as np, torch
sys.path.insert(0, '.')
from tfqmrgpu_b200 import api, synthetic, _lib as L
n = int(os.environ.get('N', '32'))
sp = synthetic.Stencil27(n, 32, 32, 2, sigma=8.0, dtype=np.float32, device='cuda')
h = api.Handle(); pl = api.BsrsvPlan(h, sp.mb, sp.rpA, sp.ciA, sp.rpX, sp.ciX, sp.rpB, sp.ciB)
pl.buffer_size_for(32, 32, 'c'); pl.set_buffer()
pl.set_matrix('A', None, 'n', raw_ptr=sp.valA_host.data_ptr()); pl.set_matrix('B', sp.valB)
print('use_tc', pl.plan_info()['use_tc'], 'split', os.environ.get('TFQMRGPU_TC_SPLIT'))
for tol in (1e-3, 3e-4, 1e-4, 5e-5, 2e-5, 1e-5):
    st = pl.solve(tol, 60); i = pl.info(); s = pl.solve_stats()
    print(f'tol {tol:g}: status {st} it {i["iterations"]} res {i["residuum"]:.3e} probes {s["probes"]}')
