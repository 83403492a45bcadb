# dev-only: time the bare block-sparse product of config 3 under experiment switches (TFQMRGPU_TC_DBG read per launch)
import sys, os, numpy as np, torch
sys.path.insert(0, '.')
from tfqmrgpu_b200 import api, synthetic, _lib as L
n = int(os.environ.get('N', '32')); lm = 32; ln = int(os.environ.get('LN', '32')); ncol = int(os.environ.get('NCOL', '2'))
sp = synthetic.Stencil27(n, lm, ln, ncol, sigma=8.0, dtype=np.float32, device='cuda')
h = api.Handle(); pl = api.BsrsvPlan(h, sp.mb, sp.rpA, sp.ciA, sp.rpX, sp.ciX, sp.rpB, sp.ciB)
pl.buffer_size_for(lm, ln, 'c'); pl.set_buffer()
pl.set_matrix('A', None, 'n', raw_ptr=sp.valA_host.data_ptr()); pl.set_matrix('B', sp.valB)
X = np.random.default_rng(0).uniform(-1, 1, sp.nnzbX*2*lm*ln).astype(np.float32)
pl.set_matrix('X', X, 'n', L.LAYOUT_RRRRIIII)
for dbg in [int(v) for v in sys.argv[1:]] or [0]:
    os.environ['TFQMRGPU_TC_DBG'] = str(dbg)
    pl.multiply(3); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); pl.multiply(10); e1.record(); torch.cuda.synchronize()
    print(f'dbg {dbg:3d}: {e0.elapsed_time(e1)/10:.3f} ms per product', flush=True)
