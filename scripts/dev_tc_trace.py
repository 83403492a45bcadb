# dev-only: timeline of one unit of CTA 0 of the tcgen05 product (config 3), from a library built with -DTFQ_TC_TRACE:
#   scripts/dev_ablate.sh trace; TFQMRGPU_LIB=$PWD/tfqmrgpu_b200/lib/ablate/libtfQMRgpu_trace.so python scripts/dev_tc_trace.py
import sys, os, ctypes, numpy as np, torch
sys.path.insert(0, '.')
from tfqmrgpu_b200 import api, synthetic, _lib as L
n, lm, ln, ncol = 32, 32, 32, 2
sp = synthetic.Stencil27(n, lm, ln, ncol, sigma=8.0, dtype=np.float32, device='cuda')
h = api.Handle(); pl = api.BsrsvPlan(h, sp.mb, sp.rpA, sp.ciA, sp.rpX, sp.ciX, sp.rpB, sp.ciB)
pl.buffer_size_for(lm, ln, 'c'); pl.set_buffer()
pl.set_matrix('A', None, 'n', raw_ptr=sp.valA_host.data_ptr()); pl.set_matrix('B', sp.valB)
X = np.random.default_rng(0).uniform(-1, 1, sp.nnzbX*2*lm*ln).astype(np.float32)
pl.set_matrix('X', X, 'n', L.LAYOUT_RRRRIIII)
pl.multiply(3); torch.cuda.synchronize()
lib = ctypes.CDLL(os.environ['TFQMRGPU_LIB'])
buf = np.zeros(64*24, dtype=np.int64)
assert 0 == lib.tfq_tc_trace_dump(buf.ctypes.data_as(ctypes.c_void_p), buf.size)
t = buf.reshape(64, 24)
ne = int((t[:, 0] != 0).sum()); t = t[:ne]; t0 = t[t != 0].min()
names = ['c0.start', 'c0.ldx', 'c0.stage', 'c0.split', 'c0.A', 'c0.lo', 'c0.ready', 'c7.start', 'c7.ldx', 'c7.stage', 'c7.split', 'c7.A', 'c7.lo', 'c7.ready',
         '-', 'm.top', 'm.ready', 'm.issued']      # MMA warp: loop top, stage barrier passed, 8 MMAs + commits issued (the copy warp is not traced)
print('entry ' + ' '.join(f'{s:>9s}' for s in names))
for e in range(ne):
    print(f'{e:5d} ' + ' '.join(f'{int(t[e, k] - t0) if t[e, k] else -1:9d}' for k in range(18)))
d = np.diff(t[:, 6]); print('ready-to-ready per entry:', d.tolist())
