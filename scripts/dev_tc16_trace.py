# dev-only: per-role cycle accounting of CTA 0 of the fp16-pair tcgen05 product (config 3), from a library built with -DTFQ_TC16_TRACE:
#   scripts/dev_ablate16.sh trace; TFQMRGPU_LIB=$PWD/tfqmrgpu_b200/lib/ablate/libtfQMRgpu16_trace.so python scripts/dev_tc16_trace.py
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
buf = np.zeros(64, dtype=np.int64)
assert 0 == lib.tfq_tc16_trace_dump(buf.ctypes.data_as(ctypes.c_void_p), buf.size)
t = buf.reshape(8, 8)
info = pl.plan_info()
ne = info['nEntries']/148.
roles = [('copy warp', ['wait done', 'issue copy']), ('MMA warp 0', ['wait a/x full', 'issue', 'wait acc free', 'segment setup']),
         ('MMA warp 1', ['wait a/x full', 'issue', 'wait acc free', 'segment setup']),
         ('converters hi', ['wait X landed', 'LDS + transform', 'wait stage free', 'st + arrive']), ('converters lo', ['wait X landed', 'LDS + transform', 'wait stage free', 'st + arrive']),
         ('epilogue warp', ['wait acc full', 'tmem ld + sum', 'store'])]
for r, (name, laps) in enumerate(roles):
    tot = t[r].sum()
    print(f'{name:14s} total {tot:9d} cycles = {tot/ne:7.1f} per entry of the CTA | ' + ', '.join(f'{l} {t[r, k]/ne:6.1f}' for k, l in enumerate(laps) if l != '-'))
