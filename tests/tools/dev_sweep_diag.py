# Diagnostic (one GPU): fp32 sweep cases with hundreds of right-hand sides that stall on the GPU - product and solve against the oracle.
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import orclib as O
from tfqmrgpu_b200 import api, synthetic, _lib as L

n = int(sys.argv[1]) if len(sys.argv) > 1 else 6
cases = ((4, 5, 510), (8, 64, 512), (4, 4, 512), (8, 32, 512)) if n < 10 else ((4, 5, 510),)
maxit = 30 if n < 10 else 14
for (lm, ln, rhs) in cases:
    ncols = rhs//ln
    sp = synthetic.Stencil27(n, lm, ln, ncols, sigma=8.0, dtype=np.float32, device="cuda")
    A = sp.valA_host.numpy()
    op = O.OraclePlan(sp.mb, sp.rpA, sp.ciA, sp.rpX, sp.ciX, sp.rpB, sp.ciB)
    Ai = O.import_blocks(A, sp.nnzbA, lm, lm, var="A"); Bi = O.import_blocks(sp.valB, sp.nnzbB, lm, ln, var="B")
    for small in (("1", "0") if n < 10 else ("1",)):
        os.environ["TFQMRGPU_SMALL"] = small
        h = api.Handle(); pl = api.BsrsvPlan(h, sp.mb, sp.rpA, sp.ciA, sp.rpX, sp.ciX, sp.rpB, sp.ciB)
        pl.buffer_size_for(lm, ln, "c"); pl.set_buffer()
        info = pl.plan_info()
        pl.set_matrix("A", A); pl.set_matrix("B", sp.valB)
        # product on a random X
        rng = np.random.default_rng(3)
        X = rng.standard_normal((pl.nnzbX, 2, lm, ln)).astype(np.float32)
        pl.set_matrix("X", X, "n", L.LAYOUT_RRRRIIII)
        pl.multiply(1)
        Y = pl.get_vector("Y", "n", L.LAYOUT_RRRRIIII).reshape(pl.nnzbX, 2, lm, ln)
        Yo = O.multiply(Ai, X, op.starts, op.pairs.reshape(-1), lm, ln, nthreads=8)
        perr = float(np.abs(Y - Yo).max()); pscale = float(np.abs(Yo).max())
        st = pl.solve(1e-2, maxit)
        gi = pl.info(); v3 = pl.get_v3().copy(); rs = pl.rhs_status()
        o = O.solve(op, lm, ln, Ai, Bi, v3, 1e-2, maxit)
        print(f"{lm}x{ln} rhs {rhs} small={small} (use_small {info['use_small']}, gmax {info['gmax']}, units {info['nUnits']}, tiles {info['nTiles']}): "
              f"product err {perr:.2e} of {pscale:.2e}; GPU status {st} it {gi['iterations']} res {gi['residuum']:.2e} bad {int((rs<0).sum())} | "
              f"oracle status {o['status']} it {o['iterations']} res {o['residuum']:.2e} bad {int((o['rhs_status']<0).sum())}", flush=True)
        pl.close(); h.close()
