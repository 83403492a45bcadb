"""Device-side generation of the synthetic benchmark systems (BASELINE.json configs 3-5).

The 27-point block-stencil operator of config 3 has 7.25 GB of A values; hashing them with numpy takes
minutes, so the SAME counter-based hash as ``problems._hash_uniform`` is evaluated with torch on the GPU
(int64 wrap-around arithmetic, bit-identical values, checked in tests/test_gpu_parity.py) and the result
is parked in pinned host memory, from where the C-ABI's ``setMatrix`` uploads it like any caller's data.
Torch is used for memory and elementwise generation only; nothing here is on the solver's hot path.
"""
from __future__ import annotations

import numpy as np

from . import problems as P

_M64 = 0xFFFFFFFFFFFFFFFF


def _s64(v: int) -> int:
    v &= _M64
    return v - (1 << 64) if v >= (1 << 63) else v


def hash_uniform_torch(idx, seed: int):
    """torch int64 tensor of counters -> float64 uniform(-1,1), bit-identical to problems._hash_uniform."""
    import torch

    def lsr(x, k):
        return (x >> k) & ((1 << (64 - k)) - 1)
    x = idx + _s64(int(seed)*0x9E3779B97F4A7C15)
    x = x ^ lsr(x, 30); x = x*_s64(0xBF58476D1CE4E5B9)
    x = x ^ lsr(x, 27); x = x*_s64(0x94D049BB133111EB)
    x = x ^ lsr(x, 31)
    return lsr(x, 11).to(torch.float64)*(2.0/9007199254740992.0) - 1.0


class Stencil27:
    """Periodic n^3 grid, 27 A blocks per block row, X dense in ``ncols`` block columns, one unit B block per
    block column.  ``valA_host`` is a pinned torch tensor
    [nnzbA, lm, lm, 2] in the caller's RIRIRIRI layout; all index arrays are numpy int32."""

    def __init__(self, n, lm, ln, ncols, sigma=8.0, seed=1234, dtype=np.float32, device="cuda", pin=True,
                 chunk_blocks=1 << 15, col0=0, ncols_global=None, with_values=True):
        import torch
        self.n, self.lm, self.ln, self.ncols, self.sigma, self.seed = n, lm, ln, ncols, sigma, seed
        self.mb = n**3
        self.rpA, self.ciA = P.stencil27_pattern(n)
        self.nnzbA = int(self.ciA.size)
        self.rpX = (ncols*np.arange(self.mb + 1)).astype(np.int32)
        # a shard of a wider problem: global block columns [col0, col0 + ncols) of ncols_global
        ncols_global = ncols_global or ncols
        cols = col0 + np.arange(ncols, dtype=np.int32)
        self.ciX = np.tile(cols, self.mb)
        self.nnzbX = int(self.ciX.size)
        brow = (cols.astype(np.int64)*(self.mb//ncols_global))
        counts = np.zeros(self.mb, np.int64); counts[brow] += 1
        self.rpB = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
        self.ciB = cols[np.argsort(brow, kind="stable")]
        self.nnzbB = ncols
        valB = np.zeros((ncols, lm, ln, 2), dtype=dtype)
        for j in range(ln):
            valB[:, j % lm, j, 0] = 1
        self.valB = valB
        tdt = torch.float64 if dtype == np.float64 else torch.float32
        self._a_bytes = self.nnzbA*lm*lm*2*(8 if dtype == np.float64 else 4)
        self.valA_host = None
        self._gen = (tdt, pin, torch.device(device), chunk_blocks)
        if not with_values:      # a rank that uploads only a range of the replicated operator calls values_of() itself
            return
        self.valA_host = self.values_of(0, self.nnzbA)

    def values_of(self, block0, block1):
        """Pinned host tensor [block1 - block0, lm, lm, 2] with the A blocks [block0, block1) (the same values whatever the range)."""
        import torch
        tdt, pin, dev, chunk_blocks = self._gen
        lm, sigma, seed = self.lm, self.sigma, self.seed
        out = torch.empty((block1 - block0, lm, lm, 2), dtype=tdt, pin_memory=pin)
        per = lm*lm*2
        rp_d = torch.from_numpy(self.rpA.astype(np.int64)).to(dev)
        ci_d = torch.from_numpy(self.ciA.astype(np.int64)).to(dev)
        eye = torch.eye(lm, dtype=torch.float64, device=dev)
        ar = torch.arange(per, dtype=torch.int64, device=dev)
        for b0 in range(block0, block1, chunk_blocks):
            b1 = min(block1, b0 + chunk_blocks)
            blocks = torch.arange(b0, b1, dtype=torch.int64, device=dev)
            u = hash_uniform_torch(blocks[:, None]*per + ar[None, :], seed).view(b1 - b0, lm, lm, 2)*0.05
            rows = torch.searchsorted(rp_d, blocks, right=True) - 1
            shift = torch.where(rows == ci_d[b0:b1], 27.0 + sigma, -1.0).to(torch.float64)
            u[..., 0] += shift[:, None, None]*eye[None]
            out[b0 - block0:b1 - block0].copy_(u.to(tdt))
        if dev.type == "cuda":
            torch.cuda.synchronize(dev)
        return out

    @property
    def a_bytes(self):
        return self._a_bytes
