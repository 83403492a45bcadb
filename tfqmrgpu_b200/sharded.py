"""Multi-GPU tfQMR by sharding the independent right-hand-side block columns.

Every tfQMR scalar is per right-hand-side column and ``A*X`` never mixes block columns, so the block
columns of X/B are partitioned into contiguous ranges (balanced by X blocks per column), A is
replicated, and every rank (one process per GPU) solves its own sub-problem through the ordinary
C-ABI with NO data-path collective.  NCCL (via ``torch.distributed``) is only used when the caller asks
for the gathered X.  The shadow vector v3 of a shard is the matching slice of the 1-GPU cuRAND stream,
so a shard reproduces the 1-GPU numbers of its columns.

Convergence control: the reference's iteration counter and probe schedule are GLOBAL - one maximum over
all right-hand sides decides (core.hxx:239-299).  With a process group (``dist``) every rank registers an
exchange (``tfqmrgpux_bsrsv_setShardExchange``): after K4 and after N3 the ranks all-gather three numbers
per shard on the solver's stream (NCCL) and take the same decision, so an N-rank run has the single-GPU
iteration count and - vector tiles are cut per block column, whatever the shard - the single-GPU bits.
Without a process group (ranks solved one after the other, e.g. on one GPU) every shard runs the rule on
its own columns and may stop earlier than the global maximum would.

The pure index logic (``partition_columns``, ``shard_pattern``, ``global_block_index``) is numpy-only and
is what the world_size-2 gloo tests exercise on CPU.
"""
from __future__ import annotations

import numpy as np


def dense_column_ids(ciX: np.ndarray) -> tuple[np.ndarray, int]:
    """Dense renumbering of the used block-column values (tfqmrgpu.cu:254-314): colindx, nCols."""
    uniq, inv = np.unique(np.asarray(ciX), return_inverse=True)
    return inv.astype(np.int64), int(uniq.size)


def partition_columns(colindx: np.ndarray, ncols: int, world: int) -> list[tuple[int, int]]:
    """Contiguous ranges [c0, c1) of dense block-column ids, one per rank, balanced by the number of X
    blocks; every rank gets at least one column while columns last (ranks beyond ncols get (n, n))."""
    counts = np.bincount(colindx, minlength=ncols).astype(np.float64)
    cum = np.concatenate([[0.], np.cumsum(counts)])
    bounds = [0]
    for r in range(1, world):
        target = cum[-1]*r/world
        c = int(np.searchsorted(cum, target, side="left"))
        c = min(max(c, bounds[-1] + 1), ncols - (world - r))  # keep >= 1 column for the remaining ranks
        c = max(c, bounds[-1])                                 # (unless columns ran out)
        bounds.append(min(c, ncols))
    bounds.append(ncols)
    return [(bounds[r], bounds[r + 1]) for r in range(world)]


def shard_pattern(rp: np.ndarray, ci_dense: np.ndarray, c0: int, c1: int, index_offset: int = 0):
    """Restrict a BSR pattern (dense column ids) to columns [c0, c1).
    Returns (rowptr' (zero-based), kept block indices in the parent's order)."""
    rp = np.asarray(rp, np.int64) - index_offset
    keep = (ci_dense >= c0) & (ci_dense < c1)
    sel = np.flatnonzero(keep)
    rows = np.repeat(np.arange(rp.size - 1), np.diff(rp))
    newrp = np.concatenate([[0], np.cumsum(np.bincount(rows[sel], minlength=rp.size - 1))]).astype(np.int32)
    return newrp, sel


class ShardSpec:
    """Index bookkeeping of one rank's sub-problem."""

    def __init__(self, rpX, ciX, rpB, ciB, rank: int, world: int, index_offset: int = 0):
        ciX = np.asarray(ciX); ciB = np.asarray(ciB)
        colindx, ncols = dense_column_ids(ciX)
        self.ncols_global = ncols
        self.ranges = partition_columns(colindx, ncols, world)
        self.c0, self.c1 = self.ranges[rank]
        uniq = np.unique(ciX)
        lut = {int(v): i for i, v in enumerate(uniq)}
        colB = np.array([lut.get(int(v), -1) for v in ciB], np.int64) if ciB.size else np.zeros(0, np.int64)
        self.rpX, self.selX = shard_pattern(rpX, colindx, self.c0, self.c1, index_offset)
        self.rpB, self.selB = shard_pattern(rpB, colB, self.c0, self.c1, index_offset)
        self.ciX = ciX[self.selX].astype(np.int32)
        self.ciB = ciB[self.selB].astype(np.int32)
        self.nnzbX_global = int(ciX.size)
        self.empty = (self.c1 <= self.c0)


def scatter_shards(shards: list[np.ndarray], sels: list[np.ndarray], nnzbX_global: int) -> np.ndarray:
    """Assemble the global X (caller block order) from per-rank X arrays in THEIR caller order."""
    first = next(s for s in shards if s is not None and s.size)
    out = np.zeros((nnzbX_global,) + first.shape[1:], first.dtype)
    for x, sel in zip(shards, sels):
        if x is not None and sel.size:
            out[sel] = x
    return out


class NcclExchange:
    """The per-iteration exchange of a one-process-per-GPU run (``tfqmrgpux_bsrsv_setShardExchange``): whenever the solver needs
    the other shards' convergence monitors it calls back, and the ranks all-gather ``world`` blocks of 4 doubles on the solver's
    stream with NCCL (``torch.distributed``).  Keep the object alive as long as the plan solves."""

    def __init__(self, plan, dist, rank, world, n_rhs_global, device, group=None):
        """group: the process group (communicator) of the exchange.  Give it its own one (dist.new_group()) when other collectives of
        the caller - e.g. the distribution of the next system's operator - may be in flight on other streams: operations on ONE NCCL
        communicator execute in issue order, and the solve would wait behind them."""
        import torch
        self.slots = torch.zeros(2*2*world*4, dtype=torch.float64, device=device)
        self.mine = torch.zeros(4, dtype=torch.float64, device=device)
        self.calls = 0
        self.cpu_s = 0.0                  # host time spent in the hook (diagnostics)
        base = self.slots.data_ptr()

        def hook(ptr, count, stream):
            import time
            t0 = time.perf_counter()
            off = (ptr - base)//8
            view = self.slots[off:off + count]
            ext = torch.cuda.ExternalStream(stream, device=device) if stream else torch.cuda.default_stream(device)
            with torch.cuda.stream(ext):
                self.mine.copy_(view[4*rank:4*rank + 4])
                dist.all_gather_into_tensor(view, self.mine, group=group)
            self.calls += 1
            self.cpu_s += time.perf_counter() - t0
            return 0
        plan.set_shard_exchange(rank, world, n_rhs_global, base, hook)


class ShardedBsrsv:
    """One rank of a block-column-sharded solve (needs a GPU; uses torch only for device memory and the
    optional NCCL gather)."""

    def __init__(self, mb, lm, ln, precision, rpA, ciA, valA, transA, rpX, ciX, rpB, ciB, valB, transB,
                 rank=0, world=1, index_offset=0, device=None, global_v3=True, dist=None):
        import torch
        from . import api
        self.torch = torch
        self.rank, self.world = rank, world
        self.lm, self.ln, self.precision = lm, ln, precision
        self.spec = ShardSpec(rpX, ciX, rpB, ciB, rank, world, index_offset)
        self.device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.handle = api.Handle(torch.cuda.current_stream(self.device).cuda_stream)
        self.plan = None
        self._col_global, _ = dense_column_ids(np.asarray(ciX))
        if self.spec.empty:
            return
        s = self.spec
        rpA0 = np.asarray(rpA, np.int32) - index_offset
        ciA0 = np.asarray(ciA, np.int32) - index_offset
        self.plan = api.BsrsvPlan(self.handle, mb, rpA0, ciA0, s.rpX, s.ciX, s.rpB, s.ciB, 0, 0)
        if world > 1:
            # choose the product kernel like the unsharded plan would (vector tiles are cut per block column: they agree anyway)
            max_cols = int(np.diff(np.asarray(rpX, np.int64)).max())
            self.plan.set_shard_hints(0, max_cols)
        if dist is not None and world > 1:
            self._register_exchange(dist, rank, world, s.ncols_global*ln)
        nbytes = self.plan.buffer_size_for(lm, ln, precision)
        self.workspace = torch.empty(nbytes + 256, dtype=torch.uint8, device=self.device)
        base = self.workspace.data_ptr()
        self.ws_ptr = (base + 255) & ~255
        self.plan.set_buffer(self.ws_ptr, keep_alive=self.workspace)
        if global_v3 and world > 1:
            # slice of the 1-GPU shadow vector: regenerate the global stream, keep this rank's blocks
            import ctypes as C
            n = s.nnzbX_global*2*lm*ln
            full = torch.empty(n, dtype=torch.float32, device=self.device)
            st = self.plan.lib.tfqmrgpux_randomShadow(self.handle.h, C.c_void_p(full.data_ptr()), C.c_size_t(n))
            assert st == 0, st
            sel = torch.from_numpy(s.selX).to(self.device)
            mine = full.view(s.nnzbX_global, 2*lm*ln).index_select(0, sel).contiguous()
            torch.cuda.current_stream(self.device).synchronize()
            self.plan.set_v3(mine.data_ptr(), on_device=True)
            del full, mine
        valB = np.asarray(valB)
        self.plan.set_matrix("A", valA, transA)
        self.plan.set_matrix("B", valB[s.selB], transB)

    def _register_exchange(self, dist, rank, world, n_rhs_global):
        """All-gather of the shards' convergence monitors on the solver's stream (the library calls this twice per iteration)."""
        self._exchange = NcclExchange(self.plan, dist, rank, world, n_rhs_global, self.device)

    def solve(self, threshold, max_iterations):
        if self.plan is None:
            return 0
        return self.plan.solve(threshold, max_iterations)

    def info(self):
        if self.plan is None:
            return dict(residuum=0., iterations=0, flops=0., flops_all=0.)
        return self.plan.info()

    def x_device(self):
        """This rank's X window as a torch tensor [nnzbX_local, 2, lm, ln] in STORAGE (column-sorted) order."""
        torch = self.torch
        off, length = self.plan.window("X")
        dt = torch.float64 if self.precision == "z" else torch.float32
        start = (self.ws_ptr - self.workspace.data_ptr()) + off
        return self.workspace[start:start + length].view(dt).view(-1, 2, self.lm, self.ln)

    def gather_x(self, dist=None):
        """All ranks receive the global X [nnzbX_global, 2, lm, ln] (internal RRRRIIII block layout) in the
        caller's block order.  One NCCL all_gather of the padded X windows; no host round trip."""
        torch = self.torch
        s = self.spec
        dt = torch.float64 if self.precision == "z" else torch.float32
        if self.plan is not None:
            perm = torch.from_numpy(self.plan.plan_array(4).astype(np.int64)).to(self.device)  # caller -> storage
            mine = self.x_device().index_select(0, perm)                                        # caller order of the shard
        else:
            mine = torch.zeros((0, 2, self.lm, self.ln), dtype=dt, device=self.device)
        if dist is None or self.world == 1:
            out = torch.zeros((s.nnzbX_global, 2, self.lm, self.ln), dtype=dt, device=self.device)
            out[torch.from_numpy(s.selX).to(self.device)] = mine
            return out
        # sizes of all shards from the (deterministic) partition
        counts = [int(((self._col_global >= c0) & (self._col_global < c1)).sum()) for (c0, c1) in s.ranges]
        nmax = max(counts)
        pad = torch.zeros((nmax, 2, self.lm, self.ln), dtype=dt, device=self.device)
        pad[:mine.shape[0]] = mine
        allx = torch.empty((self.world, nmax, 2, self.lm, self.ln), dtype=dt, device=self.device)
        dist.all_gather_into_tensor(allx, pad)
        out = torch.zeros((s.nnzbX_global, 2, self.lm, self.ln), dtype=dt, device=self.device)
        for r, (c0, c1) in enumerate(s.ranges):
            sel = torch.from_numpy(np.flatnonzero((self._col_global >= c0) & (self._col_global < c1))).to(self.device)
            out[sel] = allx[r, :sel.numel()]
        return out

    def close(self):
        if self.plan is not None:
            self.plan.close()
        self.handle.close()
