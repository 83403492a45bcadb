"""Several GPUs behind the C-ABI (the reference itself has no multi-GPU code).

  * one process, several devices: ``tfqmrgpux_bsrsv_setDevices`` / ``TFQMRGPU_NUM_GPUS`` - the ordinary entry points drive
    all shards, the caller keeps ONE workspace.  Device lists may name a device twice, so these tests also run (and are run by
    the driver) on a box with a single GPU: every shard then has its own stream, workspace and sub-plan on that GPU and the
    whole exchange machinery (mapped host slots, cross-stream events, decide kernel, device-to-device pushes) is exercised;
  * one process per GPU: ``tfqmrgpux_bsrsv_setShardExchange`` with NCCL (``tfqmrgpu_b200/sharded.py``), needs two GPUs.

The bar: a column-sharded solve reproduces the single-GPU run - status, iteration count (the reference's rule is GLOBAL,
core.hxx:239-299), flop count, residual, per-RHS status - and X bit for bit.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

from tfqmrgpu_b200 import api, problems as P, _lib as L

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.fixture(autouse=True)
def _per_kernel_path(monkeypatch):
    """The bit-for-bit comparisons below are about the sharding machinery: both sides run the per-kernel path.  (A small system
    on ONE GPU would otherwise take the resident solver, which groups its sums differently and agrees to rounding only:
    test_gpu_parity.py::test_resident_solver_matches_the_per_kernel_path.)"""
    monkeypatch.setenv("TFQMRGPU_RESIDENT", "0")


def _run(prob, prec, tol, maxit, devices=None, env_ngpu=None, monkeypatch=None, layout=L.LAYOUT_RIRIRIRI):
    dt = np.float64 if prec == "z" else np.float32
    vA = P.interleave(prob.A.val, dt); vB = P.interleave(prob.B.val, dt)
    if env_ngpu:
        monkeypatch.setenv("TFQMRGPU_NUM_GPUS", str(env_ngpu))
    h = api.Handle()
    pl = api.BsrsvPlan(h, prob.mb, prob.A.rowptr, prob.A.colind, prob.X.rowptr, prob.X.colind, prob.B.rowptr, prob.B.colind)
    if env_ngpu:
        monkeypatch.delenv("TFQMRGPU_NUM_GPUS")
    if devices is not None:
        pl.set_devices(len(devices), devices)
    ndev = len(pl.get_devices())
    pl.buffer_size_for(prob.lm, prob.ln, prec); pl.set_buffer()
    pl.set_matrix("A", vA); pl.set_matrix("B", vB)
    st = pl.solve(tol, maxit)
    info = pl.info()
    X = pl.get_matrix("X", "n", layout).copy()
    rhs = pl.rhs_status()
    pl.close(); h.close()
    return dict(st=st, info=info, X=X, rhs=rhs, ndev=ndev)


def _same_device_list(n):
    import torch
    have = torch.cuda.device_count()
    return [i % have for i in range(n)]


@pytest.mark.parametrize("lm,ln,prec,ncols,tol", [(8, 8, "z", 5, 1e-9), (32, 32, "c", 4, 1e-4), (16, 32, "z", 3, 1e-9), (4, 4, "c", 7, 1e-4),
                                                  (64, 64, "c", 2, 1e-3)],
                         ids=["8x8z", "32x32c", "16x32z", "4x4c", "64x64c"])
@pytest.mark.parametrize("nshards", [2, 3])
def test_set_devices_reproduces_the_single_gpu_run(lm, ln, prec, ncols, tol, nshards):
    """X dense in its block columns (every block row holds all of them - the reference's use case): bit for bit."""
    prob = P.random_system(14, lm, ln, ncols=ncols, pX=1.0, seed=lm + 3*ncols, unsorted=True)
    one = _run(prob, prec, tol, 200)
    many = _run(prob, prec, tol, 200, devices=_same_device_list(nshards))
    assert one["ndev"] == 0 and many["ndev"] == min(nshards, ncols)
    assert many["st"] == one["st"] == 0
    assert many["info"]["iterations"] == one["info"]["iterations"]            # the global rule, not a per-shard one
    assert many["info"]["flops"] == one["info"]["flops"]
    assert many["info"]["residuum"] == one["info"]["residuum"]
    assert np.array_equal(many["rhs"], one["rhs"])
    assert np.array_equal(many["X"], one["X"])                               # bit for bit


@pytest.mark.parametrize("lm,ln,prec,tol", [(8, 8, "z", 1e-9), (32, 32, "c", 1e-4), (4, 4, "c", 1e-4)], ids=["8x8z", "32x32c", "4x4c"])
def test_set_devices_ragged_x_pattern(lm, ln, prec, tol):
    """Block rows of X with different block columns: the product kernels group a row's entries by the columns that share a
    unit, so the shards' sums may be ordered differently - same status, iterations within one, X equal to rounding."""
    prob = P.random_system(14, lm, ln, ncols=5, seed=lm + 15, unsorted=True)
    one = _run(prob, prec, tol, 200)
    many = _run(prob, prec, tol, 200, devices=_same_device_list(2))
    assert many["st"] == one["st"] == 0 and abs(many["info"]["iterations"] - one["info"]["iterations"]) <= 1
    assert np.array_equal(many["rhs"], one["rhs"])
    assert np.abs(many["X"] - one["X"]).max() <= 10*tol*np.abs(one["X"]).max()


def test_num_gpus_environment_variable_and_second_solve(monkeypatch):
    prob = P.random_system(12, 8, 8, ncols=4, pX=1.0, seed=21, unsorted=True)
    one = _run(prob, "z", 1e-9, 200)
    import torch
    n = min(2, torch.cuda.device_count())
    if n < 2:
        pytest.skip("TFQMRGPU_NUM_GPUS names devices 0..N-1: needs two GPUs")
    many = _run(prob, "z", 1e-9, 200, env_ngpu=2, monkeypatch=monkeypatch)
    assert many["ndev"] == 2 and many["st"] == 0 and many["info"]["iterations"] == one["info"]["iterations"]
    assert np.array_equal(many["X"], one["X"])


def test_multi_device_max_iterations_status_and_layouts():
    """status 9 with iterations_needed = MaxIt like the reference (core.hxx:170-171), other download layouts, a second solve."""
    prob = P.random_system(12, 8, 8, ncols=4, pX=1.0, seed=5, unsorted=True)
    vA = P.interleave(prob.A.val, np.float64); vB = P.interleave(prob.B.val, np.float64)
    outs = []
    for devices in (None, _same_device_list(2)):
        h = api.Handle()
        pl = api.BsrsvPlan(h, prob.mb, prob.A.rowptr, prob.A.colind, prob.X.rowptr, prob.X.colind, prob.B.rowptr, prob.B.colind)
        if devices:
            pl.set_devices(len(devices), devices)
        pl.buffer_size_for(8, 8, "z"); pl.set_buffer()
        pl.set_matrix("A", vA); pl.set_matrix("B", vB)
        st3 = pl.solve(1e-9, 3); i3 = pl.info()
        x3 = pl.get_matrix("X", "t", L.LAYOUT_RRIIRRII).copy()
        st = pl.solve(1e-9, 200); i = pl.info()
        x = pl.get_matrix("X", "n", L.LAYOUT_RRRRIIII).copy()
        outs.append((st3, i3["iterations"], x3, st, i["iterations"], x, i["flops_all"]))
        pl.close(); h.close()
    a, b = outs
    assert a[0] == b[0] == L.STATUS_MAX_ITERATIONS and a[1] == b[1] == 3 and np.array_equal(a[2], b[2])
    assert a[3] == b[3] == 0 and a[4] == b[4] and np.array_equal(a[5], b[5]) and a[6] == b[6]


def test_two_real_devices_strong_scaling_problem():
    """Two physical GPUs: a stencil problem with 4 block columns split 2 + 2, against the same problem on one GPU."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from tfqmrgpu_b200 import synthetic
    sp = synthetic.Stencil27(8, 32, 32, 4, sigma=8.0, dtype=np.float32, device="cuda")
    res = []
    for devices in (None, [0, 1]):
        h = api.Handle()
        pl = api.BsrsvPlan(h, sp.mb, sp.rpA, sp.ciA, sp.rpX, sp.ciX, sp.rpB, sp.ciB)
        if devices:
            pl.set_devices(2, devices)
        pl.buffer_size_for(32, 32, "c"); pl.set_buffer()
        pl.set_matrix("A", None, "n", raw_ptr=sp.valA_host.data_ptr()); pl.set_matrix("B", sp.valB)
        st = pl.solve(1e-3, 100)
        res.append((st, pl.info(), pl.get_matrix("X").copy()))
        pl.close(); h.close()
    assert res[0][0] == res[1][0] == 0 and res[0][1]["iterations"] == res[1][1]["iterations"]
    assert np.array_equal(res[0][2], res[1][2])


def test_one_process_per_gpu_with_nccl_exchange_and_gather(tmp_path):
    """torchrun, 2 ranks, NCCL: tfqmrgpu_b200/sharded.py registers the exchange (all-gather of the shards' convergence monitors
    on the solver's stream) and gathers X with NCCL; rank 0 compares with its own single-GPU solve of the whole problem."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    out = tmp_path / "result.txt"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29731", os.path.join(HERE, "tools", "sharded_nccl_check.py"), str(out)]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    text = out.read_text()
    assert "OK" in text, text
