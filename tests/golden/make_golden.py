#!/usr/bin/env python
"""Generates the committed golden fixtures from the UNMODIFIED reference (run in the build container,
where /root/reference and oracle/_ref exist; the GPU box only sees the generated files).

  plan_unordered.npz / plan_reordered.npz : the reference's own multiplication plans
        test/multiplication/plan_{un,re}ordered.14-287-16 as arrays (starts, pairs, y order)
  FD_problem.xml                          : output of the reference's generate_FD_example (defaults)
  ref_solves.npz                          : results of the reference's CPU build (HAS_NO_CUDA) through
        its own C-ABI on the Julia known-answer test, FD_problem.xml and seeded random systems:
        status, iterations, residual, flops, workspace bytes, the v3 it drew and X (internal layout)

usage: python tests/golden/make_golden.py
"""
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
REF = os.environ.get("TFQMR_REFERENCE", "/root/reference")

import orclib as O  # noqa: E402
from tfqmrgpu_b200 import problems as P  # noqa: E402


def golden_cases():
    """(name, problem, precision, tol, maxit, transA, transB) - shared with the tests."""
    fd = P.read_xml(os.path.join(HERE, "FD_problem.xml"))
    return [
        ("julia_z", P.julia_kat(), "z", 1.2e-8, 210, "n", "n"),
        ("julia_c", P.julia_kat(), "c", 1.2e-5, 210, "n", "n"),
        ("fd_z", fd, "z", fd.tolerance, 2000, "t", "t"),
        ("rand8x8_z", P.random_system(10, 8, 8, seed=11, unsorted=True), "z", 1e-9, 200, "n", "n"),
        ("rand4x5_z", P.random_system(9, 4, 5, seed=12), "z", 1e-9, 200, "n", "n"),
        ("rand16x32_c", P.random_system(6, 16, 32, seed=13), "c", 1e-4, 200, "n", "n"),
    ]


def main():
    # 1. multiplication plans
    for name in ("plan_unordered", "plan_reordered"):
        starts, pairs, nY, nA, nX, yorder = P.read_multiplication_plan(
            os.path.join(REF, "test", "multiplication", name + ".14-287-16"))
        np.savez_compressed(os.path.join(HERE, name + ".npz"), starts=starts, pairs=pairs,
                            nnz=np.array([nY, nA, nX]), yorder=yorder.astype(np.int64))
        print(name, nY, nA, nX, pairs.shape)
    # 2. FD problem from the reference generator
    gen = os.path.join(ROOT, "oracle", "_ref", "generate_FD_example")
    subprocess.check_call([gen], cwd=os.path.join(ROOT, "oracle", "_ref"), stdout=subprocess.DEVNULL)
    os.replace(os.path.join(ROOT, "oracle", "_ref", "FD_problem.xml"), os.path.join(HERE, "FD_problem.xml"))
    # 3. reference CPU solves
    ref = O.ref_cpu()
    out = {}
    for name, prob, prec, tol, maxit, tA, tB in golden_cases():
        dt = np.float64 if prec == "z" else np.float32
        r = ref.solve(prob.mb, prob.lm, prob.ln, prob.A.rowptr, prob.A.colind, P.interleave(prob.A.val, dt),
                      prob.X.rowptr, prob.X.colind, prob.B.rowptr, prob.B.colind, P.interleave(prob.B.val, dt),
                      tol, maxit, prec, transA=tA, trans_b=tB)
        out[name + "_X"] = r["X"]
        out[name + "_v3"] = r["v3"]
        out[name + "_scalars"] = np.array([r["status"], r["iterations"], r["residuum"], r["flops"], r["buffer_size"]], np.float64)
        for k in ("starts", "pairs", "subset", "colindx"):
            out[name + "_" + k] = r["lists"][k]
        print(name, "status", r["status"], "it", r["iterations"], "res", r["residuum"], "flops", r["flops"])
    np.savez_compressed(os.path.join(HERE, "ref_solves.npz"), **out)


if __name__ == "__main__":
    main()
