"""CPU tests of the measurement plumbing: the device-side generator's hash is bit-identical to the numpy one (so the CPU
sample and the GPU workload are the same operator), shard parameters, and the JSON contract of `bench.py --impl reference`
(the reference's CPU path on the bounded sample)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from tfqmrgpu_b200 import problems as P, synthetic

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_torch_hash_matches_numpy_hash_bit_for_bit():
    d = P.stencil27(4, 8, 8, 2, sigma=1.0)
    sp = synthetic.Stencil27(4, 8, 8, 2, sigma=1.0, device="cpu", pin=False)
    assert np.array_equal(sp.valA_host.numpy(), d["valA"])
    for k in ("rpA", "ciA", "rpX", "ciX", "rpB", "ciB", "valB"):
        assert np.array_equal(getattr(sp, k), d[k]), k
    d64 = P.stencil27(3, 4, 5, 1, sigma=8.0, dtype=np.float64)
    sp64 = synthetic.Stencil27(3, 4, 5, 1, sigma=8.0, dtype=np.float64, device="cpu", pin=False)
    assert np.array_equal(sp64.valA_host.numpy(), d64["valA"])


def test_shard_of_a_wider_problem():
    sp = synthetic.Stencil27(4, 8, 8, 2, device="cpu", pin=False, col0=2, ncols_global=8, with_values=False)
    assert sp.valA_host is None and sp.a_bytes == sp.nnzbA*8*8*2*4
    assert sorted(set(sp.ciX.tolist())) == [2, 3] and sorted(sp.ciB.tolist()) == [2, 3]
    assert np.flatnonzero(np.diff(sp.rpB)).tolist() == [16, 24]          # unit blocks of columns 2, 3 in rows c * (64 // 8)
    diag5 = int(sp.rpA[5] + np.flatnonzero(sp.ciA[sp.rpA[5]:sp.rpA[6]] == 5)[0])      # diagonal block of block row 5
    off5 = diag5 + 1 if diag5 + 1 < sp.rpA[6] else diag5 - 1
    rows = P.stencil27_values_rows(sp.rpA, sp.ciA, 8, 8.0, np.array([diag5, off5]))
    assert rows.shape == (2, 8, 8, 2) and abs(rows[0, 0, 0, 0] - 35.0) < 0.06 and abs(rows[1, 0, 0, 0] + 1.0) < 0.06
    assert abs(rows[0, 0, 1, 0]) < 0.06 and abs(rows[0, 3, 3, 1]) < 0.06                  # 0.05 * uniform noise elsewhere


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "liboracle.so")), reason="oracle not built")
def test_reference_arm_json_contract():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "tfqmr_solve_throughput" and line["unit"] == "GFLOP/s"
    assert line["higher_is_better"] is True and line["value"] > 0 and line["vs_baseline"] is None
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] == 1
    assert line["e2e"] == {"value": line["value"], "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in line["config"] and line["gpu_launches"] == 0
    # other ranks of a torchrun launch exit silently
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=60, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""
