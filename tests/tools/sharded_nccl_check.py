# Two ranks, one GPU each, NCCL (launched by tests/test_multi_gpu.py through torchrun): column-sharded solve with the global
# iteration rule through the exchange hook, X gathered with NCCL, compared on rank 0 with the single-GPU solve.
import os, sys
os.environ["TFQMRGPU_RESIDENT"] = "0"      # bit-for-bit comparison of the per-kernel path (the resident solver agrees to rounding)
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from tfqmrgpu_b200 import problems as P
from tfqmrgpu_b200.sharded import ShardedBsrsv

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
ok = True
lines = []
for (lm, ln, prec, tol) in ((8, 8, "z", 1e-9), (32, 32, "c", 1e-4)):
    prob = P.random_system(14, lm, ln, ncols=6, pX=1.0, seed=lm, unsorted=True)     # X dense in its block columns: bit for bit
    dt = np.float64 if prec == "z" else np.float32
    vA = P.interleave(prob.A.val, dt); vB = P.interleave(prob.B.val, dt).reshape(prob.B.nnzb, -1)
    args = (prob.mb, lm, ln, prec, prob.A.rowptr, prob.A.colind, vA, "n", prob.X.rowptr, prob.X.colind, prob.B.rowptr, prob.B.colind, vB, "n")
    sh = ShardedBsrsv(*args, rank=rank, world=world, device=dev, dist=dist)
    st = sh.solve(tol, 200)
    it = sh.info()["iterations"]
    X = sh.gather_x(dist).cpu().numpy()
    sh.close()
    if 0 == rank:
        one = ShardedBsrsv(*args, rank=0, world=1, device=dev)
        st1 = one.solve(tol, 200); it1 = one.info()["iterations"]
        X1 = one.gather_x(None).cpu().numpy()
        one.close()
        same = bool(np.array_equal(X, X1))
        lines.append(f"{lm}x{ln}{prec}: status {st}/{st1} iterations {it}/{it1} bit-identical {same}")
        ok = ok and st == st1 == 0 and it == it1 and same
    its = [None]*world
    dist.all_gather_object(its, it)
    ok = ok and len(set(its)) == 1        # every rank ran the same number of iterations
dist.barrier()
if 0 == rank:
    with open(sys.argv[1], "w") as f:
        f.write("\n".join(lines) + ("\nOK\n" if ok else "\nFAILED\n"))
dist.destroy_process_group()
