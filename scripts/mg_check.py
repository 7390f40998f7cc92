"""torchrun worker: ONE graph sharded over the ranks (one process per GPU, ssb_graph_attach_comm); rank 0 prints a JSON
line with the result digest and, with MG_OUT, saves the estimates for the parity test."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from semantic_slam_b200 import GraphSLAM, synth
from semantic_slam_b200 import distributed as ssbd

name = sys.argv[1] if len(sys.argv) > 1 else "cfg1"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 5
precond = int(sys.argv[3]) if len(sys.argv) > 3 else 3
tol = float(sys.argv[4]) if len(sys.argv) > 4 else 1e-8
repeat = int(sys.argv[5]) if len(sys.argv) > 5 else 0
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
spec = synth.make_config_graph(name)
g = GraphSLAM(device=local, preconditioner=precond, pcg_tol=tol)
synth.load_graph(g, spec)
ssbd.attach(g)
g.prepare()
g.snapshot()
torch.cuda.synchronize(); dist.barrier()
t0 = time.perf_counter()
g.optimize_resident(iters)
torch.cuda.synchronize(); dist.barrier()
dt = time.perf_counter() - t0
P, X = g.get_all(spec.n_poses, spec.n_landmarks)
st = dict(g.stats)
hist = g.history.copy()
ms = []
for _ in range(repeat):
    g.restore()
    g.optimize_resident(iters)
    ms.append(g.stats["ms_device"])
dig = torch.tensor([float(np.abs(P).sum()), float(np.abs(X).sum()), st["chi2_final"], float((P * np.arange(P.size).reshape(P.shape)).sum())],
                   device="cuda", dtype=torch.float64)
lo = dig.clone(); hi = dig.clone()
dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
if rank == 0:
    out = os.environ.get("MG_OUT")
    if out:
        np.savez(out, poses=P, landmarks=X, history=hist)
    print(json.dumps({"world": world, "config": name, "iterations": st["iterations"], "seconds": dt, "ms_device": st["ms_device"],
                      "ms_pcg": st["ms_pcg"], "pcg_iters": st["total_pcg_iters"], "trials": st["total_trials"],
                      "chi2_final": st["chi2_final"], "ranks_identical": bool(torch.equal(lo, hi)),
                      "shard_info": [g.shard_info(world, r) for r in range(world)], "resident_ms": ms,
                      "history_chi2": hist[:, 1].tolist()}), flush=True)
dist.destroy_process_group()
