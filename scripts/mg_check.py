"""torchrun worker: sharded LM on N GPUs; rank 0 prints a JSON line with the result digest."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from semantic_slam_b200 import GraphSLAM, synth
from semantic_slam_b200 import distributed as ssbd

name = sys.argv[1] if len(sys.argv) > 1 else "cfg1"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 5
scale = int(sys.argv[3]) if len(sys.argv) > 3 else 1
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
spec = synth.make_config_graph(name, scale) if name in synth.CONFIGS else None
g = GraphSLAM(device=local, preconditioner=0)
synth.load_graph(g, spec)
ssbd.attach(g)
g.prepare()
torch.cuda.synchronize(); dist.barrier()
t0 = time.perf_counter()
g.optimize_resident(iters)
torch.cuda.synchronize(); dist.barrier()
dt = time.perf_counter() - t0
P, X = g.get_all(spec.n_poses, spec.n_landmarks)
dig = torch.tensor([float(np.abs(P).sum()), float(np.abs(X).sum()), g.stats["chi2_final"]], device="cuda", dtype=torch.float64)
lo = dig.clone(); hi = dig.clone()
dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
if rank == 0:
    out = os.environ.get("MG_OUT")
    if out:
        np.savez(out, poses=P, landmarks=X, history=g.history)
    print(json.dumps({"world": world, "config": name, "scale": scale, "iterations": g.iterations, "seconds": dt,
                      "ms_device": g.stats["ms_device"], "pcg_iters": g.stats["total_pcg_iters"],
                      "chi2_final": g.stats["chi2_final"], "ranks_identical": bool(torch.equal(lo, hi)),
                      "history_chi2": g.history[:, 1].tolist()}), flush=True)
dist.destroy_process_group()
