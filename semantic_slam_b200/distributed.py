"""Multi-GPU plumbing (one process per GPU): torch.distributed carries the NCCL unique id to every rank, the C
library builds its own communicator from it — used only to hand over the cudaIpc handles of the peer arenas — and
shards the graph by contiguous keyframe range (include/ssb.h: ssb_graph_attach_comm; csrc/ssb_peer.cuh).  Every
rank must hold the same graph (replay the same add_* calls) and call prepare / optimize / chi2 collectively."""
from __future__ import annotations

import ctypes as C
import numpy as np

from . import _lib
from ._lib import ip, check


def shard_ranges(n_poses: int, n_landmarks: int, world: int, rank: int):
    """(ps, pe): keyframe range [ps, pe) owned by `rank`."""
    out = np.zeros(4, dtype=np.int32)
    check(_lib.lib().ssb_shard_ranges(n_poses, n_landmarks, world, rank, out.ctypes.data_as(ip)), "ssb_shard_ranges")
    return int(out[0]), int(out[1])


def shard_plan(n_poses, n_landmarks, pl_pose, pl_lm, pp_i, pp_j, world, rank):
    """The sharding plan of `rank` as ssb_graph_prepare derives it (host-only; include/ssb.h: ssb_shard_plan)."""
    a = [np.ascontiguousarray(x, dtype=np.int32) for x in (pl_pose, pl_lm, pp_i, pp_j)]
    out = np.zeros(8, dtype=np.int32)
    ghosts = np.zeros(max(n_poses, 1), dtype=np.int32)
    push_to = np.zeros(world, dtype=np.int32)
    check(_lib.lib().ssb_shard_plan(n_poses, n_landmarks, a[0].ctypes.data_as(ip), a[1].ctypes.data_as(ip), a[0].size,
                                    a[2].ctypes.data_as(ip), a[3].ctypes.data_as(ip), a[2].size, world, rank,
                                    out.ctypes.data_as(ip), ghosts.ctypes.data_as(ip), push_to.ctypes.data_as(ip)),
          "ssb_shard_plan")
    n_own = int(out[1] - out[0])
    return {"own": (int(out[0]), int(out[1])), "local_poses": int(out[2]), "owned_landmarks": int(out[3]),
            "touched_landmarks": int(out[4]), "local_edges": int(out[5]), "u_pushes": int(out[6]), "v_pushes": int(out[7]),
            "ghosts": ghosts[: int(out[2]) - n_own].copy(), "push_to": push_to}


def spec_index_lists(spec):
    """(n_poses, n_landmarks, pl_pose, pl_lm, pp_i, pp_j) of a synth.GraphSpec: per-kind indices in creation order."""
    kind_idx = np.zeros(spec.vkind.size, dtype=np.int64)
    kind_idx[spec.vkind == 0] = np.arange(int((spec.vkind == 0).sum()))
    kind_idx[spec.vkind == 1] = np.arange(int((spec.vkind == 1).sum()))
    pl = spec.ekind == 1
    pp = spec.ekind == 0
    return (spec.n_poses, spec.n_landmarks, kind_idx[spec.evi[pl]], kind_idx[spec.evj[pl]], kind_idx[spec.evi[pp]],
            kind_idx[spec.evj[pp]])


def attach(graph, group=None):
    """Attach an NCCL communicator spanning the torch.distributed group to `graph` (GraphSLAM)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if world == 1:
        graph.attach_comm(0, 1, b"\0" * 128)
        return
    buf = C.create_string_buffer(128)
    if rank == 0:
        check(_lib.lib().ssb_comm_unique_id(buf), "ssb_comm_unique_id")
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    t = torch.tensor(list(buf.raw), dtype=torch.uint8, device=dev)
    dist.broadcast(t, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    graph.attach_comm(rank, world, bytes(t.cpu().tolist()))
