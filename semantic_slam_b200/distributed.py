"""Multi-GPU plumbing (one process per GPU): torch.distributed carries the NCCL unique id to every
rank, the C library builds its own communicator from it and shards the Schur PCG by contiguous
keyframe range (include/ssb.h: ssb_graph_attach_comm, ssb_shard_ranges).  Every rank must hold the
same graph (replay the same add_* calls)."""
from __future__ import annotations

import ctypes as C
import numpy as np

from . import _lib
from ._lib import ip, check


def shard_ranges(n_poses: int, n_landmarks: int, world: int, rank: int):
    """(ps, pe, ls, le): keyframe range [ps, pe) and landmark range [ls, le) owned by `rank`."""
    out = np.zeros(4, dtype=np.int32)
    check(_lib.lib().ssb_shard_ranges(n_poses, n_landmarks, world, rank, out.ctypes.data_as(ip)), "ssb_shard_ranges")
    return tuple(int(v) for v in out)


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
    dist.broadcast(t, src=0, group=group)
    graph.attach_comm(rank, world, bytes(t.cpu().tolist()))
