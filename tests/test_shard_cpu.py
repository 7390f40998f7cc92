"""CPU (gloo, world_size 2) coverage of the N>1 path: the keyframe-range partition of the C library
and the collective pattern of the sharded Schur PCG (all-gather p / v, all-reduce the two dot
products), emulated in numpy on the oracle's system so it runs without a GPU."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_ranges_partition():
    from semantic_slam_b200.distributed import shard_ranges
    for Np, Nl, world in [(100, 20, 2), (10000, 1986, 8), (7, 3, 4), (5, 0, 8), (1, 1, 1)]:
        covered_p, covered_l = [], []
        for r in range(world):
            ps, pe, ls, le = shard_ranges(Np, Nl, world, r)
            assert 0 <= ps <= pe <= Np and 0 <= ls <= le <= Nl
            covered_p += list(range(ps, pe))
            covered_l += list(range(ls, le))
        assert covered_p == list(range(Np)), "keyframe ranges must tile [0, Np) contiguously"
        assert covered_l == list(range(Nl))


def _worker(rank, world, port, outq):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import scipy.sparse as sp
    import oracle
    from semantic_slam_b200 import synth
    from semantic_slam_b200.distributed import shard_ranges
    dist.init_process_group("gloo", rank=rank, world_size=world)
    spec = synth.make_config_graph("cfg1")
    o = oracle.OracleGraphSLAM()
    synth.load_graph(o, spec)
    H, b, off = o.sparse_system()
    lam = 0.5
    pidx, lidx = [], []
    for v in range(spec.vkind.size):
        if off[v] < 0:
            continue
        (pidx if spec.vkind[v] == 0 else lidx).extend(range(off[v], off[v] + (6 if spec.vkind[v] == 0 else 3)))
    pidx, lidx = np.array(pidx), np.array(lidx)
    Hpp = H[pidx][:, pidx].toarray() + lam * np.eye(len(pidx))
    Hpl = H[pidx][:, lidx].toarray()
    Hll = H[lidx][:, lidx].toarray() + lam * np.eye(len(lidx))
    Np, Nl = len(pidx) // 6, len(lidx) // 3
    Winv = np.zeros_like(Hll)
    for l in range(Nl):
        Winv[3 * l:3 * l + 3, 3 * l:3 * l + 3] = np.linalg.inv(Hll[3 * l:3 * l + 3, 3 * l:3 * l + 3])
    S = Hpp - Hpl @ Winv @ Hpl.T
    g = b[pidx] - Hpl @ (Winv @ b[lidx])
    ps, pe, ls, le = shard_ranges(Np, Nl, world, rank)
    cp, cl = -(-Np // world), -(-Nl // world)

    def allgather(vec_owned, chunk, width, total):
        buf = torch.zeros(chunk * width, dtype=torch.float64)
        buf[: vec_owned.size] = torch.from_numpy(vec_owned)
        outs = [torch.zeros_like(buf) for _ in range(world)]
        dist.all_gather(outs, buf)
        return torch.cat(outs).numpy()[: total * width]

    def allreduce(x):
        t = torch.tensor([x], dtype=torch.float64)
        dist.all_reduce(t)
        return float(t[0])

    rows = slice(6 * ps, 6 * pe)
    Dinv = np.zeros((Np, 6, 6))
    for i in range(Np):
        Dinv[i] = np.linalg.inv(S[6 * i:6 * i + 6, 6 * i:6 * i + 6])
    appM = lambda r_own: np.einsum("nij,nj->ni", Dinv[ps:pe], r_own.reshape(-1, 6)).reshape(-1)
    x = np.zeros(6 * (pe - ps)); r = g[rows].copy(); z = appM(r)
    p = allgather(z, cp, 6, Np)
    rz = allreduce(float(r @ z)); rz0 = rz
    for it in range(2000):
        # phase 1: owned landmarks  v = W Hlp p ; all-gather v
        v_own = (Winv[3 * ls:3 * le, 3 * ls:3 * le] @ (Hpl[:, 3 * ls:3 * le].T @ p))
        v = allgather(v_own, cl, 3, Nl)
        # phase 2: owned poses q = (Hpp) p - Hpl v ; all-reduce p.q
        q = Hpp[rows] @ p - Hpl[rows] @ v
        pq = allreduce(float(p[rows] @ q))
        a = rz / pq
        x += a * p[rows]; r -= a * q; z = appM(r)
        rzn = allreduce(float(r @ z))
        p_own = z + (rzn / rz) * p[rows]
        p = allgather(p_own, cp, 6, Np)
        rz = rzn
        if rz <= 1e-24 * rz0:
            break
    xfull = allgather(x, cp, 6, Np)
    ref = np.linalg.solve(S, g)
    err = float(np.abs(xfull - ref).max() / max(1.0, np.abs(ref).max()))
    outq.put((rank, err, it))
    dist.destroy_process_group()


def test_sharded_pcg_collective_pattern_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    world, port = 2, 29611
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, err, it in res:
        assert err < 1e-8, (rank, err, it)
