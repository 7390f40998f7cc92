"""CPU coverage of the N>1 path (no GPU): the keyframe-range sharding plan the C library derives on every rank
(ssb_shard_plan — the same code ssb_graph_prepare runs), checked for its invariants and, over gloo with world_size 2,
for sufficiency: every rank applies its rows of the Schur complement using ONLY the keyframes the plan gives it (own +
ghosts) and the PCG built on that distributed product must reproduce the dense solve."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_ranges_partition():
    from semantic_slam_b200.distributed import shard_ranges
    for Np, world in [(100, 2), (10000, 8), (7, 4), (5, 8), (1, 1)]:
        covered = []
        for r in range(world):
            ps, pe = shard_ranges(Np, 0, world, r)
            assert 0 <= ps <= pe <= Np
            covered += list(range(ps, pe))
        assert covered == list(range(Np)), "keyframe ranges must tile [0, Np) contiguously"


@pytest.mark.parametrize("name,world", [("cfg1", 2), ("cfg1", 4), ("cfg2", 8)])
def test_plan_invariants(name, world):
    from semantic_slam_b200 import synth
    from semantic_slam_b200.distributed import shard_plan, spec_index_lists
    spec = synth.make_config_graph(name)
    Np, Nl, pl_p, pl_l, pp_i, pp_j = spec_index_lists(spec)
    plans = [shard_plan(Np, Nl, pl_p, pl_l, pp_i, pp_j, world, r) for r in range(world)]
    owner = np.zeros(Np, dtype=int)
    for r, P in enumerate(plans):
        owner[P["own"][0]:P["own"][1]] = r
    assert sum(P["owned_landmarks"] for P in plans) == Nl, "every landmark is eliminated by exactly one rank"
    for s, P in enumerate(plans):
        ghosts = P["ghosts"]
        assert P["local_poses"] == P["own"][1] - P["own"][0] + ghosts.size
        assert np.all(owner[ghosts] != s) and np.all(np.diff(ghosts) > 0)
        # whoever owns a ghost of rank s pushes it to s, and nothing else
        for r, Q in enumerate(plans):
            if r != s:
                assert Q["push_to"][s] == int((owner[ghosts] == r).sum())
        # every keyframe that shares a landmark with an own keyframe, or is a pose-pose neighbour, is local
        own = np.zeros(Np, dtype=bool)
        own[P["own"][0]:P["own"][1]] = True
        touched = np.zeros(Nl, dtype=bool)
        touched[pl_l[own[pl_p]]] = True
        need = np.zeros(Np, dtype=bool)
        need[pl_p[touched[pl_l]]] = True
        need[pp_j[own[pp_i]]] = True
        need[pp_i[own[pp_j]]] = True
        need &= ~own
        assert np.array_equal(np.flatnonzero(need), ghosts)
        assert P["touched_landmarks"] >= int(touched.sum())
        assert P["local_edges"] == int(touched[pl_l].sum()) + 0 * P["u_pushes"]
    assert sum(P["u_pushes"] for P in plans) == sum(P["ghosts"].size for P in plans)


def _worker(rank, world, port, outq):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import oracle
    from semantic_slam_b200 import synth
    from semantic_slam_b200.distributed import shard_plan, spec_index_lists
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
    Nfree, Nl = len(pidx) // 6, len(lidx) // 3
    Winv = np.zeros_like(Hll)
    for l in range(Nl):
        Winv[3 * l:3 * l + 3, 3 * l:3 * l + 3] = np.linalg.inv(Hll[3 * l:3 * l + 3, 3 * l:3 * l + 3])
    S = Hpp - Hpl @ Winv @ Hpl.T
    g = b[pidx] - Hpl @ (Winv @ b[lidx])
    # the plan of the product, for the full graph (keyframe 0 is fixed: free keyframe k = keyframe k + 1)
    Np, Nl_, pl_p, pl_l, pp_i, pp_j = spec_index_lists(spec)
    P = shard_plan(Np, Nl_, pl_p, pl_l, pp_i, pp_j, world, rank)
    ps, pe = P["own"]
    local = np.zeros(Np, dtype=bool)
    local[ps:pe] = True
    local[P["ghosts"]] = True
    own_free = np.arange(max(ps, 1), pe) - 1                      # free-keyframe indices this rank owns
    visible = np.flatnonzero(local[1:])                           # free keyframes whose values this rank may read
    rows = (6 * own_free[:, None] + np.arange(6)).reshape(-1)
    cols = (6 * visible[:, None] + np.arange(6)).reshape(-1)
    counts = [None] * world
    dist.all_gather_object(counts, int(rows.size))

    def exchange(vec_own):
        """every rank contributes its own rows; a rank then reads ONLY what the plan makes local to it"""
        parts = [None] * world
        dist.all_gather_object(parts, (rows, vec_own))
        full = np.full(6 * Nfree, np.nan)
        for rr, vv in parts:
            full[rr] = vv
        seen = np.full(6 * Nfree, np.nan)
        seen[cols] = full[cols]
        return seen

    def allreduce(x):
        t = torch.tensor([x], dtype=torch.float64)
        dist.all_reduce(t)
        return float(t[0])

    Dinv = [np.linalg.inv(S[6 * i:6 * i + 6, 6 * i:6 * i + 6]) for i in own_free]
    appM = lambda r_own: np.concatenate([Dinv[k] @ r_own[6 * k:6 * k + 6] for k in range(len(Dinv))]) if len(Dinv) else r_own
    Srows = S[rows]
    support = np.flatnonzero(np.abs(Srows).sum(0) > 0)
    assert np.all(np.isin(support, cols)), "the plan's ghosts must cover the support of this rank's rows of S"
    x = np.zeros(rows.size); r = g[rows].copy(); z = appM(r); p_own = z.copy()
    rz = allreduce(float(r @ z)); rz0 = rz
    it = 0
    for it in range(3000):
        p_seen = exchange(p_own)
        q = Srows[:, cols] @ p_seen[cols]
        assert np.all(np.isfinite(q))
        pq = allreduce(float(p_own @ q))
        a = rz / pq
        x += a * p_own; r -= a * q; z = appM(r)
        rzn = allreduce(float(r @ z))
        p_own = z + (rzn / rz) * p_own
        rz = rzn
        if rz <= 1e-24 * rz0:
            break
    parts = [None] * world
    dist.all_gather_object(parts, (rows, x))
    xfull = np.zeros(6 * Nfree)
    for rr, vv in parts:
        xfull[rr] = vv
    ref = np.linalg.solve(S, g)
    err = float(np.abs(xfull - ref).max() / max(1.0, np.abs(ref).max()))
    outq.put((rank, err, it))
    dist.destroy_process_group()


def test_sharded_pcg_on_the_plan_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    world, port = 2, 29611
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, err, it in res:
        assert err < 1e-8, (rank, err, it)
