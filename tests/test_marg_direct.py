"""The direct form of the landmark marginals (semantic_slam_b200/csrc/ssb_marg_direct.cuh: poses eliminated by a block-bidiagonal
Cholesky of the odometry chain, dense landmark system inverted by block Gauss-Jordan) run on the CPU: tests/md_emulate.cpp
compiles the SAME kernel source and launch sequence with one std::thread per CUDA thread, and the result is compared with the
oracle's marginals (graph_slam.cpp:221-234).  The GPU run of the same code is tests/test_gpu_graph.py."""
import os
import subprocess

import numpy as np
import pytest
import scipy.sparse as sp

import oracle
from semantic_slam_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def emulator(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("md") / "md_emulate")
    subprocess.check_call(["g++", "-std=c++20", "-O2", "-pthread", "-Wall", "-o", exe, os.path.join(ROOT, "tests", "md_emulate.cpp")])
    return exe


def _problem(n_kf, n_lm, seed, its, flip_every=2):
    """device-layout arrays of the linearised system, rebuilt from the oracle's sparse H"""
    spec = synth.make_graph(n_kf, n_lm, seed=synth.SEED_BASE + seed)
    o = oracle.OracleGraphSLAM(threads=1)
    ids = synth.load_graph(o, spec)
    o.optimize(its)
    vk = spec.vkind
    pose_v = [v for v in range(vk.size) if vk[v] == 0]
    lm_v = [v for v in range(vk.size) if vk[v] == 1]
    return problem_from_oracle(o, ids, pose_v, lm_v, flip_every)


def problem_from_oracle(o, ids, pose_v, lm_v, flip_every=2):
    """pose_v / lm_v: positions in `ids` of the keyframe / landmark vertices, keyframes in chain order"""
    H, b, off_by_id = o.sparse_system()
    H = sp.csr_matrix(H)
    off = {v: off_by_id[ids[v]] for v in list(pose_v) + list(lm_v)}
    Np, Nl = len(pose_v), len(lm_v)

    def blk(vr, vc, dr, dc):
        if off[vr] < 0 or off[vc] < 0:
            return np.zeros((dr, dc))
        return H[off[vr]:off[vr] + dr, off[vc]:off[vc] + dc].toarray()

    Hpp = np.stack([blk(v, v, 6, 6) if off[v] >= 0 else np.eye(6) for v in pose_v])           # fixed keyframe: identity
    Hll = np.stack([blk(v, v, 3, 3)[np.triu_indices(3)] for v in lm_v])
    # pose-pose edges (k-1, k); every flip_every-th one stored the other way round to exercise both roles
    pp, inc = [], [[] for _ in range(Np)]
    Hoff = []
    for k in range(1, Np):
        a, c = (k, k - 1) if (k % flip_every == 0) else (k - 1, k)
        e = len(pp)
        pp.append((a, c))
        Hoff.append(blk(pose_v[a], pose_v[c], 6, 6))
        inc[a].append((e << 1 | 0, c))
        inc[c].append((e << 1 | 1, a))
    rowptr = np.zeros(Np + 1, np.int32)
    rowptr[1:] = np.cumsum([len(x) for x in inc])
    idx = np.array([c for x in inc for c, _ in x], np.int32)
    other = np.array([o_ for x in inc for _, o_ in x], np.int32)
    # pose-landmark edges in L-order (landmark-major, then by keyframe): one record per non-zero H(l, p) block
    lm_rowptr, edge_pose, HplL = [0], [], []
    Hc = H.tocsc()
    owner = np.full(H.shape[0], -1, np.int64)                   # scalar row of H -> keyframe
    for p_, pv in enumerate(pose_v):
        if off[pv] >= 0:
            owner[off[pv]:off[pv] + 6] = p_
    for l, v in enumerate(lm_v):
        rows = np.unique(Hc[:, off[v]:off[v] + 3].nonzero()[0])
        ps = np.unique(owner[rows])
        ps = [int(p_) for p_ in ps if p_ >= 0]
        for p in ps:
            edge_pose.append(p)
            HplL.append(blk(v, pose_v[p], 3, 6))
        lm_rowptr.append(len(edge_pose))
    return dict(o=o, ids=ids, lm_v=lm_v, Np=Np, Nl=Nl, rowptr=rowptr, idx=idx, other=other, lm_rowptr=np.array(lm_rowptr, np.int32),
                edge_pose=np.array(edge_pose, np.int32), Hoff=np.array(Hoff), Hpp=Hpp, Hll=Hll, HplL=np.array(HplL))


def write_problem(pr, req, path):
    """the binary layout tests/md_emulate.cpp and scripts/dbg/md_time.cu read"""
    with open(path, "wb") as f:
        np.array([pr["Np"], pr["Nl"], pr["edge_pose"].size, pr["Hoff"].shape[0], pr["idx"].size, req.size], np.int32).tofile(f)
        for a in (pr["rowptr"], pr["idx"], pr["other"], pr["lm_rowptr"], pr["edge_pose"], req.astype(np.int32)):
            a.astype(np.int32).tofile(f)
        for a in (pr["Hoff"], pr["Hpp"], pr["Hll"], pr["HplL"]):
            np.ascontiguousarray(a, dtype=np.float64).tofile(f)


def _run(exe, pr, req, tmp):
    fin, fout = os.path.join(tmp, "p.bin"), os.path.join(tmp, "o.bin")
    write_problem(pr, req, fin)
    out = subprocess.run([exe, fin, fout], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    status = np.fromfile(fout, dtype=np.int32, count=2)
    M = np.fromfile(fout, dtype=np.float64, offset=8).reshape(-1, 3, 3)
    return status, M, out.stdout


@pytest.mark.parametrize("n_kf,n_lm,seed", [(40, 8, 1), (150, 30, 6), (400, 90, 7)])
def test_emulated_kernels_match_the_oracle(emulator, tmp_path, n_kf, n_lm, seed):
    pr = _problem(n_kf, n_lm, seed, its=3)
    req = np.arange(pr["Nl"])[::-1].copy()                     # every landmark, asked in reverse order
    status, M, log = _run(emulator, pr, req, str(tmp_path))
    assert status.tolist() == [0, 0], log
    lmids = np.array([pr["ids"][v] for v in pr["lm_v"]], dtype=np.int32)[req]
    Mo = pr["o"].computeLandmarkMarginals(lmids, method="g2o")
    assert np.abs(M - Mo).max() <= 1e-9 * np.abs(Mo).max(), log


def test_emulated_kernels_with_repeated_edges(emulator, tmp_path):
    """a landmark matched by two detections of one keyframe gives two edges between the same pair (it happens in the per-frame
    stream): the sweep must add both records of a keyframe before moving on"""
    pr = _problem(40, 8, 1, its=3)
    rp, ep, Hl = pr["lm_rowptr"], pr["edge_pose"], pr["HplL"]
    new_rp, new_ep, new_H = [0], [], []
    for l in range(pr["Nl"]):
        for e in range(rp[l], rp[l + 1]):
            if (e % 3) == 0:                                    # every third record split into two unequal parts
                new_ep += [ep[e], ep[e]]
                new_H += [0.25 * Hl[e], 0.75 * Hl[e]]
            else:
                new_ep.append(ep[e])
                new_H.append(Hl[e])
        new_rp.append(len(new_ep))
    ref_status, ref_M, _ = _run(emulator, pr, np.arange(pr["Nl"]), str(tmp_path))
    pr2 = dict(pr, lm_rowptr=np.array(new_rp, np.int32), edge_pose=np.array(new_ep, np.int32), HplL=np.array(new_H))
    status, M, log = _run(emulator, pr2, np.arange(pr["Nl"]), str(tmp_path))
    assert status.tolist() == [0, 0] and ref_status.tolist() == [0, 0], log
    assert np.abs(M - ref_M).max() <= 1e-12 * np.abs(ref_M).max()


def test_emulated_kernels_flag_a_singular_system(emulator, tmp_path):
    pr = _problem(40, 8, 1, its=2)
    pr["Hll"][3] = 0.0                                          # a landmark whose block vanishes
    pr["HplL"][pr["lm_rowptr"][3]:pr["lm_rowptr"][4]] = 0.0
    status, M, log = _run(emulator, pr, np.arange(2), str(tmp_path))
    assert status[1] != 0, log


def test_emulated_kernels_are_race_free_under_thread_sanitizer(tmp_path):
    """the emulator gives every CUDA thread an OS thread and every __syncthreads() / shuffle a barrier, so ThreadSanitizer sees
    exactly the happens-before edges the CUDA code has inside a CTA: a missing barrier between conflicting shared-memory (or
    same-CTA global-memory) accesses would be reported as a data race"""
    exe = str(tmp_path / "md_emulate_tsan")
    r = subprocess.run(["g++", "-std=c++20", "-O1", "-g", "-pthread", "-fsanitize=thread", "-o", exe, os.path.join(ROOT, "tests", "md_emulate.cpp")],
                       capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("ThreadSanitizer is not available: " + r.stderr[-200:])
    pr = _problem(150, 30, 6, its=3)                            # 2 x 2 tiles: the tile product, both Gauss-Jordan branches
    fin, fout = str(tmp_path / "p.bin"), str(tmp_path / "o.bin")
    write_problem(pr, np.arange(pr["Nl"]), fin)
    env = dict(os.environ, TSAN_OPTIONS="halt_on_error=0 report_signal_unsafe=0 exitcode=66")
    out = subprocess.run([exe, fin, fout], capture_output=True, text=True, timeout=900, env=env)
    assert "ThreadSanitizer" not in out.stderr and out.returncode == 0, out.stderr[-2000:]
    assert "status 0 0" in out.stdout
