"""Sharded (keyframe-range) LM over NCCL on >= 2 GPUs vs the single-GPU result and the oracle."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

import oracle
from semantic_slam_b200 import GraphSLAM, synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("world", [2])
def test_sharded_lm_matches_single_gpu_and_oracle(world, tmp_path):
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs")
    out = str(tmp_path / "mg.npz")
    env = dict(os.environ, MG_OUT=out)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29531", os.path.join(ROOT, "scripts", "mg_check.py"), "cfg1", "6"]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1]
    res = json.loads(line)
    assert res["ranks_identical"], "all ranks must hold bit-identical estimates"
    d = np.load(out)
    spec = synth.make_config_graph("cfg1")
    g = GraphSLAM(preconditioner=0)
    o = oracle.OracleGraphSLAM()
    synth.load_graph(g, spec)
    synth.load_graph(o, spec)
    g.optimize(6)
    o.optimize(6)
    P1, X1 = g.get_all(spec.n_poses, spec.n_landmarks)
    Po, Xo = o.get_all(spec.n_poses, spec.n_landmarks)
    assert np.abs(d["poses"] - P1).max() <= 1e-8 * max(1.0, np.abs(P1).max())
    assert np.abs(d["poses"] - Po).max() <= 1e-5 * max(1.0, np.abs(Po).max())
    assert np.abs(d["landmarks"] - Xo).max() <= 1e-5 * max(1.0, np.abs(Xo).max())
    assert np.allclose(d["history"][:, 1], o.history[:, 1], rtol=1e-8)
