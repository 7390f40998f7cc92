"""Writes direct-marginals problems (250 / 1 000 / 4 000 keyframes) from the oracle's linearised system and runs the per-kernel
timing tool scripts/dbg/md_time on them (build it first, see its header).  Run under gpurun."""
import os, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from test_marg_direct import _problem, write_problem

exe = os.path.join(ROOT, "scripts", "dbg", "md_time")
for n_kf, n_lm, seed in ((250, 40, 5), (1000, 100, 6), (4000, 400, 8)):
    pr = _problem(n_kf, n_lm, seed, its=3, flip_every=10**9)
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "p.bin")
        write_problem(pr, np.arange(pr["Nl"]), path)
        print(subprocess.run([exe, path], capture_output=True, text=True, timeout=120).stdout, flush=True)
