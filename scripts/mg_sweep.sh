#!/bin/bash
# builder-run multi-GPU evidence (the driver's test box has one GPU): N ranks = N GPUs, one process each
N=$1
OUT=gpurun_out/r02_mg$N
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29551 scripts/mg_check.py cfg2 20 3 1e-6 3 > $OUT/cfg2.log 2>&1; echo rc=$? >> $OUT/cfg2.log
timeout 500 $TR --master-port 29552 scripts/mg_check.py cfg4 5 3 1e-6 2 > $OUT/cfg4.log 2>&1; echo rc=$? >> $OUT/cfg4.log
timeout 600 python -m pytest tests/test_multi_gpu.py -x -q -m gpu -k "processes and $N" > $OUT/pytest.log 2>&1; echo rc=$? >> $OUT/pytest.log
timeout 600 $TR --master-port 29553 bench.py --gpus $N --steps 5 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo rc=$? >> $OUT/bench.err
grep -h "^{\|rc=" $OUT/cfg2.log $OUT/cfg4.log | cut -c1-700; tail -3 $OUT/pytest.log; tail -c 600 $OUT/bench.json; tail -2 $OUT/bench.err
