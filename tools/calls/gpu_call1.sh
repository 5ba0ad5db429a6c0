#!/bin/bash
# round 2, GPU call 1: TF probe, full GPU test suite, smoke, bench N=1
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( python -c "import tensorflow as tf; print('tensorflow', tf.__version__)" 2>&1 | tail -1; python -c "import keras; print('keras', keras.__version__)" 2>&1 | tail -1; python -c "import polars; print('polars', polars.__version__)" 2>&1 | tail -1; python --version; nproc; nvidia-smi --query-gpu=name,driver_version --format=csv,noheader ) > gpurun_out/tf_probe.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -s -p no:cacheprovider > gpurun_out/gputest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/gputest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
tail -3 gpurun_out/gputest.log; cat gpurun_out/smoke.log | tail -2; cat gpurun_out/bench_n1.json | head -c 1500
