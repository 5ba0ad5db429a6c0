"""One eager train step of the headline workload between cudaProfilerStart/Stop, for ncu:

  ncu --profile-from-start off --set full --clock-control none --import-source on -o gpurun_out/r02_step \
      python tools/profile_step.py [--workload NAME] [--dp-path]
  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
      python tools/profile_step.py

(CUDA-graph replay is switched off so that every kernel is a separate, named launch.)
"""
import argparse
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "ebnerd-benchmark_b200")]
os.environ["EBK_NO_GRAPH"] = "1"

import torch  # noqa: E402

import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default=bench.DEFAULT_WORKLOAD)
ap.add_argument("--steps", type=int, default=1)
ap.add_argument("--no-defer", action="store_true", help="keep the wgrad GEMM on the main stream (serial kernel list)")
args = ap.parse_args()
if args.no_defer:
    os.environ["EBK_DEFER_WGRAD"] = "0"
w = bench.WORKLOADS[args.workload]
torch.cuda.set_device(0)
model, eng, host, dev, B, C_ = bench.build_model(w, 0)
for i in range(3):
    eng.train_step_dev(dev[i % len(dev)][0], dev[i % len(dev)][1], B, C_)
torch.cuda.synchronize()
torch.cuda.profiler.start()
for i in range(args.steps):
    eng.train_step_dev(dev[(3 + i) % len(dev)][0], dev[(3 + i) % len(dev)][1], B, C_)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled", args.steps, "step(s) of", args.workload)
