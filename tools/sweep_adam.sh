#!/bin/bash
# sweep of the fused embedding-Adam kernel's slice width (CH) and grid cap on the bench workload
for ch in 1 2 3 6; do for cap in 4 8 16; do
  r=$(EBK_ADAM_CH=$ch EBK_ADAM_CAP=$cap python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['kernel_ms_per_step']['news.adam'], d['ms_per_step'])")
  echo "CH=$ch CAP=$cap adam_ms,step_ms = $r"
done; done
