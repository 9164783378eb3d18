#!/bin/bash
# round-2 call 1: measure the prepared HT_DETECT_PIPE experiment and the whole-step DRAM traffic of the r01 design
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader > $O/r02c1_gpu.txt
for p in 0 2 4 8; do
  HT_DETECT_PIPE=$p python bench.py --steps 5 --warmup 3 --workload detect --no-cpu-baseline > $O/r02c1_pipe$p.json 2> $O/r02c1_pipe$p.err
done
HT_OVERLAP=4 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/r02c1_overlap4.json 2>> $O/r02c1_pipe0.err
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/r02c1_base.json 2>> $O/r02c1_pipe0.err
# whole-step DRAM bytes, 256 frames per step (one step of detect_track30), all kernels
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv \
  --log-file $O/r02c1_dram_step.csv python tools/profile_run.py --frames 256 --iters 1 > $O/r02c1_dram.log 2>&1
for f in $O/r02c1_pipe*.json $O/r02c1_overlap4.json $O/r02c1_base.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1], round(d["value"]), d["ms_per_step"], d["kernel_ms_per_step"])
except Exception as e: print(sys.argv[1], "ERR", e)
PY
done
