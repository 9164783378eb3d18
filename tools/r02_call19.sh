#!/bin/bash
# round-2 call 19: stage 2 in quad form now that k_cascade is LSU-bound (73 % of the wavefront peak); bench's unpipelined block
O=gpurun_out; mkdir -p $O
run() { tag=$1; shift; env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline $BARGS > $O/r02c19_$tag.json 2> $O/r02c19_$tag.err; }
V=$PWD/headtrackr_b200/variants
HT_LIB=$V/libht_quad3.so timeout 600 python -m pytest tests/test_gpu_detect.py tests/test_gpu_quads.py tests/test_gpu_golden.py -m gpu -q --timeout 600 > $O/r02c19_pytest_quad3.log 2>&1; tail -2 $O/r02c19_pytest_quad3.log
BARGS="--workload detect"
run det
run det_quad3 HT_LIB=$V/libht_quad3.so
run det_b
run det_quad3_b HT_LIB=$V/libht_quad3.so
BARGS=""
run full
for f in $O/r02c19_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1].split("r02c19_")[1], round(d["value"]), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), d["kernel_ms_per_step"], d.get("unpipelined",{}).get("value"))
except Exception as e: print(sys.argv[1], "ERR", e); print(open(sys.argv[1].replace('.json','.err')).read()[-1500:])
PY
done
