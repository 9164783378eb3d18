#!/bin/bash
# round-2 call 20: PRMT byte extraction in k_resample; survivor groups split ({4} {5}, {6} {7}); new masking test
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -x > $O/r02c20_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 $O/r02c20_pytest.log
V=$PWD/headtrackr_b200/variants
HT_LIB=$V/libht_split2.so timeout 600 python -m pytest tests/test_gpu_detect.py tests/test_gpu_quads.py tests/test_gpu_golden.py -m gpu -q --timeout 600 > $O/r02c20_pytest_split2.log 2>&1; tail -2 $O/r02c20_pytest_split2.log
run() { tag=$1; shift; env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline $BARGS > $O/r02c20_$tag.json 2> $O/r02c20_$tag.err; }
BARGS="--workload detect"
run det
run det_split1 HT_LIB=$V/libht_split1.so
run det_split2 HT_LIB=$V/libht_split2.so
run det_b
run det_split1_b HT_LIB=$V/libht_split1.so
run det_split2_b HT_LIB=$V/libht_split2.so
for f in $O/r02c20_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1].split("r02c20_")[1], round(d["value"]), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), d["kernel_ms_per_step"])
except Exception as e: print(sys.argv[1], "ERR", e); print(open(sys.argv[1].replace('.json','.err')).read()[-1500:])
PY
done
