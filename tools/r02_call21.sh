#!/bin/bash
# round-2 call 21: validation of the last change (getWeights skips the division for empty model bins)
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > $O/r02c21_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 $O/r02c21_pytest.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/r02c21_full.json 2> $O/r02c21_full.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --pipeline 0 > $O/r02c21_full_nopipe.json 2> $O/r02c21_full_nopipe.err
for f in $O/r02c21_*.json; do python - "$f" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1].split("r02c21_")[1], round(d["value"]), round(d["ms_per_step"],3), d["kernel_ms_per_step"], (d.get("unpipelined") or {}).get("value"))
PY
done
