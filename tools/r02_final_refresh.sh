#!/bin/bash
# round-2 final refresh after the k_resample change: tests, memcheck of smoke(), the bench lines BASELINE.md quotes, launch list
O=gpurun_out; mkdir -p $O; P=r02i
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > $O/${P}_pytest_gpu.txt 2>&1
echo "pytest rc=$?"; tail -3 $O/${P}_pytest_gpu.txt
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python __graft_entry__.py smoke > $O/${P}_memcheck_smoke.txt 2>&1; echo "memcheck rc=$?"; tail -4 $O/${P}_memcheck_smoke.txt
b() { tag=$1; shift; timeout 600 python bench.py "$@" > $O/${P}_bench_$tag.json 2> $O/${P}_bench_$tag.err; echo "$tag rc=$?"; }
b n1 --steps 10 --warmup 3
b reference --impl reference --steps 2 --warmup 1
b detect --steps 10 --warmup 3 --workload detect
b full_320 --steps 10 --warmup 3 --width 320 --height 240 --no-cpu-baseline
b d720_i3 --steps 10 --warmup 3 --workload detect720 --interval 3 --no-cpu-baseline
b d720_i5 --steps 10 --warmup 3 --workload detect720 --interval 5 --no-cpu-baseline
for f in $O/${P}_bench_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1].split("_bench_")[1], round(d["value"]), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), d.get("kernel_ms_per_step"), d.get("batch_parity"), (d.get("unpipelined") or {}).get("value"))
except Exception as e: print(sys.argv[1], "ERR", e); print(open(sys.argv[1].replace('.json','.err')).read()[-800:])
PY
done
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/${P}_launches_bench.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --pipeline 0 > $O/${P}_launches.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_resample -c 1 -o $O/${P}_resample -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --workload detect > $O/${P}_ncu_res.log 2>&1
ncu -i $O/${P}_resample.ncu-rep --page details --csv > $O/${P}_resample_details.csv 2>/dev/null
ls $O | grep ${P}_ | head -30
