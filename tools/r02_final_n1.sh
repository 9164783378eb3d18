#!/bin/bash
# round-2 final single-GPU measurement: everything profiles/ holds for the final build (prefix r02h_)
O=gpurun_out; mkdir -p $O; P=r02h
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > $O/${P}_pytest_gpu.txt 2>&1
echo "pytest rc=$?"; tail -3 $O/${P}_pytest_gpu.txt
python __graft_entry__.py smoke 2>&1 | tail -2
b() { tag=$1; shift; timeout 600 python bench.py "$@" > $O/${P}_bench_$tag.json 2> $O/${P}_bench_$tag.err; echo "$tag rc=$?"; }
b n1 --steps 10 --warmup 3
b reference --impl reference --steps 2 --warmup 1
b n1_pipelined --steps 10 --warmup 3 --pipeline 1 --no-cpu-baseline
b detect --steps 10 --warmup 3 --workload detect --no-cpu-baseline
b full_320 --steps 10 --warmup 3 --width 320 --height 240 --no-cpu-baseline
b full_1280 --steps 5 --warmup 3 --width 1280 --height 720 --batch 256 --no-cpu-baseline
b d720_i3 --steps 10 --warmup 3 --workload detect720 --interval 3 --no-cpu-baseline
b d720_i5 --steps 10 --warmup 3 --workload detect720 --interval 5 --no-cpu-baseline
b streams1 --steps 3 --warmup 3 --workload streams --streams 1 --no-cpu-baseline
b streams64 --steps 3 --warmup 3 --workload streams --streams 64 --no-cpu-baseline
for f in $O/${P}_bench_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1].split("_bench_")[1], round(d["value"]), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), d.get("kernel_ms_per_step"))
except Exception as e: print(sys.argv[1], "ERR", e); print(open(sys.argv[1].replace('.json','.err')).read()[-800:])
PY
done
# launch list of the bench command (cold-cache, serialised: shares, not absolutes)
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/${P}_launches_bench.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/${P}_launches.log 2>&1
# whole-step DRAM traffic without ncu's cache flush between kernels
timeout 400 ncu --cache-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
  --log-file $O/${P}_dram_step_1024.csv python tools/profile_run.py --frames 1024 --iters 2 > $O/${P}_dram.log 2>&1
# full captures, one launch each, on the bench's own batch
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_cascade -c 1 -o $O/${P}_cascade -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --workload detect > $O/${P}_ncu_casc.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_gray -c 1 -o $O/${P}_gray -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $O/${P}_ncu_gray.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_resample -c 1 -o $O/${P}_resample -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --workload detect > $O/${P}_ncu_res.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_track$ --launch-skip 5 -c 1 -o $O/${P}_track -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $O/${P}_ncu_track.log 2>&1
for k in cascade gray resample track; do
  ncu -i $O/${P}_$k.ncu-rep --page details --csv > $O/${P}_${k}_details.csv 2>/dev/null
  ncu -i $O/${P}_$k.ncu-rep --page raw --csv > $O/${P}_${k}_raw.csv 2>/dev/null
done
ncu -i $O/${P}_cascade.ncu-rep --page source --print-source cuda,sass --csv > $O/${P}_cascade_cs.csv 2>/dev/null
ls -la $O | grep ${P}_ | awk '{print $5, $9}' | tail -40
