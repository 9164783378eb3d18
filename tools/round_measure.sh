#!/bin/bash
# One GPU-box call that produces everything profiles/ holds for a round:
#   gpurun --timeout 900 -- 'bash tools/round_measure.sh r01'
# (numbers printed under ncu are never bench values; the bench lines come from the plain runs)
R=${1:-rXX}
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee $O/pytest_gpu_$R.txt
python bench.py --steps 5 --warmup 3 > $O/bench_${R}_n1.json 2> $O/bench_${R}_n1.err
python bench.py --steps 5 --warmup 3 --workload detect --no-cpu-baseline > $O/bench_${R}_detect.json 2>> $O/bench_${R}_n1.err
python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_${R}_reference.json 2>> $O/bench_${R}_n1.err
tail -c 600 $O/bench_${R}_n1.json
# launch list of the bench command (cold-cache, serialised: shares, not absolutes)
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/launches_${R}.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/launches_${R}.log 2>&1
# full captures of the two kernels that changed this round (148 frames, strict tracking)
HT_TRACK_MEMO=0 timeout 300 ncu --set full --clock-control none --import-source on -k regex:^k_track$ -s 1 -c 1 -f \
  -o $O/prof_track_${R}_final python tools/profile_run.py --frames 148 --iters 2 > $O/ncu_track_$R.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_resample -s 7 -c 1 -f \
  -o $O/prof_resample_${R}_final python tools/profile_run.py --frames 148 --iters 2 > $O/ncu_resample_$R.log 2>&1
ls -la $O | tail -12
