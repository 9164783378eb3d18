#!/bin/bash
# round-2 call 6: light tier for k_track, packed histogram counters in k_gray, profiles of the 2-CTA track launch
O=gpurun_out; mkdir -p $O
timeout 1200 python -m pytest tests/test_gpu_detect.py tests/test_gpu_quads.py tests/test_gpu_track.py tests/test_gpu_stream.py tests/test_gpu_golden.py -q --timeout 900 > $O/r02c6_pytest.log 2>&1
echo "pytest rc=$?"; tail -5 $O/r02c6_pytest.log
run() { tag=$1; shift; env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline $BARGS > $O/r02c6_$tag.json 2> $O/r02c6_$tag.err; }
BARGS=""
run full
run full_light2 HT_TRACK_LIGHT=2
run full_light3 HT_TRACK_LIGHT=3
run full_light4 HT_TRACK_LIGHT=4
run full_light2_128 HT_TRACK_LIGHT=2,128
run full_light2_mid8 HT_TRACK_LIGHT=2 HT_TRACK_MID=8
run full_light2_h32 HT_TRACK_LIGHT=2 HT_TRACK_HEAVY=32
run full_light2_nt128 HT_TRACK_LIGHT=2 HT_TRACK_NT=128
run full_light15 HT_TRACK_LIGHT=1
for f in $O/r02c6_full*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1].split("r02c6_")[1], round(d["value"]), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), d["kernel_ms_per_step"], d["clocks"])
except Exception as e: print(sys.argv[1], "ERR", e)
PY
done
HT_TRACK_HEAVY=0 HT_TRACK_MID=0 timeout 400 ncu --set full --clock-control none --import-source on -k 'regex:^k_track$' -s 1 -c 1 -f \
  -o $O/r02c6_track2 python tools/profile_run.py --frames 1024 --iters 2 > $O/r02c6_ncu_track.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_gray -s 1 -c 1 -f \
  -o $O/r02c6_gray python tools/profile_run.py --frames 1024 --iters 2 > $O/r02c6_ncu_gray.log 2>&1
ls $O | grep -c r02c6
