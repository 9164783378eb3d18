#!/bin/bash
# round-2 call 4: v3 cascade (bit masks, 256-thread CTAs, TH=8) as the default build; track experiments; profiles
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > $O/r02c4_pytest.log 2>&1
echo "pytest rc=$?"; tail -6 $O/r02c4_pytest.log
run() { tag=$1; shift; env "$@" timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline $BARGS > $O/r02c4_$tag.json 2> $O/r02c4_$tag.err; }
BARGS=""
run full
run full_mid8 HT_TRACK_MID=8
run full_mid16 HT_TRACK_MID=16
run full_heavy32_mid8 HT_TRACK_HEAVY=32 HT_TRACK_MID=8
run full_nt512 HT_TRACK_NT=512 HT_TRACK_CLUSTER=1
run full_c4 HT_TRACK_CLUSTER=4
run full_minb3 HT_LIB=variants/libht_minb3.so
run full_minb3_ov2 HT_LIB=variants/libht_minb3.so HT_OVERLAP=2
run full_minb3_ov4 HT_LIB=variants/libht_minb3.so HT_OVERLAP=4
run full_ov4 HT_OVERLAP=4
run full_th12 HT_LIB=variants/libht_th12.so
run full_th12_minb2_ov4 HT_LIB=variants/libht_th12_minb2.so HT_OVERLAP=4
run full_tma HT_TMA=1
run full_mb256 HT_WAVE_MB=256
BARGS="--workload detect"
run det
BARGS="--workload detect720 --interval 3"
run d720_i3
BARGS="--workload detect720 --interval 5"
run d720_i5
BARGS="--width 320 --height 240"
run full_320
BARGS="--workload streams --streams 1 --stream-frames 120"
run streams1
BARGS="--workload streams --streams 64 --stream-frames 120"
run streams64
for f in $O/r02c4_full*.json $O/r02c4_det.json $O/r02c4_d720*.json $O/r02c4_streams*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1].split("r02c4_")[1], round(d["value"]), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), d["kernel_ms_per_step"])
except Exception as e: print(sys.argv[1], "ERR", e)
PY
done
timeout 300 ncu --cache-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
  --log-file $O/r02c4_dram_step.csv python tools/profile_run.py --frames 1024 --iters 2 > $O/r02c4_dram.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_cascade -s 1 -c 1 -f \
  -o $O/r02c4_cascade python tools/profile_run.py --frames 1024 --iters 2 > $O/r02c4_ncu_cascade.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k 'regex:^k_track$' -s 2 -c 1 -f \
  -o $O/r02c4_track python tools/profile_run.py --frames 1024 --iters 2 > $O/r02c4_ncu_track.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_gray -s 1 -c 1 -f \
  -o $O/r02c4_gray python tools/profile_run.py --frames 1024 --iters 2 > $O/r02c4_ncu_gray.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_resample -s 8 -c 1 -f \
  -o $O/r02c4_resample python tools/profile_run.py --frames 1024 --iters 2 > $O/r02c4_ncu_resample.log 2>&1
ls $O | grep -c r02c4
