#!/bin/bash
# round-2, 8 GPUs of one box: the BASELINE configs that name 8 GPUs (config 4: 1280x720, 4096 frames sharded, interval 3 and 5;
# config 5: one 640x480 stream per GPU, and 8 per GPU).  The default workload's 1-8 curve is the driver's own SCALE run.
O=gpurun_out; mkdir -p $O; P=r02h
tr() { tag=$1; port=$2; shift 2; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $port bench.py --gpus 8 "$@" > $O/${P}_bench_n8_$tag.json 2> $O/${P}_bench_n8_$tag.err; echo "$tag rc=$?"; tail -c 300 $O/${P}_bench_n8_$tag.json | head -c 300; echo; }
tr d720_i3 29611 --steps 5 --warmup 3 --workload detect720 --interval 3
tr d720_i5 29612 --steps 5 --warmup 3 --workload detect720 --interval 5
tr streams1 29613 --steps 3 --warmup 3 --workload streams --streams 1
tr streams8 29614 --steps 3 --warmup 3 --workload streams --streams 8
tr full 29615 --steps 5 --warmup 3
for f in $O/${P}_bench_n8_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1].split("_bench_")[1], round(d["value"]), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), d.get("shard_check"), [r["ms_per_step"] for r in d["per_rank"]])
except Exception as e: print(sys.argv[1], "ERR", e); print(open(sys.argv[1].replace('.json','.err')).read()[-800:])
PY
done
