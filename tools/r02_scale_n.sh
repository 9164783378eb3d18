#!/bin/bash
# default workload at N GPUs (the driver's SCALE run does the same at round end): bash tools/r02_scale_n.sh N
N=$1; O=gpurun_out; mkdir -p $O
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2971$N bench.py --gpus $N --steps 10 --warmup 3 > $O/r02i_bench_n$N.json 2> $O/r02i_bench_n$N.err; echo "rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2972$N bench.py --gpus $N --steps 10 --warmup 3 --width 320 --height 240 > $O/r02i_bench_n${N}_320.json 2> $O/r02i_bench_n${N}_320.err; echo "rc=$?"
for f in $O/r02i_bench_n${N}.json $O/r02i_bench_n${N}_320.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1].split("_bench_")[1], round(d["value"]), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), d.get("shard_check"), [r["ms_per_step"] for r in d["per_rank"]], (d.get("unpipelined") or {}).get("value"))
except Exception as e: print(sys.argv[1], "ERR", e); print(open(sys.argv[1].replace('.json','.err')).read()[-800:])
PY
done
