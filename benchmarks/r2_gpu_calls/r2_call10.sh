#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/c10_pytest.log 2>&1
tail -6 gpurun_out/c10_pytest.log
for o in "bloom_part_tile=0" "bloom_part_tile=256"; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 --no-e2e --opt $o > gpurun_out/c10_bench_n2_$o.json 2> gpurun_out/c10_bench_n2_$o.err
python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/c10_bench_n2_$o.json") if l.startswith("{")][-1])
    print("$o headline", d["value"]/1e9, d["ms_per_step"], "parity", d["parity"]["parity"], {k:(round(v["total_ms"],1)) for k,v in d["roofline"]["kernels"].items()})
except Exception as e: print("$o failed", e, open("gpurun_out/c10_bench_n2_$o.err").read()[-1500:])
PY
done
