#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 benchmarks/sharded_check.py --keys 1000000 > gpurun_out/c8_sharded_check_n2.txt 2>&1
tail -8 gpurun_out/c8_sharded_check_n2.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 --config5 > gpurun_out/c8_bench_n2.json 2> gpurun_out/c8_bench_n2.err
tail -5 gpurun_out/c8_bench_n2.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/c8_bench_n2.json") if l.startswith("{")][-1])
print("headline", d["value"]/1e9, d["ms_per_step"], "parity", d["parity"])
print("e2e", d["e2e"])
print("kernels", {k:(round(v["total_ms"],1)) for k,v in d["roofline"]["kernels"].items()})
print("config5", d.get("config5"))
PY
