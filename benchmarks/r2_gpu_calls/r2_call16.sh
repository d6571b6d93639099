#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/c16_pytest.log 2>&1
tail -4 gpurun_out/c16_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 benchmarks/sharded_check.py --keys 1000000 > gpurun_out/c16_sharded_check_n2.txt 2>&1
tail -2 gpurun_out/c16_sharded_check_n2.txt
for ck in 134217728 67108864; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 --chunk-keys $ck > gpurun_out/c16_bench_n2_$ck.json 2> gpurun_out/c16_bench_n2_$ck.err
python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/c16_bench_n2_$ck.json") if l.startswith("{")][-1])
    print("N=2 chunk=$ck headline", d["value"]/1e9, d["ms_per_step"], "parity", d["parity"]["parity"], {k:(round(v["total_ms"],1)) for k,v in d["roofline"]["kernels"].items()}, "e2e", d["e2e"]["value"])
except Exception as e: print("$ck failed", e, open("gpurun_out/c16_bench_n2_$ck.err").read()[-1500:])
PY
done
