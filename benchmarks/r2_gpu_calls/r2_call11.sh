#!/bin/bash
mkdir -p gpurun_out
for wl in 28 29; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 --no-e2e --window-log2 $wl > gpurun_out/c11_bench_n2_wl$wl.json 2> gpurun_out/c11_bench_n2_wl$wl.err
python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/c11_bench_n2_wl$wl.json") if l.startswith("{")][-1])
    print("N=2 wl=$wl headline", d["value"]/1e9, d["ms_per_step"], "parity", d["parity"]["parity"], {k:(round(v["total_ms"],1)) for k,v in d["roofline"]["kernels"].items()})
except Exception as e: print("wl=$wl failed", e, open("gpurun_out/c11_bench_n2_wl$wl.err").read()[-1500:])
PY
done
for o in "bloom_part_tile=512" "bloom_window_log2_bits=28" ; do
timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-micro --no-parts --no-parity --opt $o > gpurun_out/c11_bench_n1_$o.json 2> gpurun_out/c11_bench_n1_$o.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/c11_bench_n1_$o.json")); print("N=1 $o", d["value"]/1e9, d["ms_per_step"], {k:(round(v["total_ms"],1)) for k,v in d["roofline"]["kernels"].items()})
except Exception as e: print("$o failed", e)
PY
done
