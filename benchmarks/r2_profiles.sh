#!/bin/bash
# round-2 ncu evidence: launch list of the default bench + one `--set full` capture per hot kernel (1 GPU)
mkdir -p gpurun_out
NCU="ncu --clock-control none"
timeout 900 $NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-micro --cuckoo-log2 26 > gpurun_out/r2_launches_bench.log 2>&1
prof() {  # part, kernel regex, launch skip, count
  timeout 900 $NCU --set full --import-source on -k "regex:$2" -s $3 -c $4 -o gpurun_out/r2_prof_$1 python benchmarks/profile_parts.py $1 > gpurun_out/r2_prof_$1.log 2>&1
  ncu -i gpurun_out/r2_prof_$1.ncu-rep --page raw --csv > gpurun_out/r2_prof_$1.raw.csv 2>/dev/null
  ncu -i gpurun_out/r2_prof_$1.ncu-rep --page source --csv > gpurun_out/r2_prof_$1.source.csv 2>/dev/null
  rm -f gpurun_out/r2_prof_$1.ncu-rep   # (gpurun brings back at most 64 MiB: keep the exported pages, not the reports)
}
prof bloom_insert "bloom_part4|bloom_apply2" 2 2
prof bloom_insert_w286 "bloom_part4" 2 1
prof bloom_check "bloom_check_fixed16" 0 1
prof bloom_query "bloom_probe2|bloom_part4" 4 2
prof cms "cms_add_fixed16|cms_check_fixed16" 0 2
prof cuckoo "cuckoo_claim_fixed16|cuckoo_insert_kernel|cuckoo_check_fixed16" 4 3
prof cbloom "cbloom_add_fixed16|cbloom_check_fixed16" 0 2
ls -la gpurun_out | grep r2_
ls -la gpurun_out | grep r2_
