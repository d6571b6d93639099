#!/bin/bash
# ncu --set full on the cuckoo insert kernels after the r2 rewrite (launch 3-4 = the batch that takes the table from 50 % to 93 %)
mkdir -p gpurun_out
NCU="ncu --clock-control none"
timeout 900 $NCU --set full --import-source on -k "regex:cuckoo_claim_fixed16|cuckoo_insert_kernel" -s 2 -c 2 -o gpurun_out/r2b_prof_cuckoo python benchmarks/profile_parts.py cuckoo > gpurun_out/r2b_prof_cuckoo.log 2>&1
ncu -i gpurun_out/r2b_prof_cuckoo.ncu-rep --page raw --csv > gpurun_out/r2b_prof_cuckoo.raw.csv 2>/dev/null
ncu -i gpurun_out/r2b_prof_cuckoo.ncu-rep --page source --csv > gpurun_out/r2b_prof_cuckoo.source.csv 2>/dev/null
rm -f gpurun_out/r2b_prof_cuckoo.ncu-rep
tail -3 gpurun_out/r2b_prof_cuckoo.log
ls -la gpurun_out/r2b_prof_cuckoo*
