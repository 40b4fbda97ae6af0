#!/bin/bash
# ncu evidence for the union chain (gpurun helper; writes gpurun_out/r2_s2_*): launch list of one small step + full capture
cd "$(dirname "$0")/.."
S="python bench.py --workload pubmed_posplus_union --links 8000 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --configs none"
timeout 120 $S > gpurun_out/r2_s2_union_small.json 2> gpurun_out/r2_s2_union_small.err && \
timeout 250 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 40 --csv --log-file gpurun_out/r2_s2_union_launches.csv $S > gpurun_out/r2_s2_ncu1.log 2>&1
echo "launch list rc=$?"
timeout 280 ncu --set full --clock-control none --import-source on -k regex:chain_kernel -c 3 -f -o gpurun_out/r2_s2_chain $S > gpurun_out/r2_s2_ncu2.log 2>&1
echo "full capture rc=$?"; ls -la gpurun_out/r2_s2_chain.ncu-rep
