#!/bin/bash
# A/B runs of the union chain's tuning knobs + one ncu capture (gpurun helper; writes gpurun_out/r2_u2_*)
cd "$(dirname "$0")/.."
B="python bench.py --workload pubmed_posplus_union --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --configs none"
for v in "POLICY=1" "POLICY=2" "SLABS=2" "SLABS=4"; do
  env S3GRL_CHAIN_$v S3GRL_BENCH_DEBUG=1 timeout 120 $B > gpurun_out/r2_u2_$v.json 2> gpurun_out/r2_u2_$v.err
  echo "$v rc=$? $(python tools/bench_summary.py gpurun_out/r2_u2_$v.json 2>&1 | head -1) $(grep ccn_chain gpurun_out/r2_u2_$v.err | head -1)"
done
S="python bench.py --workload pubmed_posplus_union --links 8000 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --configs none"
timeout 120 $S > gpurun_out/r2_u2_small.json 2> gpurun_out/r2_u2_small.err && \
timeout 250 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:chain -c 16 --csv --log-file gpurun_out/r2_u2_launches.csv $S > gpurun_out/r2_u2_ncu1.log 2>&1
echo "launch list rc=$?"; grep -c chain gpurun_out/r2_u2_launches.csv
timeout 280 ncu --set full --clock-control none --import-source on -k regex:chain_kernel -c 3 -f -o gpurun_out/r2_u2_chain $S > gpurun_out/r2_u2_ncu2.log 2>&1
echo "full capture rc=$?"; ls -la gpurun_out/*.ncu-rep
