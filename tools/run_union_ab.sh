#!/bin/bash
# A/B runs of library variants on the union workload (gpurun helper; writes gpurun_out/r2_ab_*). Usage: run_union_ab.sh tag lib...
cd "$(dirname "$0")/.."
B="python bench.py --workload pubmed_posplus_union --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --configs none"
tag=$1; shift
for lib in default "$@"; do
  name=$(basename $lib .so)
  if [ "$lib" = default ]; then unset S3GRL_LIB; else export S3GRL_LIB=$PWD/$lib; fi
  S3GRL_BENCH_DEBUG=1 timeout 120 $B > gpurun_out/r2_${tag}_$name.json 2> gpurun_out/r2_${tag}_$name.err
  echo "$name rc=$? $(python tools/bench_summary.py gpurun_out/r2_${tag}_$name.json 2>&1 | head -1) $(grep ccn_chain gpurun_out/r2_${tag}_$name.err | head -1)"
done
