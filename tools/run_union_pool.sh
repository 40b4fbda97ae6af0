#!/bin/bash
# Union chain: shared-memory placement vs the pooled route (S3GRL_CHAIN_POOL_CW[:SPLIT[:POOL_X[:SMEM_X]]]) — gpurun helper,
# writes gpurun_out/r2_pool_*.  Usage: run_union_pool.sh "cw[:split[:x[:sx]]] ..."
cd "$(dirname "$0")/.."
timeout 120 python -m pytest tests/test_gpu_parity.py -x -q -k "${S3_POOL_TESTS:-pooled or union_chain_matches}" > gpurun_out/r2_pool_pytest.log 2>&1
echo "pytest rc=$? $(tail -1 gpurun_out/r2_pool_pytest.log)"
B="python bench.py --workload pubmed_posplus_union --steps 2 --warmup 2 --no-e2e --no-cpu-baseline --configs none"
for v in ${1:-0 8 16 32 16:1 32:1}; do
  IFS=: read cw split x sx <<< "$v"; split=${split:-0}; x=${x:-0}; sx=${sx:-0}
  S3GRL_CHAIN_POOL_CW=$cw S3GRL_CHAIN_POOL_SPLIT=$split S3GRL_CHAIN_POOL_X=$x S3GRL_CHAIN_SMEM_X=$sx S3GRL_BENCH_DEBUG=1 timeout 100 $B > gpurun_out/r2_pool_cw${cw}_s${split}_x${x}_sx$sx.json 2> gpurun_out/r2_pool_cw${cw}_s${split}_x${x}_sx$sx.err
  echo "cw=$cw split=$split x=$x sx=$sx rc=$? $(python tools/bench_summary.py gpurun_out/r2_pool_cw${cw}_s${split}_x${x}_sx$sx.json 2>&1 | head -1) $(grep ccn_chain gpurun_out/r2_pool_cw${cw}_s${split}_x${x}_sx$sx.err | head -1 | cut -c1-70)"
done
