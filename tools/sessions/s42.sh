#!/bin/bash
# session 42: group size of the two-level blocked LDL^T with the bulk tile kernel
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for g in 2 3 6; do
  B200_WIDE_GROUP=$g timeout 200 python bench.py --workload sparse_c3 --steps 2 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/s42_bench_c3_g$g.json 2> gpurun_out/s42_bench_c3_g$g.err
done
echo done
