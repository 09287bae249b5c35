#!/bin/bash
# session 36: partition count of the multistage backend with the fused solve; handles in flight in the suite
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for k in 4 6 7; do
  B200_MS_SEGMENTS=$k timeout 300 python bench.py --workload multistage --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/s36_bench_ms_k$k.json 2> gpurun_out/s36_bench_ms_k$k.err
done
B200_SUITE_THREADS=24 timeout 600 python tools/mm_suite.py > gpurun_out/s36_mm_suite_t24.json 2> gpurun_out/s36_mm_suite_t24.err
echo done
