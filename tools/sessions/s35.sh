#!/bin/bash
# session 35: dense e2e with the library's H2D turnstile vs bench-side turns
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python bench.py --workload dense --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/s35_bench_dense.json 2> gpurun_out/s35_bench_dense.err
B200_E2E_TURNS=1 timeout 300 python bench.py --workload dense --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/s35_bench_dense_turns.json 2> gpurun_out/s35_bench_dense_turns.err
B200_E2E_CHUNKS=6 timeout 300 python bench.py --workload dense --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/s35_bench_dense_c6.json 2> gpurun_out/s35_bench_dense_c6.err
grep "e2e rep" gpurun_out/s35_*.err
