#!/bin/bash
# session 18: weights of the DMMA contraction staged in smem, load-first IP vector kernels, concurrent handles in the MM-shaped suite
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
export PATH=/usr/local/cuda/bin:$PATH
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s18_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/s18_pytest.log
B200_TIMELINE=1 timeout 300 python tools/timeline.py --workload multistage --out gpurun_out/s18_timeline_ms.raw > gpurun_out/s18_timeline_ms.txt 2>&1
timeout 300 python bench.py --workload dense --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/s18_bench_dense.json 2> gpurun_out/s18_bench_dense.err
timeout 300 python bench.py --workload multistage --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/s18_bench_ms.json 2> gpurun_out/s18_bench_ms.err
B200_SUITE_THREADS=12 timeout 600 python tools/mm_suite.py > gpurun_out/s18_mm_suite_t12.json 2> gpurun_out/s18_mm_suite_t12.err
B200_SUITE_THREADS=4 timeout 600 python tools/mm_suite.py > gpurun_out/s18_mm_suite_t4.json 2> gpurun_out/s18_mm_suite_t4.err
tail -3 gpurun_out/s18_pytest.log
