#!/bin/bash
# session 32: window updates of the whole-GPU LDL^T split (next panel's columns first, the rest beside the next panel)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
export PATH=/usr/local/cuda/bin:$PATH
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/s32_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/s32_pytest.log
timeout 300 python bench.py --workload sparse_c3 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/s32_bench_c3.json 2> gpurun_out/s32_bench_c3.err
B200_WIDE_NO_WINDOW_SPLIT=1 timeout 300 python bench.py --workload sparse_c3 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/s32_bench_c3_nosplit.json 2> gpurun_out/s32_bench_c3_nosplit.err
B200_WIDE_GROUP=8 timeout 300 python bench.py --workload sparse_c3 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/s32_bench_c3_g8.json 2> gpurun_out/s32_bench_c3_g8.err
B200_TIMELINE=1 timeout 600 python tools/timeline.py --workload sparse_c3 --out gpurun_out/s32_timeline_c3.raw > gpurun_out/s32_timeline_c3.txt 2>&1
tail -n 3 gpurun_out/s32_pytest.log
