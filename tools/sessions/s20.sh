#!/bin/bash
# session 20: fused multistage solve (one launch per backend solve)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
export PATH=/usr/local/cuda/bin:$PATH
timeout 900 python -m pytest tests/test_gpu_multistage.py tests/test_gpu_full_size.py -m gpu -x -q > gpurun_out/s20_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/s20_pytest.log
timeout 300 python bench.py --workload multistage --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/s20_bench_ms.json 2> gpurun_out/s20_bench_ms.err
B200_MS_FUSED_SOLVE=0 timeout 300 python bench.py --workload multistage --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/s20_bench_ms_unfused.json 2> gpurun_out/s20_bench_ms_unfused.err
B200_TIMELINE=1 timeout 300 python tools/timeline.py --workload multistage --out gpurun_out/s20_timeline_ms.raw > gpurun_out/s20_timeline_ms.txt 2>&1
tail -3 gpurun_out/s20_pytest.log
