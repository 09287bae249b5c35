#!/bin/bash
# session 34: IP-solver vectors from slabs (setup time); full GPU tests; multistage + dense e2e
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/s34_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/s34_pytest.log
timeout 300 python bench.py --workload multistage --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/s34_bench_ms.json 2> gpurun_out/s34_bench_ms.err
timeout 300 python bench.py --workload dense --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/s34_bench_dense.json 2> gpurun_out/s34_bench_dense.err
tail -n 3 gpurun_out/s34_pytest.log
