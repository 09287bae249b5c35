#!/bin/bash
# session 39: GPU tests after the symbolic-phase changes (golden permutation / factor checks); where config 3's setup time goes
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/s39_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/s39_pytest.log
B200_DEBUG_SYMBOLIC=1 timeout 600 python tools/sparse_big_probe.py 10000 0.01 0 1 > gpurun_out/s39_c3_probe.txt 2>&1
timeout 300 python bench.py --workload sparse_c3 --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/s39_bench_c3.json 2> gpurun_out/s39_bench_c3.err
tail -n 3 gpurun_out/s39_pytest.log; grep -v "big front" gpurun_out/s39_c3_probe.txt | tail -n 25
