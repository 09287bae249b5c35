#!/bin/bash
# session 24: bulk DMMA tile kernel with a dedicated producer warp
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
export PATH=/usr/local/cuda/bin:$PATH
timeout 900 python -m pytest tests/test_gpu_dense.py -m gpu -x -q > gpurun_out/s24_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/s24_pytest.log
timeout 300 python bench.py --workload dense --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/s24_bench_dense.json 2> gpurun_out/s24_bench_dense.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gemm_nt_t64_bulk" --launch-skip 0 -c 9 -o gpurun_out/s24_t64_bulk -f python tools/dense_probe.py 256 > gpurun_out/s24_ncu.log 2>&1
tail -3 gpurun_out/s24_pytest.log
