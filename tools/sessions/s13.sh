#!/bin/bash
# GPU session 13: pipelined graph replays (multistage), trapezoid upload of P, ozaki L2 evidence
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
export PATH=/usr/local/cuda/bin:$PATH
timeout 900 python -m pytest tests/test_gpu_multistage.py tests/test_gpu_full_size.py tests/test_gpu_dense.py tests/test_adapter_header.py -m gpu -q > gpurun_out/s13_pytest.log 2>&1
echo "rc=$?" >> gpurun_out/s13_pytest.log
timeout 300 python bench.py --workload multistage --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/s13_bench_ms.json 2> gpurun_out/s13_bench_ms.err
timeout 600 python bench.py --workload dense --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/s13_bench_dense.json 2> gpurun_out/s13_bench_dense.err
B200_DENSE_ASSEMBLE=ozaki timeout 600 ncu --set full --clock-control none --kernel-name regex:"oz_gemm_kernel" --launch-skip 3 --launch-count 1 -o gpurun_out/r02b_ncu_ozaki python bench.py --workload dense --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/s13_ncu_ozaki.log 2>&1
ncu -i gpurun_out/r02b_ncu_ozaki.ncu-rep --page raw --csv > gpurun_out/r02b_ncu_ozaki_raw.csv 2>/dev/null
rm -f gpurun_out/r02b_ncu_ozaki.ncu-rep
