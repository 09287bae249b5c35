#!/bin/bash
# session 31: device timeline of config 3 (whole-GPU sparse schedule); memcheck of the new kernels
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
export PATH=/usr/local/cuda/bin:$PATH
B200_TIMELINE=1 timeout 600 python tools/timeline.py --workload sparse_c3 --out gpurun_out/s31_timeline_c3.raw > gpurun_out/s31_timeline_c3.txt 2>&1
timeout 900 compute-sanitizer --tool memcheck --log-file gpurun_out/s31_memcheck_dense_bulk.log python -m pytest tests/test_gpu_dense.py -m gpu -x -q -k "factor_solve_eval_parity and (dmma-dims6 or dmma-dims7)" > gpurun_out/s31_memcheck_dense.out 2>&1
timeout 900 compute-sanitizer --tool memcheck --log-file gpurun_out/s31_memcheck_ms_fused.log python -m pytest tests/test_gpu_multistage.py -m gpu -x -q -k "partition" > gpurun_out/s31_memcheck_ms.out 2>&1
tail -3 gpurun_out/s31_memcheck_dense.out gpurun_out/s31_memcheck_ms.out; tail -3 gpurun_out/s31_memcheck_dense_bulk.log gpurun_out/s31_memcheck_ms_fused.log
