#!/bin/bash
# session 44: ncu launch list of one dense config-2 solve, final kernels (library kernels only)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
export PATH=/usr/local/cuda/bin:$PATH
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:"b200::" -c 900 --csv --log-file gpurun_out/s44_launches_dense_c2.csv python tools/dense_probe.py 256 > gpurun_out/s44_ncu_dense.log 2>&1
tail -n 2 gpurun_out/s44_ncu_dense.log; wc -l gpurun_out/s44_launches_dense_c2.csv
