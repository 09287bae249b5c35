#!/bin/bash
# session 43: final bench record with the corrected roofline kernel name; ncu launch lists of the final dense and multistage solves
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
export PATH=/usr/local/cuda/bin:$PATH
(time timeout 900 python bench.py --steps 10 --warmup 3) > gpurun_out/s43_bench_all.json 2> gpurun_out/s43_bench_all.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/s43_launches_dense_c2.csv python tools/dense_probe.py 256 > gpurun_out/s43_ncu_dense.log 2>&1
B200_NO_GRAPH=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/s43_launches_multistage_c4.csv python tools/ms_probe.py 128 1 > gpurun_out/s43_ncu_ms.log 2>&1
ls -la gpurun_out/s43_*
