#!/bin/bash
# session 19: MM-shaped suite with concurrent handles (after the shared-memory opt-in fix); source-level ncu of the partitioned multistage kernels
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
export PATH=/usr/local/cuda/bin:$PATH
B200_SUITE_THREADS=12 timeout 600 python tools/mm_suite.py > gpurun_out/s19_mm_suite_t12.json 2> gpurun_out/s19_mm_suite_t12.err
B200_SUITE_THREADS=4 timeout 600 python tools/mm_suite.py > gpurun_out/s19_mm_suite_t4.json 2> gpurun_out/s19_mm_suite_t4.err
B200_NO_GRAPH=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"msp_|msw_factor_chain|msw_solve" --launch-skip 24 -c 12 -o gpurun_out/s19_ms_part -f python tools/ms_probe.py 128 1 > gpurun_out/s19_ncu.log 2>&1
tail -2 gpurun_out/s19_mm_suite_t12.err
