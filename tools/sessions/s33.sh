#!/bin/bash
# session 33: where the multistage e2e setup time goes
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
B200_TIMING=1 timeout 300 python tools/ms_probe.py 128 3 > gpurun_out/s33_ms_probe.txt 2>&1
tail -n 40 gpurun_out/s33_ms_probe.txt
