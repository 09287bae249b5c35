#!/bin/bash
# session 17: source-level ncu capture of the dense assembly kernel and the Cholesky updates (where does the DMMA pipe idle?)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
export PATH=/usr/local/cuda/bin:$PATH
timeout 800 ncu --set full --clock-control none --import-source on -k regex:"gemm_nt_t64_kernel" --launch-skip 0 -c 9 -o gpurun_out/s17_t64 -f python tools/dense_probe.py 256 > gpurun_out/s17_ncu.log 2>&1
ls -la gpurun_out/ | tail -5
