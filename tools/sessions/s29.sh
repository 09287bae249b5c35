#!/bin/bash
# session 29: iteration parity of the dense path at full size, 32 instances, with and without the Cholesky look-ahead
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python tools/dense_iter_parity.py 32 > gpurun_out/s29_dense_iter_parity.txt 2>&1
B200_CHOL_LOOKAHEAD=0 timeout 900 python tools/dense_iter_parity.py 32 > gpurun_out/s29_dense_iter_parity_nola.txt 2>&1
tail -2 gpurun_out/s29_dense_iter_parity.txt gpurun_out/s29_dense_iter_parity_nola.txt
