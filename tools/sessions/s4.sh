#!/bin/bash
# GPU session 4: launch lists of the multistage kernels (partition vs sequential); multi-workload bench after the capture-mutex fix
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
export PATH=/usr/local/cuda/bin:$PATH
for v in part nopart; do
  if [ $v = nopart ]; then export B200_MS_NO_PARTITION=1; else export B200_MS_NO_PARTITION=0; fi
  B200_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name regex:"ms" -c 260 --csv --log-file gpurun_out/s4_launches_ms_$v.csv \
     python bench.py --workload multistage --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/s4_ncu_ms_$v.log 2>&1
done
unset B200_MS_NO_PARTITION
( time timeout 1500 python bench.py --steps 5 --warmup 3 ) > gpurun_out/s4_bench_all.json 2> gpurun_out/s4_bench_all.err
timeout 300 python -m pytest tests/test_adapter_header.py -m gpu -q > gpurun_out/s4_pytest_adapter.log 2>&1
