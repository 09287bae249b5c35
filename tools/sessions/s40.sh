#!/bin/bash
# session 40: validation of the final round-2 build: GPU tests, smoke, default bench (both arms)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/s40_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/s40_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/s40_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/s40_smoke.log
(time timeout 900 python bench.py --steps 10 --warmup 3) > gpurun_out/s40_bench_all.json 2> gpurun_out/s40_bench_all.err
(time timeout 600 python bench.py --impl reference --steps 3 --warmup 1) > gpurun_out/s40_bench_reference.json 2> gpurun_out/s40_bench_reference.err
tail -n 3 gpurun_out/s40_pytest.log; tail -n 2 gpurun_out/s40_smoke.log
