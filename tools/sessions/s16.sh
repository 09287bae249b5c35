#!/bin/bash
# session 16: live device timelines of one captured IP iteration (multistage C4, sparse batch 148, dense C2)
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
for w in multistage sparse dense; do
  B200_TIMELINE=1 timeout 600 python tools/timeline.py --workload $w --out gpurun_out/s16_timeline_$w.raw > gpurun_out/s16_timeline_$w.txt 2>&1
done
B200_TIMELINE=1 B200_MS_NO_PARTITION=1 timeout 600 python tools/timeline.py --workload multistage --out gpurun_out/s16_timeline_ms_nopart.raw > gpurun_out/s16_timeline_ms_nopart.txt 2>&1
tail -5 gpurun_out/s16_timeline_multistage.txt
