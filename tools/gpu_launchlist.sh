#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
OUT=gpurun_out
for w in ${WL:-c3}; do
python tools/profile_step.py $w 1 > $OUT/plain_$w.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $OUT/launches_$w.csv \
    python tools/profile_step.py $w 1 > $OUT/ncu_launch_$w.log 2>&1
done
