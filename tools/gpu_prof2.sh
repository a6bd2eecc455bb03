#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
OUT=gpurun_out
for v in 0 5; do
python tools/sort_bench.py 26 $v > $OUT/plain_sort_$v.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_onesweep_pass -s 9 -c 1 -o $OUT/prof_sort_v$v \
   python tools/sort_bench.py 26 $v > $OUT/ncu_sort_$v.log 2>&1
done
for w in c5 c3; do
python tools/profile_step.py $w 1 > $OUT/plain_$w.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $OUT/launches_$w.csv \
    python tools/profile_step.py $w 1 > $OUT/ncu_launch_$w.log 2>&1
done
ls -la $OUT
