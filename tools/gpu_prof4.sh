#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
OUT=gpurun_out
python tools/profile_step.py c2 2 > $OUT/plain_c2.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $OUT/launches_c2.csv \
    python tools/profile_step.py c2 1 > $OUT/ncu_launch_c2.log 2>&1
python tools/profile_step.py c2 1 >> $OUT/plain_c2.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_rerank|k_init_keys|k_onesweep" -c 3 -o $OUT/prof_c2_final \
    python tools/profile_step.py c2 1 > $OUT/ncu_full_c2.log 2>&1
ls -la $OUT | grep -E "prof_c2_final|launches_c2"
