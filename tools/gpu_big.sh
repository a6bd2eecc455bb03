#!/bin/bash
# C4 (2 GiB block) parity + timing, then the launch list + full ncu capture of the current kernels on C2.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
OUT=gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 1400 -p no:cacheprovider -k "2147483648" -s > $OUT/pytest_c4.log 2>&1; echo "pytest c4 rc=$?"; tail -n 4 $OUT/pytest_c4.log
timeout 600 python tools/profile_step.py c4 2 > $OUT/step_c4.log 2>&1; echo "c4 rc=$?"; tail -n 1 $OUT/step_c4.log | cut -c1-1200
python tools/profile_step.py c2 2 > $OUT/plain_c2.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $OUT/launches_c2.csv \
    python tools/profile_step.py c2 1 > $OUT/ncu_launch_c2.log 2>&1
python tools/profile_step.py c2 1 >> $OUT/plain_c2.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_onesweep_pass|k_rerank|k_emit_bwt_window|k_init_keys" -c 5 -o $OUT/prof_c2_v4 \
    python tools/profile_step.py c2 1 > $OUT/ncu_full_c2.log 2>&1
ls -la $OUT | grep -E "prof_c2_v4|launches_c2"
