#!/bin/bash
# ncu evidence for one workload (run under gpurun, 1 GPU): launch list + full capture of the hot kernels.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
W=${1:-c2}
python tools/profile_step.py $W 3 > gpurun_out/plain_$W.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$W.csv \
    python tools/profile_step.py $W 1 > gpurun_out/ncu_launch_$W.log 2>&1
python tools/profile_step.py $W 1 >> gpurun_out/plain_$W.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_onesweep_pass -s 1 -c 1 -o gpurun_out/prof_pass_$W \
    python tools/profile_step.py $W 1 > gpurun_out/ncu_pass_$W.log 2>&1
python tools/profile_step.py $W 1 >> gpurun_out/plain_$W.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_rerank|k_emit_bwt|k_init_keys|k_build_keys|k_symbol" -c 5 -o gpurun_out/prof_other_$W \
    python tools/profile_step.py $W 1 > gpurun_out/ncu_other_$W.log 2>&1
cat gpurun_out/plain_$W.log | tail -n 3
ls -la gpurun_out/
