#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
OUT=gpurun_out
python tools/profile_step.py c3 1 > $OUT/plain_c3.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_rerank|k_build_keys|k_partition|k_scatter_ranks" -c 6 -o $OUT/prof_rerank_c3 \
   python tools/profile_step.py c3 1 > $OUT/ncu_rerank_c3.log 2>&1
ls -la $OUT | tail -3
