#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
OUT=gpurun_out
python tools/profile_step.py ${1:-c2} 1 > $OUT/plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_rerank" -c 2 -o $OUT/prof_rerank_final \
    python tools/profile_step.py ${1:-c2} 1 > $OUT/ncu_rerank.log 2>&1
ls -la $OUT | grep prof_rerank_final
