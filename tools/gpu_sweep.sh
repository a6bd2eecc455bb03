#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
OUT=gpurun_out
echo "== counter" > $OUT/sweep.log
timeout 600 python tools/sort_bench.py 27 ${SWEEP:-5} >> $OUT/sweep.log 2>&1
echo "== blockidx" >> $OUT/sweep.log
DARK_BWT_TILE_BY_BLOCKIDX=1 timeout 600 python tools/sort_bench.py 27 ${SWEEP:-5} >> $OUT/sweep.log 2>&1
grep -E "==|variant" $OUT/sweep.log | cut -c1-120
