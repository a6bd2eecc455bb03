#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
: > gpurun_out/pass_trace.log
for ws in ${WS:-0 1}; do for v in ${VARS:-1}; do echo "== warp_specialized=$ws" | tee -a gpurun_out/pass_trace.log; DARK_BWT_WARP_SPECIALIZED=$ws DARK_BWT_EARLY_LOOKBACK=$ws timeout 300 python tools/pass_trace.py 27 $v 2>&1 | tee -a gpurun_out/pass_trace.log; done; done
for ws in ${WS:-0 1}; do echo "== sort_bench warp_specialized=$ws"; DARK_BWT_WARP_SPECIALIZED=$ws DARK_BWT_EARLY_LOOKBACK=$ws timeout 300 python tools/sort_bench.py 27 ${VARS:-1} 2>&1 | grep variant | cut -c1-120; done
