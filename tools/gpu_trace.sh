#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for v in ${VARS:-1 5 0}; do timeout 300 python tools/pass_trace.py 27 $v 2>&1 | tee -a gpurun_out/pass_trace.log; done
