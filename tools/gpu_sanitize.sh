#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TOOL=${1:-memcheck}
python tools/sanitize_probe.py > gpurun_out/sanitize_plain.log 2>&1 || { echo "plain run failed"; tail gpurun_out/sanitize_plain.log; exit 1; }
for env in "" "DARK_BWT_BUCKETED=1 DARK_BWT_EMIT_WINDOW_MB=1 DARK_BWT_FORCE_U64_STATUS=1"; do
  echo "== env: $env" | tee -a gpurun_out/sanitize_$TOOL.log
  env $env timeout 900 compute-sanitizer --tool $TOOL --error-exitcode 9 python tools/sanitize_probe.py >> gpurun_out/sanitize_$TOOL.log 2>&1; echo "rc=$?" | tee -a gpurun_out/sanitize_$TOOL.log
done
grep -E "ERROR SUMMARY|rc=|probe ok|== env" gpurun_out/sanitize_$TOOL.log
