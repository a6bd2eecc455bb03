#!/bin/bash
# Run on the GPU box (via gpurun): staged GPU tests, each stage in its own process under a timeout
# so that a hung kernel in one stage does not lose the others.  Logs -> gpurun_out/.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
OUT=gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > $OUT/gpu.txt 2>&1
nproc > $OUT/nproc.txt; free -g >> $OUT/nproc.txt
run() { # name, timeout, command...
  local name=$1 to=$2; shift 2
  echo "=== $name" | tee -a $OUT/summary.txt
  timeout $to "$@" > $OUT/$name.log 2>&1
  local rc=$?
  echo "rc=$rc" | tee -a $OUT/summary.txt
  tail -n 4 $OUT/$name.log | tee -a $OUT/summary.txt
}
: > $OUT/summary.txt
PT="python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 200 -p no:cacheprovider"
run smoke 300 python -c "import __graft_entry__ as g; g.smoke()"
run t_kat 300 $PT -k "known_answers or error_behaviour"
run t_sort 300 $PT -k "radix_sort"
run t_blocks 300 $PT -k "emission or verifier"
run t_small 600 $PT -k "small_random or edge"
run t_medium 600 $PT -k "medium"
run t_misc 300 $PT -k "u64 or idempotent"
run t_large 900 $PT -k "large" -s
run bench 600 python bench.py --steps 5 --warmup 3
cat $OUT/bench.log | tail -n 1 > $OUT/bench.json
