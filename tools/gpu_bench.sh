#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
OUT=gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 600 -p no:cacheprovider -k "batch or known" > $OUT/pytest_batch.log 2>&1; echo "pytest rc=$?"; tail -n 3 $OUT/pytest_batch.log
timeout 900 python bench.py --steps ${STEPS:-10} --warmup 3 > $OUT/bench.log 2> $OUT/bench.err; echo "bench rc=$?"; tail -n 1 $OUT/bench.log > $OUT/bench.json; cat $OUT/bench.json; tail -n 5 $OUT/bench.err
