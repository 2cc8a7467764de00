#!/bin/bash
# Runs the whole GPU test-suite, one process per test file (a CUDA fault poisons the context of the process that hit it).
# Usage: tools/gpu_tests.sh [extra pytest args]; summary in gpurun_out/pytest.log
mkdir -p gpurun_out
: > gpurun_out/pytest.log
for f in $(grep -l "mark.gpu" tests/test_*.py | sort); do
  echo "=== $f" >> gpurun_out/pytest.log
  timeout 1500 python -m pytest "$f" -m gpu -q --timeout 900 -p no:cacheprovider "$@" >> gpurun_out/pytest.log 2>&1
  echo "exit=$?" >> gpurun_out/pytest.log
done
grep -E "^===|passed|failed|error|exit=" gpurun_out/pytest.log
