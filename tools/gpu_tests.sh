#!/bin/bash
# Runs the GPU test-suite in separate processes (a CUDA fault poisons the context of the process that hit it).
mkdir -p gpurun_out
for sel in "tests/test_gpu_kernels.py -k 'not attention'" "tests/test_gpu_kernels.py -k attention" \
           "tests/test_gpu_path.py -k 'ctc_prefix'" "tests/test_gpu_path.py -k 'beam_search or batched_decode'" \
           "tests/test_gpu_path.py -k 'encoder'" "tests/test_gpu_path.py -k 'full_path'"; do
  echo "=== $sel" >> gpurun_out/pytest.log
  eval timeout 900 python -m pytest $sel -m gpu -q --timeout 600 -p no:cacheprovider >> gpurun_out/pytest.log 2>&1
  echo "exit=$?" >> gpurun_out/pytest.log
done
grep -E "^===|passed|failed|exit=" gpurun_out/pytest.log
