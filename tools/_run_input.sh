set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_input.py -x -q > gpurun_out/input_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/input_tests.log
timeout 300 python tools/bench_input_pipeline.py > gpurun_out/input_bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/input_bench.log
tail -30 gpurun_out/input_tests.log; cat gpurun_out/input_bench.log
