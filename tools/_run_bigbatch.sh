mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_path.py -x -q -k "large_decode or sharded_evaluation or long_utterances" > gpurun_out/bigbatch_tests.log 2>&1; echo "tests rc=$?"
tail -15 gpurun_out/bigbatch_tests.log
