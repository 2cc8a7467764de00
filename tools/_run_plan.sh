mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_path.py -x -q -k "large_decode or sharded_evaluation" > gpurun_out/plan_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/plan_tests.log
timeout 600 python tools/eval_cfg3.py > gpurun_out/cfg3_plan_1gpu.json 2> gpurun_out/cfg3_plan_1gpu.err; echo "rc=$?"; cat gpurun_out/cfg3_plan_1gpu.json; tail -2 gpurun_out/cfg3_plan_1gpu.err
timeout 600 python tools/eval_cfg3.py --host-inputs > gpurun_out/cfg3_plan_1gpu_host.json 2> gpurun_out/cfg3_plan_1gpu_host.err; echo "rc=$?"; cat gpurun_out/cfg3_plan_1gpu_host.json; tail -2 gpurun_out/cfg3_plan_1gpu_host.err
