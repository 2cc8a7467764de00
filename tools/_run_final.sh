mkdir -p gpurun_out
timeout 300 python tools/bench_input_pipeline.py --no-oracle > gpurun_out/input_bench2.log 2>&1; echo "input bench rc=$?"
cat gpurun_out/input_bench2.log
timeout 1500 python -m pytest tests/ -x -q -m gpu > gpurun_out/final_tests.log 2>&1; echo "tests rc=$?"
tail -5 gpurun_out/final_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/final_smoke.log 2>&1; echo "smoke rc=$?"
tail -2 gpurun_out/final_smoke.log
timeout 600 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench rc=$?"
cat gpurun_out/final_bench.json
