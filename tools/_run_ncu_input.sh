mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'fbank_stack' -s 3 -c 1 -o gpurun_out/ncu_fbank -f python tools/bench_input_pipeline.py --no-oracle --iters 1 > gpurun_out/ncu_fbank.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/ncu_fbank.ncu-rep
