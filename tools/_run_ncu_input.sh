mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'video_u8|fbank_stack|wave_' -s 6 -c 4 -o gpurun_out/ncu_input -f python tools/bench_input_pipeline.py --no-oracle --iters 2 > gpurun_out/ncu_input.log 2>&1; echo "ncu rc=$?"
tail -5 gpurun_out/ncu_input.log; ls -la gpurun_out/ncu_input.ncu-rep
