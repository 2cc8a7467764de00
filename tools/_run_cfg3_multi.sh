mkdir -p gpurun_out
N=${1:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 tools/eval_cfg3.py > gpurun_out/cfg3_${N}gpu.json 2> gpurun_out/cfg3_${N}gpu.err; echo "rc=$?"
cat gpurun_out/cfg3_${N}gpu.json; tail -3 gpurun_out/cfg3_${N}gpu.err
