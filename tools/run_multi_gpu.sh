#!/bin/bash
# One multi-GPU measurement pass: configs[2] (sharded set), configs[3] (AVCocktail-shaped chunks) and the headline configs[1]
# bench under torchrun on N GPUs of one box.  Usage: tools/run_multi_gpu.sh N   (results: gpurun_out/multi_N_*.json)
N=$1
mkdir -p gpurun_out
run() {  # name, args...
  name=$1; shift
  if [ "$N" = "1" ]; then python bench.py --gpus 1 "$@" > gpurun_out/multi_${N}_${name}.json 2> gpurun_out/multi_${N}_${name}.err
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@" \
         > gpurun_out/multi_${N}_${name}.json 2> gpurun_out/multi_${N}_${name}.err; fi
  tail -n 1 gpurun_out/multi_${N}_${name}.json | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$name N=$N', round(d['value'],1), 'ms', round(d['ms_per_step'],1), d.get('ms_all_steps',''))
except Exception as e: print('$name N=$N FAILED', e)"
}
run cfg2 --workload cfg2 --steps 3 --warmup 2
run cfg3 --workload cfg3 --steps 2 --warmup 1
run cfg1 --steps 5 --warmup 3 --no-cpu-baseline --no-gpu-baseline
