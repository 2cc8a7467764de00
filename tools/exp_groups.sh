for cfg in "1 cluster 0" "2 cluster 74" "2 cluster 148" "2 cluster 100" "3 cluster 49" "4 cluster 37" "2 splitk 0"; do
  set -- $cfg
  AVSR_DECODE_GROUPS=$1 AVSR_PROJ=$2 AVSR_SM_BUDGET=$3 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-gpu-baseline > gpurun_out/exp1_$1_$2_$3.json 2> gpurun_out/exp1_$1_$2_$3.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/exp1_$1_$2_$3.json').read().strip().splitlines()[-1])
    print('groups $1 proj $2 budget $3:', 'value',round(d['value'],1),'ms',round(d['ms_per_step'],1))
except Exception as e:
    print('groups $1 proj $2 budget $3: ERR', open('gpurun_out/exp1_$1_$2_$3.err').read()[-600:])
PY
done
