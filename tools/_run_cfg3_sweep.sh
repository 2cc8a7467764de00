mkdir -p gpurun_out
for cfg in "192 12288" "256 16384" "384 24576"; do
  set -- $cfg
  timeout 600 python tools/eval_cfg3.py --max-utts $1 --max-frames $2 > gpurun_out/cfg3_mu$1.json 2> gpurun_out/cfg3_mu$1.err; echo "mu=$1 mf=$2 rc=$?"
  cat gpurun_out/cfg3_mu$1.json; tail -2 gpurun_out/cfg3_mu$1.err
done
