# N = 2 A/B of the gradient-synchronisation variants (run under `gpurun --gpus 2`): prints ms_per_step / value / e2e of each
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
run() { name=$1; shift; port=$1; shift
  timeout 600 $R --master-port $port bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/bench_n2_$name.json 2>> gpurun_out/n2.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_n2_$name.json").read().strip().splitlines()[-1])
    print("$name", d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"], d["config"]["ddp"][:90])
except Exception as e:
    print("$name FAILED", e)
PY
}
rm -f gpurun_out/n2.err
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1_ref.json 2>> gpurun_out/n2.err; python -c "import json;d=json.loads(open('gpurun_out/bench_n1_ref.json').read().strip().splitlines()[-1]);print('n1', d['ms_per_step'], d['value'], d['e2e']['ms_per_step'])"
run flat 29511 --ddp-impl flat
run torch 29512 --ddp-impl torch
NCCL_MAX_NCHANNELS=4 run flat_ch4 29513 --ddp-impl flat
tail -5 gpurun_out/n2.err
