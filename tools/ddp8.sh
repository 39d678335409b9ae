# N = 8 A/B of the gradient-synchronisation variants (run under `gpurun --gpus 8`)
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
run() { name=$1; shift; port=$1; shift
  timeout 600 $R --master-port $port bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/bench_n8_$name.json 2>> gpurun_out/n8.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_n8_$name.json").read().strip().splitlines()[-1])
    print("$name", d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"], d["config"]["ddp"][:60])
except Exception as e:
    print("$name FAILED", e)
PY
}
rm -f gpurun_out/n8.err
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1_ref8.json 2>> gpurun_out/n8.err; python -c "import json;d=json.loads(open('gpurun_out/bench_n1_ref8.json').read().strip().splitlines()[-1]);print('n1', d['ms_per_step'], d['value'], d['e2e']['ms_per_step'])"
run flat 29521 --ddp-impl flat
run torch 29522 --ddp-impl torch
grep -v Warning gpurun_out/n8.err | grep -v "return Variable" | tail -3
