# N = 1 / 2 / 4 / 8 on one box (run under `gpurun --gpus 8`): the bench lines the driver's scaling run should reproduce
run() { n=$1; port=$2
  if [ $n -eq 1 ]; then python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r02_scale_n1.json 2>> gpurun_out/n8.err
  else timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r02_scale_n$n.json 2>> gpurun_out/n8.err; fi
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_r02_scale_n$n.json").read().strip().splitlines()[-1])
    print("N=$n", d["ms_per_step"], d["value"], d["e2e"]["value"])
except Exception as e:
    print("N=$n FAILED", e)
PY
}
rm -f gpurun_out/n8.err
run 1 0; run 2 29531; run 4 29532; run 8 29533
grep -v Warning gpurun_out/n8.err | grep -v "return Variable" | grep -v "^\*" | tail -3
