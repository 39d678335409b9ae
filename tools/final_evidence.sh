set -x
python -m pytest tests -m gpu -x -q > gpurun_out/final_gputests.log 2>&1; tail -2 gpurun_out/final_gputests.log
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r02_final_n1.json 2> gpurun_out/final_bench.err
python bench.py --steps 20 --warmup 5 --no-overlap --no-cpu-baseline > gpurun_out/bench_r02_n1_no_overlap.json 2>> gpurun_out/final_bench.err
python bench.py --model medssd --steps 5 --warmup 3 > gpurun_out/bench_r02_medssd_final_n1.json 2>> gpurun_out/final_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r02_reference_arm.json 2>> gpurun_out/final_bench.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; tail -2 gpurun_out/final_smoke.log
for f in bench_r02_final_n1 bench_r02_n1_no_overlap bench_r02_medssd_final_n1 bench_r02_reference_arm; do python -c "
import json
d=json.loads(open('gpurun_out/$f.json').read().strip().splitlines()[-1]); print('$f', d.get('ms_per_step'), d.get('value'), (d.get('e2e') or {}).get('value'), (d.get('roofline') or {}).get('frac'), d.get('clocks'))"; done
