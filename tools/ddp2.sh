R="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
$R --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --no-broadcast-buffers > gpurun_out/bench_r2_n2_nobb.json 2> gpurun_out/n2.err
NCCL_MAX_NCHANNELS=4 $R --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2_n2_ch4.json 2>> gpurun_out/n2.err
NCCL_MAX_NCHANNELS=4 $R --master-port 29513 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --no-broadcast-buffers > gpurun_out/bench_r2_n2_ch4_nobb.json 2>> gpurun_out/n2.err
NCCL_MAX_NCHANNELS=2 $R --master-port 29514 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --no-broadcast-buffers --bucket-mb 64 > gpurun_out/bench_r2_n2_ch2_b64.json 2>> gpurun_out/n2.err
$R --master-port 29515 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/bench_r2_n2_eager.json 2>> gpurun_out/n2.err
