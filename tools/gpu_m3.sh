# M3 (BASELINE configs[2]): 64 walkers over the GPUs of the box, 64 / N per GPU
mkdir -p gpurun_out
N=${1:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --replicas $((64 / N)) --steps 150 --warmup 20 > gpurun_out/bench_m3_${N}gpu.json 2> gpurun_out/bench_m3_${N}gpu.err
tail -c 400 gpurun_out/bench_m3_${N}gpu.err; cut -c1-400 gpurun_out/bench_m3_${N}gpu.json
