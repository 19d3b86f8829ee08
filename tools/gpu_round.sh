set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu11.log 2>&1; tail -5 gpurun_out/pytest_gpu11.log
python bench.py > gpurun_out/bench7.json 2> gpurun_out/bench7.err; tail -c 400 gpurun_out/bench7.err; cut -c1-300 gpurun_out/bench7.json
python bench.py --workload tolparm > gpurun_out/bench7_tolparm.json 2> gpurun_out/bench7_tolparm.err; tail -c 400 gpurun_out/bench7_tolparm.err; cut -c1-200 gpurun_out/bench7_tolparm.json
python bench.py --workload water > gpurun_out/bench7_water.json 2> gpurun_out/bench7_water.err; tail -c 400 gpurun_out/bench7_water.err; cut -c1-200 gpurun_out/bench7_water.json
timeout 400 python bench.py --workload m5 --steps 200 --warmup 20 > gpurun_out/bench7_m5.json 2> gpurun_out/bench7_m5.err; tail -c 400 gpurun_out/bench7_m5.err; cut -c1-200 gpurun_out/bench7_m5.json
