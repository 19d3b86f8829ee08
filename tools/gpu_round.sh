# full GPU round: tests, driver-style bench, workload lines
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu12.log 2>&1; tail -3 gpurun_out/pytest_gpu12.log
python bench.py > gpurun_out/bench10.json 2> gpurun_out/bench10.err; tail -c 300 gpurun_out/bench10.err; cut -c1-200 gpurun_out/bench10.json
timeout 400 python bench.py --workload m5 --steps 200 --warmup 20 > gpurun_out/bench10_m5.json 2> gpurun_out/bench10_m5.err; tail -c 300 gpurun_out/bench10_m5.err; cut -c1-200 gpurun_out/bench10_m5.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench10_ref.json 2> gpurun_out/bench10_ref.err; cut -c1-200 gpurun_out/bench10_ref.json
