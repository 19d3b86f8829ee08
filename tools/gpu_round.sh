# full GPU round: smoke, tests, driver-style bench (both arms), perf probes
set -x
mkdir -p gpurun_out
T=${TAG:-r02}
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$T.log 2>&1; tail -5 gpurun_out/pytest_gpu_$T.log
python bench.py > gpurun_out/bench_$T.json 2> gpurun_out/bench_$T.err; tail -c 300 gpurun_out/bench_$T.err; cut -c1-600 gpurun_out/bench_$T.json
python -m tests.gpu_perf_probe 1 300 > gpurun_out/probe1_$T.log 2>&1; tail -25 gpurun_out/probe1_$T.log
python -m tests.gpu_perf_probe 8 200 > gpurun_out/probe8_$T.log 2>&1; tail -25 gpurun_out/probe8_$T.log
