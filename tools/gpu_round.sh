# full GPU round: smoke, tests, driver-style bench (both arms)
set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu13.log 2>&1; tail -3 gpurun_out/pytest_gpu13.log
python bench.py > gpurun_out/bench11.json 2> gpurun_out/bench11.err; tail -c 300 gpurun_out/bench11.err; cut -c1-200 gpurun_out/bench11.json
