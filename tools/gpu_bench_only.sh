set -x
mkdir -p gpurun_out
T=${TAG:-r02b}
python bench.py > gpurun_out/bench_$T.json 2> gpurun_out/bench_$T.err; echo rc=$?; tail -c 300 gpurun_out/bench_$T.err; cut -c1-400 gpurun_out/bench_$T.json
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_ref_$T.json 2> gpurun_out/bench_ref_$T.err; echo rc=$?; tail -c 300 gpurun_out/bench_ref_$T.err; cut -c1-400 gpurun_out/bench_ref_$T.json
python bench.py --workload t4l_frozen --no-cpu-baseline > gpurun_out/bench_frozen_$T.json 2> gpurun_out/bench_frozen_$T.err; echo rc=$?; tail -c 300 gpurun_out/bench_frozen_$T.err; cut -c1-400 gpurun_out/bench_frozen_$T.json
