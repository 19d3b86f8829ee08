# builder flush experiment: correctness (neighbour sets, parity) then timing against the unflushed sub-lists (BUILD_CQ=308)
mkdir -p gpurun_out
out=gpurun_out/builder1.log
: > $out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_builder1.log 2>&1; tail -4 gpurun_out/pytest_builder1.log >> $out
for cq in 88 56 152 308; do
  echo "== R=1 cq=$cq" >> $out
  BLUES_B200_BUILD_CQ=$cq timeout 120 python -m tests.gpu_perf_probe 1 600 2>&1 | grep -E "graphs|neighbor|work" | tail -4 >> $out
  echo "== R=8 cq=$cq" >> $out
  BLUES_B200_BUILD_CQ=$cq timeout 120 python -m tests.gpu_perf_probe 8 150 2>&1 | grep -E "graphs|neighbor|work" | tail -4 >> $out
done
cat $out
timeout 400 python bench.py --workload m5 --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/bench9_m5.json 2> gpurun_out/bench9_m5.err; tail -c 300 gpurun_out/bench9_m5.err; cut -c1-180 gpurun_out/bench9_m5.json
