# parameter sweep of the two run-time tunables (skin, steps per CUDA graph) on the T4L workload
mkdir -p gpurun_out
out=gpurun_out/sweep1.log
: > $out
for skin in 0.14 0.18 0.22 0.26 0.30; do
  echo "== R=1 skin=$skin" >> $out
  BLUES_B200_SKIN=$skin timeout 120 python -m tests.gpu_perf_probe 1 600 2>&1 | grep -E "graphs|rebuilds|pair  |neighbor" | tail -5 >> $out
done
for gs in 8 16; do
  echo "== R=1 graph_steps=$gs" >> $out
  BLUES_B200_GRAPH_STEPS=$gs timeout 120 python -m tests.gpu_perf_probe 1 600 2>&1 | grep -E "graphs" | tail -2 >> $out
done
for skin in 0.10 0.14 0.18; do
  echo "== R=8 skin=$skin" >> $out
  BLUES_B200_SKIN=$skin timeout 120 python -m tests.gpu_perf_probe 8 150 2>&1 | grep -E "graphs|rebuilds|pair  |neighbor" | tail -5 >> $out
done
cat $out
python bench.py --workload tolparm > gpurun_out/bench8_tolparm.json 2> gpurun_out/bench8_tolparm.err; tail -c 300 gpurun_out/bench8_tolparm.err; cut -c1-200 gpurun_out/bench8_tolparm.json
