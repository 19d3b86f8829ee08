# the other SURVEY 8(d) configurations, one bench line each (GPU arm only)
mkdir -p gpurun_out
for w in ${WORKLOADS:-t4l_tol5e4 tolparm water m5 m5_t4l}; do
  timeout 600 python bench.py --workload $w --no-cpu-baseline --batched 0 --m3-walkers 0 > gpurun_out/bench_w_$w.json 2> gpurun_out/bench_w_$w.err
  echo "$w rc=$?"; grep -v "^\[W" gpurun_out/bench_w_$w.err | tail -2 | cut -c1-300
  python -c "
import json; d=json.load(open('gpurun_out/bench_w_$w.json')); print('$w', round(d['value'],1), 'steps/s', round(d['ms_per_step'],4), 'ms/step e2e', round(d['e2e']['value'],1), 'walkers/gpu', d['config']['replicas_per_gpu'], 'blown', d['blown_up_walkers']); print(d['kernels_us_per_step'])"
done
