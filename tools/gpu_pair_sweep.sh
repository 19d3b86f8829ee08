# pair-kernel variants (BLUES_B200_PAIR, see enqueue_eval) on the T4L workload: parity of forces first, then timing
mkdir -p gpurun_out
out=gpurun_out/pair_sweep.log
: > $out
for v in 0 32 34 132 1030 1130 1140 1120; do
  echo "== variant $v parity" >> $out
  BLUES_B200_PAIR=$v timeout 200 python - >> $out 2>&1 <<'PY'
from tests import gpu_checks as gc
for name in ('tol_parm', 't4l_surrogate'):
    o = gc.compare_forces(name, alchemical=False)
    print(name, 'energy_rel %.2e force_max_rel %.2e force_rms_rel %.2e' % (o['energy_rel'], o['force_max_rel'], o['force_rms_rel']))
PY
  echo "== variant $v R=1" >> $out
  BLUES_B200_PAIR=$v timeout 200 python -m tests.gpu_perf_probe 1 600 2>&1 | grep -E "graphs|pair  |neighbor|profiling" | tail -5 >> $out
  echo "== variant $v R=8" >> $out
  BLUES_B200_PAIR=$v timeout 200 python -m tests.gpu_perf_probe 8 150 2>&1 | grep -E "graphs|pair  |neighbor|profiling" | tail -5 >> $out
done
cat $out
