# A/B of run-time toggles on the same box
mkdir -p gpurun_out
out=gpurun_out/ab5.log
: > $out
for rep in 1 2; do
for pdl in 0 1; do
  echo "== PDL=$pdl R=1 rep=$rep" >> $out
  BLUES_B200_PDL=$pdl timeout 120 python -m tests.gpu_perf_probe 1 600 2>&1 | grep -E "graphs|work" | tail -2 >> $out
done
done
echo "== PDL=1 R=8" >> $out
BLUES_B200_PDL=1 timeout 120 python -m tests.gpu_perf_probe 8 150 2>&1 | grep -E "graphs" | tail -1 >> $out
cat $out
BLUES_B200_PDL=1 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
