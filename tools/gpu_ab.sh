# same-box A/B of builds of the library (tools/variants/*.so)
mkdir -p gpurun_out
out=gpurun_out/ab4.log
: > $out
cp blues_b200/libblues_b200.so /tmp/lib_keep.so
for rep in 1 2 3; do
for v in tools/variants/*.so; do
  cp $v blues_b200/libblues_b200.so
  echo "== $v R=1 rep=$rep" >> $out
  timeout 120 python -m tests.gpu_perf_probe 1 600 2>&1 | grep -E "graphs|pair  " | tail -2 >> $out
done
done
for v in tools/variants/*.so; do
  cp $v blues_b200/libblues_b200.so
  for skin in 0.14 0.12; do
  echo "== $v R=8 skin=$skin" >> $out
  BLUES_B200_SKIN=$skin timeout 120 python -m tests.gpu_perf_probe 8 150 2>&1 | grep -E "graphs|pair  " | tail -2 >> $out
  done
done
cp /tmp/lib_keep.so blues_b200/libblues_b200.so
cat $out
