# same-box A/B of builds of the library (tools/variants/*.so)
mkdir -p gpurun_out
out=gpurun_out/ab2.log
: > $out
cp blues_b200/libblues_b200.so /tmp/lib_keep.so
for rep in 1 2; do
for v in tools/variants/*.so; do
  cp $v blues_b200/libblues_b200.so
  echo "== $v R=1 rep=$rep" >> $out
  timeout 120 python -m tests.gpu_perf_probe 1 600 2>&1 | grep -E "graphs|neighbor" | tail -2 >> $out
done
done
for v in tools/variants/*.so; do
  cp $v blues_b200/libblues_b200.so
  echo "== $v R=8" >> $out
  BLUES_B200_BUILD_CQ=56 timeout 120 python -m tests.gpu_perf_probe 8 150 2>&1 | grep -E "graphs|neighbor" | tail -2 >> $out
done
cp /tmp/lib_keep.so blues_b200/libblues_b200.so
cat $out
