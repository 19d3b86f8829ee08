# same-box A/B of two builds of the library: LIBS="blues_b200/libA.so blues_b200/libB.so" (copied over the product .so in turn)
mkdir -p gpurun_out
cp blues_b200/libblues_b200.so /tmp/lib_product.so
for rep in ${REPS:-1 2}; do
for lib in /tmp/lib_product.so $LIBS; do
  cp $lib blues_b200/libblues_b200.so
  for r in ${R:-1 8}; do
    st=${STEPS1:-600}; [ "$r" != "1" ] && st=150
    echo "== $(basename $lib) R=$r rep=$rep"
    timeout 200 python -m tests.gpu_perf_probe $r $st 2>&1 | grep -E "graphs|neighbor" | tail -2
  done
done
done
cp /tmp/lib_product.so blues_b200/libblues_b200.so
