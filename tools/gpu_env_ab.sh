# same-box A/B of environment toggles: VARIANTS="A=1;B=2 C=3;..." (space-separated variants, ';' between assignments, 'base' = none)
# R="1 8" walkers, STEPS1 / STEPS8 steps per timed run, TAG log name
mkdir -p gpurun_out
out=gpurun_out/ab_${TAG:-x}.log
: > $out
for rep in 1 2; do
for v in ${VARIANTS:-base}; do
  for r in ${R:-1 8}; do
    st=${STEPS1:-400}; [ "$r" != "1" ] && st=${STEPS8:-150}
    echo "== $v R=$r rep=$rep" >> $out
    ( [ "$v" != "base" ] && export $(echo $v | tr ';' ' '); timeout 200 python -m tests.gpu_perf_probe $r $st 2>&1 | grep -E "graphs|work|integrate|pair |neighbor|pme_spread|fft|sum" | tail -${TAILN:-9} >> $out )
  done
done
done
cat $out
