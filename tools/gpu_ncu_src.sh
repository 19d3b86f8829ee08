# source-page CSV of one kernel: KERNEL=<regex> TAG=<name> [SKIP=n] [R=walkers]; the report itself stays on the box
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"${KERNEL}" -s ${SKIP:-2} -c 1 -f \
    -o /tmp/prof_${TAG} python -m tests.gpu_ncu_target ${R:-1} 30 > gpurun_out/ncu_${TAG}.log 2>&1
tail -2 gpurun_out/ncu_${TAG}.log
ncu -i /tmp/prof_${TAG}.ncu-rep --page source --csv > gpurun_out/prof_${TAG}_src.csv 2>/dev/null
ncu -i /tmp/prof_${TAG}.ncu-rep --page raw --csv > gpurun_out/prof_${TAG}_raw.csv 2>/dev/null
ls -la gpurun_out/prof_${TAG}*
