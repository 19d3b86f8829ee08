# one --set full capture with dense PC sampling: KERNEL=<regex> TAG=<name> [SKIP=n] [COUNT=n] [R=walkers]
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on --warp-sampling-interval 0 -k regex:"${KERNEL}" -s ${SKIP:-20} -c ${COUNT:-4} -f \
    -o gpurun_out/prof_${TAG} python -m tests.gpu_ncu_target ${R:-1} 30 > gpurun_out/ncu_${TAG}.log 2>&1
tail -2 gpurun_out/ncu_${TAG}.log
ncu -i gpurun_out/prof_${TAG}.ncu-rep --page raw --csv > gpurun_out/prof_${TAG}_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_${TAG}.ncu-rep --page source --csv > gpurun_out/prof_${TAG}_src.csv 2>/dev/null
ls -la gpurun_out/prof_${TAG}*
