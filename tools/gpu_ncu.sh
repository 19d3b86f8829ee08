# ncu evidence for profiles/: launch list of ~30 consecutive NCMC steps and a --set full capture of the top kernels
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -s 1500 -c 700 --csv \
    --log-file gpurun_out/launches_r02a.csv python -m tests.gpu_ncu_target 1 60 > gpurun_out/ncu_r02a_launch.log 2>&1
tail -2 gpurun_out/ncu_r02a_launch.log
ncu --set full --clock-control none --import-source on -k regex:"k_build_list" -s 30 -c 14 -f -o gpurun_out/prof_r02a_build \
    python -m tests.gpu_ncu_target 1 30 > gpurun_out/ncu_r02a_build.log 2>&1
tail -2 gpurun_out/ncu_r02a_build.log
ncu --set full --clock-control none --import-source on -k regex:"k_pair|k_integrate|k_pme_spread|k_pme_gather5" -s 130 -c 10 -f -o gpurun_out/prof_r02a_top \
    python -m tests.gpu_ncu_target 1 30 > gpurun_out/ncu_r02a_top.log 2>&1
tail -2 gpurun_out/ncu_r02a_top.log
ncu -i gpurun_out/prof_r02a_build.ncu-rep --page raw --csv > gpurun_out/prof_r02a_build_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_r02a_top.ncu-rep --page raw --csv > gpurun_out/prof_r02a_top_raw.csv 2>/dev/null
ls -la gpurun_out/
