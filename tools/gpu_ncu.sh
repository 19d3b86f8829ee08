# ncu evidence for profiles/: launch list of ~30 consecutive NCMC steps and a --set full capture of the top kernels
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -s 1500 -c 700 --csv \
    --log-file gpurun_out/launches_r02f.csv python -m tests.gpu_ncu_target 1 60 > gpurun_out/ncu_r02f_launch.log 2>&1
tail -2 gpurun_out/ncu_r02f_launch.log
ncu --set full --clock-control none --import-source on -k regex:"k_build_list" -s 30 -c 14 -f -o gpurun_out/prof_r02f_build \
    python -m tests.gpu_ncu_target 1 30 > gpurun_out/ncu_r02f_build.log 2>&1
tail -2 gpurun_out/ncu_r02f_build.log
ncu --set full --clock-control none --import-source on -k regex:"k_pair4|k_integrate$|k_pme_spread|k_pme_gather5|k_pme_convolve" -s 420 -c 12 -f -o gpurun_out/prof_r02f_top \
    python -m tests.gpu_ncu_target 1 30 > gpurun_out/ncu_r02f_top.log 2>&1
tail -2 gpurun_out/ncu_r02f_top.log
ncu -i gpurun_out/prof_r02f_build.ncu-rep --page raw --csv > gpurun_out/prof_r02f_build_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_r02f_top.ncu-rep --page raw --csv > gpurun_out/prof_r02f_top_raw.csv 2>/dev/null
ls -la gpurun_out/
# gpurun brings back at most 64 MiB: the reports stay on the box, their CSV pages come home
ncu -i gpurun_out/prof_r02f_top.ncu-rep --page source --csv -k regex:"k_pair4" > gpurun_out/prof_r02f_pair4_src.csv 2>/dev/null
rm -f gpurun_out/*.ncu-rep
ls -la gpurun_out/
