# --set full capture of the hot kernels of a few steps (the report stays on the box, its raw page comes home)
mkdir -p gpurun_out
T=${TAG:-r02f}
ncu --set full --clock-control none --import-source on -k regex:"k_pair4|k_integrate$|k_pme_spread|k_pme_gather5|k_pme_convolve|k_sort_atoms|k_bonded|k_alch$" -s ${SKIP:-80} -c ${COUNT:-16} -f -o /tmp/prof_${T}_top \
    python -m tests.gpu_ncu_target ${R:-1} 30 > gpurun_out/ncu_${T}_top.log 2>&1
tail -2 gpurun_out/ncu_${T}_top.log
ncu -i /tmp/prof_${T}_top.ncu-rep --page raw --csv > gpurun_out/prof_${T}_top_raw.csv 2>/dev/null
ls -la gpurun_out/prof_${T}_top_raw.csv
