# Round-2 ncu captures (run under gpurun, one GPU): launch list of two sequential C2 passes, and a --set full capture
# of every kernel of the second pass, exported to CSV on the box (the .ncu-rep of 81 kernels exceeds the 64 MiB that
# travel back); source-level captures of the two dominant kernels are kept as .ncu-rep.
set -x
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv python tools/phase_profile.py C2 500 > gpurun_out/r2_launches.log 2>&1
N=$(grep -c "gpu__time_duration.sum" gpurun_out/r2_launches.csv)
HALF=$((N / 2))
ncu --set full --clock-control none -s $HALF -c $HALF -o /tmp/r2_full -f python tools/phase_profile.py C2 500 > gpurun_out/r2_full.log 2>&1
ncu -i /tmp/r2_full.ncu-rep --page raw --csv > gpurun_out/r2_full_raw.csv
gzip -f gpurun_out/r2_full_raw.csv
ncu --set full --clock-control none --import-source on -k regex:assemble_kernel -s 1 -c 1 -o gpurun_out/r2_assemble_src -f python tools/phase_profile.py C2 500 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:region_kmer_kernel -s 1 -c 1 -o gpurun_out/r2_region_kmers_src -f python tools/phase_profile.py C2 500 > /dev/null 2>&1
ls -la gpurun_out/ | tail -8
