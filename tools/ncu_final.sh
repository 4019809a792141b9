# Final-build capture of the dominant kernel (one GPU, under gpurun): --set full with source, second launch of two
# sequential C2 passes.  The .ncu-rep travels back (about 20 MB); details and raw pages are also exported on the box.
mkdir -p gpurun_out
timeout 170 ncu --set full --clock-control none --import-source on -k regex:assemble_kernel -s 1 -c 1 \
    -o gpurun_out/r2_assemble_final -f python tools/phase_profile.py C2 500 > gpurun_out/r2_assemble_final.log 2>&1
ncu -i gpurun_out/r2_assemble_final.ncu-rep --page details --csv > gpurun_out/r2_assemble_final_details.csv 2>/dev/null
ncu -i gpurun_out/r2_assemble_final.ncu-rep --page raw --csv > gpurun_out/r2_assemble_final_raw.csv 2>/dev/null
ls -la gpurun_out/ | tail -6
