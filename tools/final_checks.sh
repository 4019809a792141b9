python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r2_final_tests.log
python bench.py > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err
python bench.py --impl reference --steps 2 --warmup 0 > gpurun_out/r2_bench_reference_arm.json 2> gpurun_out/r2_bench_reference_arm.err
python tools/int_peak.py > gpurun_out/r2_int_peak.log 2>&1
BK_LIB=$PWD/breakmer_b200/lib/libbreakmer_b200_prof.so python tools/phase_profile.py C2 500 > gpurun_out/r2_phase_c2.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1
tail -3 gpurun_out/r2_final_tests.log; tail -2 gpurun_out/r2_smoke.log; tail -c 300 gpurun_out/r2_bench_default.err
