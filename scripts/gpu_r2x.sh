set -x
mkdir -p gpurun_out
timeout 300 python profiles/prof_gen.py 6 6 15 5.0 1048576 > gpurun_out/r2x_gen_m6_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_phase_a -s 4 -c 1 -o gpurun_out/prof_r2x_phasea_gen_m6t6 python profiles/prof_gen.py 6 6 15 5.0 1048576 > gpurun_out/r2x_gen_m6_ncu.log 2>&1
cat gpurun_out/r2x_gen_m6_plain.log
timeout 300 python profiles/prof_gen.py 5 3 -1 5.0 1048576 > gpurun_out/r2x_gen_m5_plain.log 2>&1
cat gpurun_out/r2x_gen_m5_plain.log
