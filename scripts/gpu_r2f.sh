set -x
mkdir -p gpurun_out
timeout 2400 python -X faulthandler -m pytest tests -m gpu -x -q > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2f_pytest.log
tail -30 gpurun_out/r2f_pytest.log
timeout 900 python bench.py --steps 2 --warmup 1 --detail --no-side > gpurun_out/r2f_bench1.json 2> gpurun_out/r2f_bench1.err; echo "rc $?"; tail -24 gpurun_out/r2f_bench1.err
timeout 600 python bench.py --workload polar --steps 1 --warmup 1 > gpurun_out/r2f_polar.json 2> gpurun_out/r2f_polar.err; echo "rc $?"; tail -5 gpurun_out/r2f_polar.err
timeout 300 python profiles/prof_fixed.py 7 10 16384 32768 2 > gpurun_out/r2f_fixed_m7_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_phase_b -s 1 -c 1 -o gpurun_out/prof_r2f_fixed_m7t10 python profiles/prof_fixed.py 7 10 16384 32768 2 > gpurun_out/r2f_fixed_m7_ncu.log 2>&1
cat gpurun_out/r2f_fixed_m7_plain.log
timeout 300 python profiles/prof_replay.py 6 6 15 0.0 65536 1 2 > gpurun_out/r2f_ct_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_phase_b -s 1 -c 1 -o gpurun_out/prof_r2f_ct_m6t6 python profiles/prof_replay.py 6 6 15 0.0 65536 1 2 > gpurun_out/r2f_ct_ncu.log 2>&1
cat gpurun_out/r2f_ct_plain.log
