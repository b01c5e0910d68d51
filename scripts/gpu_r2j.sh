set -x
mkdir -p gpurun_out
./scripts/perm_debug.bin > gpurun_out/r2j_perm.log 2>&1; cat gpurun_out/r2j_perm.log
PK_POLAR_SC_G=4 timeout 600 python -X faulthandler -m pytest tests/test_gpu_polar.py -m gpu -q -x > gpurun_out/r2j_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2j_pytest.log
tail -5 gpurun_out/r2j_pytest.log
PK_POLAR_SC_G=4 timeout 200 python profiles/prof_polar.py 1 65536 2.0 > gpurun_out/r2j_polar_plain.log 2>&1 && \
PK_POLAR_SC_G=4 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_polar_sc -s 1 -c 1 -o gpurun_out/prof_r2j_polar_sc4 python profiles/prof_polar.py 1 65536 2.0 > gpurun_out/r2j_polar_ncu.log 2>&1
cat gpurun_out/r2j_polar_plain.log; tail -3 gpurun_out/r2j_polar_ncu.log
