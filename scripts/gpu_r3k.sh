set -x
mkdir -p gpurun_out
timeout 2400 python -X faulthandler -m pytest tests -m gpu -q > gpurun_out/r3k_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r3k_pytest.log
grep -n "FAILED\|passed\|failed\|rc " gpurun_out/r3k_pytest.log | tail -6
timeout 300 python profiles/prof_polar.py 8 16384 2.0 > gpurun_out/r3k_L8_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_polar_lanes -s 1 -c 1 -o gpurun_out/prof_r3k_lanes_L8_G2 python profiles/prof_polar.py 8 16384 2.0 > gpurun_out/r3k_L8_ncu.log 2>&1
tail -1 gpurun_out/r3k_L8_plain.log
timeout 300 python profiles/prof_polar.py 1 65536 2.0 > gpurun_out/r3k_L1_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_polar_lanes -s 1 -c 1 -o gpurun_out/prof_r3k_lanes_L1_G2 python profiles/prof_polar.py 1 65536 2.0 > gpurun_out/r3k_L1_ncu.log 2>&1
tail -1 gpurun_out/r3k_L1_plain.log
