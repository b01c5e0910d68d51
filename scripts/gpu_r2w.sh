set -x
mkdir -p gpurun_out
timeout 900 python -X faulthandler -m pytest tests/test_gpu_parity.py tests/test_gpu_fer.py -m gpu -q -x -k "not 63_16 and not 63_30" > gpurun_out/r2w_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2w_pytest.log
tail -3 gpurun_out/r2w_pytest.log
timeout 300 python profiles/prof_replay.py 6 6 15 0.0 65536 1 2 > gpurun_out/r2w_ct_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --cache-control none --import-source on -k regex:k_phase_b -s 1 -c 1 -o gpurun_out/prof_r2w_ct_warm python profiles/prof_replay.py 6 6 15 0.0 65536 1 2 > gpurun_out/r2w_ct_ncu.log 2>&1
cat gpurun_out/r2w_ct_plain.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_phase_b -s 1 -c 1 -o gpurun_out/prof_r2w_ct_cold python profiles/prof_replay.py 6 6 15 0.0 65536 1 2 > gpurun_out/r2w_ct_ncu2.log 2>&1
tail -1 gpurun_out/r2w_ct_ncu2.log
