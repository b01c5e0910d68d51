set -x
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2a_gpu.txt; nproc >> gpurun_out/r2a_gpu.txt
timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2a_pytest.log
tail -5 gpurun_out/r2a_pytest.log
timeout 300 python profiles/prof_tail.py 5 3 -1 0.0 65536 > gpurun_out/r2a_tail_m5.log 2>&1
timeout 300 python profiles/prof_tail.py 7 10 15 3.0 4096 4194304 > gpurun_out/r2a_tail_m7.log 2>&1
timeout 300 python profiles/prof_tail.py 8 15 15 4.0 4096 4194304 > gpurun_out/r2a_tail_m8.log 2>&1
timeout 300 python profiles/prof_replay.py 7 10 15 3.0 4096 1 2 65536 > gpurun_out/r2a_m7_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_phase_b -s 1 -c 1 -o gpurun_out/prof_r2a_phaseb_m7t10 python profiles/prof_replay.py 7 10 15 3.0 4096 1 2 65536 > gpurun_out/r2a_m7_ncu.log 2>&1
timeout 300 python profiles/prof_replay.py 8 15 15 4.0 4096 1 2 65536 > gpurun_out/r2a_m8_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_phase_b -s 1 -c 1 -o gpurun_out/prof_r2a_phaseb_m8t15 python profiles/prof_replay.py 8 15 15 4.0 4096 1 2 65536 > gpurun_out/r2a_m8_ncu.log 2>&1
cat gpurun_out/r2a_tail_m5.log gpurun_out/r2a_m7_plain.log gpurun_out/r2a_m8_plain.log
