set -x
mkdir -p gpurun_out
timeout 120 python profiles/prof_tail.py 5 3 -1 0.0 65536 > gpurun_out/r2b_tail_m5.log 2>&1; echo "rc $?" >> gpurun_out/r2b_tail_m5.log
head -12 gpurun_out/r2b_tail_m5.log
timeout 120 python profiles/prof_tail.py 7 10 15 3.0 4096 4194304 > gpurun_out/r2b_tail_m7.log 2>&1; echo "rc $?" >> gpurun_out/r2b_tail_m7.log
head -12 gpurun_out/r2b_tail_m7.log
timeout 120 python profiles/prof_tail.py 8 15 15 4.0 4096 4194304 > gpurun_out/r2b_tail_m8.log 2>&1; echo "rc $?" >> gpurun_out/r2b_tail_m8.log
head -12 gpurun_out/r2b_tail_m8.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_long.py tests/test_gpu_fullsize.py tests/test_gpu_sweep.py -m gpu -x -q -s > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2b_pytest.log
tail -15 gpurun_out/r2b_pytest.log
timeout 300 python bench.py --steps 2 --warmup 3 --detail > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; tail -25 gpurun_out/r2b_bench.err
