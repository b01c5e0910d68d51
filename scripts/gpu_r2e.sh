set -x
mkdir -p gpurun_out
timeout 600 python -X faulthandler -m pytest tests/test_gpu_polar.py tests/test_gpu_bridge.py tests/test_gpu_ext.py -m gpu -x -q > gpurun_out/r2e_pytest1.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2e_pytest1.log
tail -25 gpurun_out/r2e_pytest1.log
timeout 600 python -X faulthandler bench.py --workload polar --steps 1 --warmup 1 > gpurun_out/r2e_polar.json 2> gpurun_out/r2e_polar.err; echo "rc $?"; tail -30 gpurun_out/r2e_polar.err
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_gpu_polar.py --deselect tests/test_gpu_bridge.py --deselect tests/test_gpu_ext.py -x > gpurun_out/r2e_pytest2.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2e_pytest2.log
tail -25 gpurun_out/r2e_pytest2.log
timeout 120 python profiles/prof_tail.py 5 3 -1 0.0 65536 > gpurun_out/r2e_tail_m5.log 2>&1; head -9 gpurun_out/r2e_tail_m5.log
timeout 120 python profiles/prof_tail.py 7 10 15 3.0 4096 4194304 > gpurun_out/r2e_tail_m7.log 2>&1; head -9 gpurun_out/r2e_tail_m7.log
timeout 120 python profiles/prof_tail.py 8 15 15 4.0 4096 4194304 > gpurun_out/r2e_tail_m8.log 2>&1; head -9 gpurun_out/r2e_tail_m8.log
timeout 900 python bench.py --workload large --steps 1 --warmup 1 > gpurun_out/r2e_large.json 2> gpurun_out/r2e_large.err; echo "rc $?"; tail -5 gpurun_out/r2e_large.err
