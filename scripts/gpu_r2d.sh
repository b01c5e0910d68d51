set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2d_pytest.log
tail -25 gpurun_out/r2d_pytest.log
timeout 120 python profiles/prof_tail.py 5 3 -1 0.0 65536 > gpurun_out/r2d_tail_m5.log 2>&1; head -9 gpurun_out/r2d_tail_m5.log
timeout 120 python profiles/prof_tail.py 7 10 15 3.0 4096 4194304 > gpurun_out/r2d_tail_m7.log 2>&1; head -9 gpurun_out/r2d_tail_m7.log
timeout 120 python profiles/prof_tail.py 8 15 15 4.0 4096 4194304 > gpurun_out/r2d_tail_m8.log 2>&1; head -9 gpurun_out/r2d_tail_m8.log
timeout 600 python bench.py --workload large --steps 1 --warmup 1 > gpurun_out/r2d_large.json 2> gpurun_out/r2d_large.err; echo "rc $?"; tail -5 gpurun_out/r2d_large.err
timeout 600 python bench.py --workload polar --steps 1 --warmup 1 > gpurun_out/r2d_polar.json 2> gpurun_out/r2d_polar.err; echo "rc $?"; tail -5 gpurun_out/r2d_polar.err
