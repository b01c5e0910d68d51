set -x
mkdir -p gpurun_out
timeout 2400 python -X faulthandler -m pytest tests -m gpu -q > gpurun_out/r2r_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2r_pytest.log
grep -n "FAILED\|passed\|failed\|rc " gpurun_out/r2r_pytest.log | tail -20
timeout 900 python bench.py --steps 2 --warmup 1 --detail > gpurun_out/r2r_bench1.json 2> gpurun_out/r2r_bench1.err; echo "rc $?"; tail -5 gpurun_out/r2r_bench1.err
timeout 600 python bench.py --workload polar --steps 1 --warmup 1 > gpurun_out/r2r_polar.json 2> gpurun_out/r2r_polar.err; echo "rc $?"; tail -5 gpurun_out/r2r_polar.err
