set -x
mkdir -p gpurun_out
timeout 2400 python -X faulthandler -m pytest tests -m gpu -q > gpurun_out/r3s_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r3s_pytest.log
grep -n "FAILED\|passed\|failed\|rc " gpurun_out/r3s_pytest.log | tail -6
timeout 600 python bench.py --workload polar --steps 1 --warmup 1 > gpurun_out/r3s_polar.json 2> gpurun_out/r3s_polar.err; echo "rc $?"
timeout 900 python bench.py --steps 2 --warmup 1 > gpurun_out/r3s_bench.json 2> gpurun_out/r3s_bench.err; echo "rc $?"
