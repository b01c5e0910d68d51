set -x
mkdir -p gpurun_out
timeout 2400 python -X faulthandler -m pytest tests -m gpu -q > gpurun_out/r3o_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r3o_pytest.log
grep -n "FAILED\|passed\|failed\|rc " gpurun_out/r3o_pytest.log | tail -6
timeout 600 python bench.py --workload polar --steps 1 --warmup 1 > gpurun_out/r3o_polar.json 2> gpurun_out/r3o_polar.err; echo "rc $?"
