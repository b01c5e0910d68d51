set -x
mkdir -p gpurun_out
timeout 2400 python -X faulthandler -m pytest tests -m gpu -q > gpurun_out/r2h_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2h_pytest.log
grep -n "FAILED\|passed\|failed" gpurun_out/r2h_pytest.log | tail -20
