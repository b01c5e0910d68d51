set -x
mkdir -p gpurun_out
timeout 900 python -X faulthandler -m pytest tests/test_gpu_polar.py -m gpu -q > gpurun_out/r2y_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2y_pytest.log
grep -n "FAILED\|passed\|failed\|rc \|Error" gpurun_out/r2y_pytest.log | tail -30
