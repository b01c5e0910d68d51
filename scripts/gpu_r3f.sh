set -x
mkdir -p gpurun_out
timeout 900 python -X faulthandler -m pytest tests/test_gpu_sweep.py -m gpu -q -x > gpurun_out/r3f_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r3f_pytest.log
tail -30 gpurun_out/r3f_pytest.log
