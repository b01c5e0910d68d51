set -x
mkdir -p gpurun_out
timeout 900 python -X faulthandler -m pytest tests/test_gpu_polar.py -m gpu -q -x -rs > gpurun_out/r3n_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r3n_pytest.log
tail -12 gpurun_out/r3n_pytest.log
