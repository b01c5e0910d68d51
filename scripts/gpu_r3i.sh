set -x
mkdir -p gpurun_out
timeout 900 python -X faulthandler -m pytest tests/test_gpu_polar.py tests/test_gpu_sweep.py tests/test_gpu_bridge.py -m gpu -q -x > gpurun_out/r3i_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r3i_pytest.log
tail -4 gpurun_out/r3i_pytest.log
timeout 200 python profiles/prof_polar.py 8 65536 2.0 | tail -1
