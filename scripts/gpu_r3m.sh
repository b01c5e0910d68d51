set -x
mkdir -p gpurun_out
timeout 900 python -X faulthandler -m pytest tests/test_gpu_polar.py -m gpu -q -x > gpurun_out/r3m_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r3m_pytest.log
tail -5 gpurun_out/r3m_pytest.log
for cfg in "1 262144" "8 65536" "32 16384" "6 65536" "24 16384"; do set -- $cfg; timeout 200 python profiles/prof_polar.py $1 $2 2.0 2>&1 | tail -1; done
PK_POLAR_LANES=0 timeout 200 python profiles/prof_polar.py 6 65536 2.0 2>&1 | tail -1
PK_POLAR_LANES=0 timeout 200 python profiles/prof_polar.py 24 16384 2.0 2>&1 | tail -1
