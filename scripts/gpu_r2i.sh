set -x
mkdir -p gpurun_out
timeout 600 python -X faulthandler -m pytest tests/test_gpu_polar.py -m gpu -q -x > gpurun_out/r2i_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2i_pytest.log
tail -15 gpurun_out/r2i_pytest.log
for G in 1 2 4; do PK_POLAR_SC_G=$G timeout 200 python profiles/prof_polar.py 1 262144 2.0 > gpurun_out/r2i_polar_G$G.log 2>&1; tail -3 gpurun_out/r2i_polar_G$G.log; done
PK_POLAR_SC_G=0 timeout 200 python profiles/prof_polar.py 1 262144 2.0 > gpurun_out/r2i_polar_old.log 2>&1; tail -3 gpurun_out/r2i_polar_old.log
