set -x
mkdir -p gpurun_out
timeout 900 python -X faulthandler -m pytest tests/test_gpu_polar.py -m gpu -q -x > gpurun_out/r3d_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r3d_pytest.log
tail -3 gpurun_out/r3d_pytest.log
for G in 1 2; do PK_POLAR_LANES_G=$G timeout 200 python profiles/prof_polar.py 1 262144 2.0 > gpurun_out/r3d_polar_L1_G$G.log 2>&1; tail -1 gpurun_out/r3d_polar_L1_G$G.log; done
for G in 2 4; do PK_POLAR_LANES_G=$G timeout 200 python profiles/prof_polar.py 8 65536 2.0 > gpurun_out/r3d_polar_L8_G$G.log 2>&1; tail -1 gpurun_out/r3d_polar_L8_G$G.log; done
PK_POLAR_LANES_G=2 timeout 200 python profiles/prof_polar.py 16 32768 2.0 > gpurun_out/r3d_polar_L16_G2.log 2>&1; tail -1 gpurun_out/r3d_polar_L16_G2.log
PK_POLAR_LANES_G=1 timeout 200 python profiles/prof_polar.py 32 16384 2.0 > gpurun_out/r3d_polar_L32_G1.log 2>&1; tail -1 gpurun_out/r3d_polar_L32_G1.log
