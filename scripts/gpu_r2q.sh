set -x
mkdir -p gpurun_out
PK_POLAR_LANES_G=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_polar_lanes -s 1 -c 1 -o gpurun_out/prof_r2q_lanes_L1_G2 python profiles/prof_polar.py 1 65536 2.0 > gpurun_out/r2q_L1_ncu.log 2>&1
tail -2 gpurun_out/r2q_L1_ncu.log
PK_POLAR_LANES_G=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_polar_lanes -s 1 -c 1 -o gpurun_out/prof_r2q_lanes_L32_G1 python profiles/prof_polar.py 32 4096 2.0 > gpurun_out/r2q_L32_ncu.log 2>&1
tail -2 gpurun_out/r2q_L32_ncu.log
