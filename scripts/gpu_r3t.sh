set -x
mkdir -p gpurun_out
timeout 300 python profiles/prof_polar.py 1 65536 2.0 > gpurun_out/r3t_L1_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_polar_lanes -s 1 -c 1 -o gpurun_out/prof_r3t_lanes_L1_G2 python profiles/prof_polar.py 1 65536 2.0 > gpurun_out/r3t_L1_ncu.log 2>&1
tail -1 gpurun_out/r3t_L1_plain.log
timeout 300 python profiles/prof_polar.py 32 4096 2.0 > gpurun_out/r3t_L32_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_polar_lanes -s 1 -c 1 -o gpurun_out/prof_r3t_lanes_L32_G1 python profiles/prof_polar.py 32 4096 2.0 > gpurun_out/r3t_L32_ncu.log 2>&1
tail -1 gpurun_out/r3t_L32_plain.log
