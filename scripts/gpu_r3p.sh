mkdir -p gpurun_out
PK_POLAR_LANES_G=1 timeout 900 python -X faulthandler -m pytest tests/test_gpu_polar.py -m gpu -q -x 2>&1 | tail -2
for cb in 0 1 0 1; do
  for cfg in "1 262144" "16 32768" "32 16384"; do
    set -- $cfg
    echo -n "cb=$cb: "; PK_POLAR_LANES_CB=$cb PK_POLAR_LANES_G=1 timeout 200 python profiles/prof_polar.py $1 $2 2.0 2>&1 | tail -1
  done
done
