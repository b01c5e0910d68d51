mkdir -p gpurun_out
timeout 900 python -X faulthandler -m pytest tests/test_gpu_polar.py -m gpu -q -x 2>&1 | tail -2
for r in 1 2; do
for cfg in "1 262144" "8 65536" "16 32768" "32 16384" "6 65536"; do
  set -- $cfg
  timeout 200 python profiles/prof_polar.py $1 $2 2.0 2>&1 | tail -1
done
done
