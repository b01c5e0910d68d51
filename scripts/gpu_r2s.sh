set -x
mkdir -p gpurun_out
timeout 600 python -X faulthandler -m pytest tests/test_gpu_fer.py -m gpu -q -k "63_39_9 or 63_51_5" > gpurun_out/r2s_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2s_pytest.log
tail -4 gpurun_out/r2s_pytest.log
timeout 900 python bench.py --workload large --steps 1 --warmup 1 > gpurun_out/r2s_large.json 2> gpurun_out/r2s_large.err; echo "rc $?"; tail -3 gpurun_out/r2s_large.err
timeout 600 python bench.py --steps 1 --warmup 1 --no-side > gpurun_out/r2s_plain.json 2> gpurun_out/r2s_plain.err && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r2s.csv python bench.py --steps 1 --warmup 1 --no-side > gpurun_out/r2s_ncu.log 2>&1
tail -2 gpurun_out/r2s_ncu.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_r2s_polar.csv python bench.py --workload polar --steps 1 --warmup 1 > gpurun_out/r2s_ncu_polar.log 2>&1
tail -2 gpurun_out/r2s_ncu_polar.log
