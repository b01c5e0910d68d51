set -x
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r3c_smoke.log 2>&1; echo "smoke rc $?"; tail -6 gpurun_out/r3c_smoke.log
timeout 2400 python -X faulthandler -m pytest tests -m gpu -q > gpurun_out/r3c_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r3c_pytest.log
grep -n "FAILED\|passed\|failed\|rc " gpurun_out/r3c_pytest.log | tail -10
( time timeout 1200 python bench.py > gpurun_out/r3c_bench_default.json 2> gpurun_out/r3c_bench_default.err ) 2>&1 | tail -4; echo "rc $?"
( time timeout 1200 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r3c_bench_ref.json 2> gpurun_out/r3c_bench_ref.err ) 2>&1 | tail -4
timeout 600 python bench.py --workload polar --steps 1 --warmup 1 > gpurun_out/r3c_polar.json 2> gpurun_out/r3c_polar.err; echo "rc $?"
