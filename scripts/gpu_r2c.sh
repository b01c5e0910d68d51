set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sweep.py -m gpu -x -q > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2c_pytest.log
tail -15 gpurun_out/r2c_pytest.log
timeout 900 python bench.py --steps 2 --warmup 1 --detail --no-side > gpurun_out/r2c_bench1.json 2> gpurun_out/r2c_bench1.err; echo "rc $?"; tail -30 gpurun_out/r2c_bench1.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 1 --detail --no-side > gpurun_out/r2c_bench2.json 2> gpurun_out/r2c_bench2.err; echo "rc $?"; tail -30 gpurun_out/r2c_bench2.err
timeout 600 python bench.py --workload large --steps 1 --warmup 1 > gpurun_out/r2c_large.json 2> gpurun_out/r2c_large.err; echo "rc $?"; tail -5 gpurun_out/r2c_large.err
