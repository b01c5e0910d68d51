set -x
mkdir -p gpurun_out
N=${N:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 2 --warmup 1 > gpurun_out/r2v_bench_$N.json 2> gpurun_out/r2v_bench_$N.err; echo "rc $?"; tail -3 gpurun_out/r2v_bench_$N.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --workload polar --steps 1 --warmup 1 > gpurun_out/r2v_polar_$N.json 2> gpurun_out/r2v_polar_$N.err; echo "rc $?"; tail -3 gpurun_out/r2v_polar_$N.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --workload large --steps 1 --warmup 1 > gpurun_out/r2v_large_$N.json 2> gpurun_out/r2v_large_$N.err; echo "rc $?"; tail -3 gpurun_out/r2v_large_$N.err
timeout 600 python -m pytest tests -m gpu -q -k "two_gpu or gpus" > gpurun_out/r2v_pytest_$N.log 2>&1; tail -3 gpurun_out/r2v_pytest_$N.log
