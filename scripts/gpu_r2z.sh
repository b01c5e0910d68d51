set -x
mkdir -p gpurun_out
N=${N:-8}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 2 --warmup 1 --no-side > gpurun_out/r2z_bench_$N.json 2> gpurun_out/r2z_bench_$N.err; echo "rc $?"; tail -3 gpurun_out/r2z_bench_$N.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --workload polar --steps 1 --warmup 1 > gpurun_out/r2z_polar_$N.json 2> gpurun_out/r2z_polar_$N.err; echo "rc $?"; tail -3 gpurun_out/r2z_polar_$N.err
