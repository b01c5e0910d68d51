set -x
mkdir -p gpurun_out
N=${N:-8}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --workload polar --steps 1 --warmup 1 > gpurun_out/r3j_polar_$N.json 2> gpurun_out/r3j_polar_$N.err; echo "rc $?"; tail -2 gpurun_out/r3j_polar_$N.err
