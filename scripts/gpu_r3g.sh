set -x
mkdir -p gpurun_out
N=${N:-8}
nvidia-smi topo -m > gpurun_out/r3g_topo.txt 2>&1; lscpu | grep -i "numa\|socket\|^CPU(s)" >> gpurun_out/r3g_topo.txt; for d in /sys/bus/pci/devices/*; do if [ "$(cat $d/vendor 2>/dev/null)" = "0x10de" ]; then echo "$d $(cat $d/numa_node)"; fi; done >> gpurun_out/r3g_topo.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 2 --warmup 1 --no-side > gpurun_out/r3g_bench_$N.json 2> gpurun_out/r3g_bench_$N.err; echo "rc $?"; tail -3 gpurun_out/r3g_bench_$N.err
