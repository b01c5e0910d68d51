mkdir -p gpurun_out
P=polar-codes-with-bch-kernel_b200
for round in 1 2; do
for v in old new; do
  cp scripts/ab/libpkb200_$v.so $P/libpkb200.so
  for cfg in "1 262144" "8 65536" "16 32768" "32 16384"; do
    set -- $cfg
    echo -n "$v r$round: "; timeout 200 python profiles/prof_polar.py $1 $2 2.0 2>&1 | tail -1
  done
done
done
