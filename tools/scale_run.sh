#!/bin/bash
# usage: tools/scale_run.sh N [N ...]   -> gpurun_out/scale_bench_nN.json, gpurun_out/scale_dist_nN.txt
# bench.py (exposure-sharded, weak scaling) and tools/dist_refocus_check.py (slice-sharded stack) on N GPUs of one box.
mkdir -p gpurun_out
port=29600
for n in "$@"; do
  port=$((port+1))
  if [ "$n" = "1" ]; then
    python bench.py --gpus 1 --steps 50 --warmup 3 --no-cpu-baseline > gpurun_out/scale_bench_n1.json 2> gpurun_out/scale_bench_n1.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port \
        bench.py --gpus $n --steps 50 --warmup 3 2> gpurun_out/scale_bench_n$n.err | grep '^{' > gpurun_out/scale_bench_n$n.json
    port=$((port+1))
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port \
        tools/dist_refocus_check.py 2> gpurun_out/scale_dist_n$n.err | grep '^{' > gpurun_out/scale_dist_n$n.txt
  fi
  echo "n=$n"; cut -c1-260 gpurun_out/scale_bench_n$n.json; [ -f gpurun_out/scale_dist_n$n.txt ] && cut -c1-200 gpurun_out/scale_dist_n$n.txt
done
