#!/bin/bash
# Builder-run throughput lines at the GPU counts BASELINE.json names (one 8 x B200 box): products-shape at 8 GPUs under the
# driver's own command (--steps 20 --warmup 5, twice) and with 200 steps, both exchanges; Reddit-shape at 8 / 4 / 2 / 1 GPUs;
# papers100M-shape at 8 GPUs.  Output: gpurun_out/r02_mg_*.json (one JSON line each).
cd "$(dirname "$0")/.."
OUT=gpurun_out
run() {  # name, ngpus, visible devices, port, extra args...
  local name=$1 n=$2 vis=$3 port=$4; shift 4
  if [ "$n" = "1" ]; then
    CUDA_VISIBLE_DEVICES=$vis timeout 900 python bench.py --gpus 1 --no-cpu-baseline --no-spmm "$@" > $OUT/r02_mg_$name.json 2> $OUT/r02_mg_$name.err
  else
    CUDA_VISIBLE_DEVICES=$vis timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 \
      --master-port $port bench.py --gpus $n --no-cpu-baseline --no-spmm "$@" > $OUT/r02_mg_$name.json 2> $OUT/r02_mg_$name.err
  fi
}
ALL=0,1,2,3,4,5,6,7
run products_n8_peer_20a 8 $ALL 29601 --steps 20 --warmup 5
run products_n8_peer_20b 8 $ALL 29602 --steps 20 --warmup 5
run products_n8_peer_200 8 $ALL 29603 --steps 200 --warmup 10
run products_n8_nccl_200 8 $ALL 29604 --steps 200 --warmup 10 --exchange nccl
run products_n8_nccl_20 8 $ALL 29605 --steps 20 --warmup 5 --exchange nccl
run reddit_n8 8 $ALL 29606 --workload reddit --steps 200 --warmup 10
# the smaller GPU counts side by side on disjoint GPUs of the same box
run reddit_n4 4 0,1,2,3 29607 --workload reddit --steps 200 --warmup 10 &
run reddit_n2 2 4,5 29608 --workload reddit --steps 200 --warmup 10 &
run reddit_n1 1 6 29609 --workload reddit --steps 200 --warmup 10 &
wait
run products_n4 4 0,1,2,3 29610 --steps 200 --warmup 10 &
run products_n2 2 4,5 29611 --steps 200 --warmup 10 &
run products_n1 1 6 29612 --steps 200 --warmup 10 &
wait
run papers_n8 8 $ALL 29613 --workload papers --steps 200 --warmup 10
for f in $OUT/r02_mg_*.json; do echo "$f: $(head -c 160 $f)"; done
