#!/bin/bash
# Round-2 GPU check: the GPU test-suite at the visible world size, then the headline bench and the
# NCCL-collectives baseline (backend=torch, autograd trainer) at 1..N GPUs.
#   gpurun --gpus N --timeout 1500 -- 'bash tools/runs/r2_check.sh N [quick]'
set -u
N=${1:-1}
MODE=${2:-full}
O=gpurun_out/r2_n$N; mkdir -p $O
export DE_B200_FLAG_TIMEOUT_CYCLES=${DE_B200_FLAG_TIMEOUT_CYCLES:-30000000000}
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv > $O/gpus.csv 2>&1
if [ "$MODE" != "quick" ]; then
  timeout 1200 python -m pytest tests -m gpu -x -q -p no:cacheprovider > $O/pytest_gpu.log 2>&1
  echo "pytest rc=$?" | tee $O/summary.txt
  tail -5 $O/pytest_gpu.log | tee -a $O/summary.txt
fi
run_bench() {  # n, tag, extra args
  local n=$1 tag=$2; shift 2
  if [ $n -eq 1 ]; then
    CUDA_VISIBLE_DEVICES=0 timeout 600 python bench.py --gpus 1 "$@" > $O/bench_${tag}_n1.log 2>&1
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 \
      --master-port $((29500 + RANDOM % 400)) bench.py --gpus $n "$@" > $O/bench_${tag}_n$n.log 2>&1
  fi
  echo "bench $tag n=$n rc=$?" | tee -a $O/summary.txt
  grep -E '^\{' $O/bench_${tag}_n$n.log | tail -1 | tee $O/bench_${tag}_n$n.json | cut -c1-400 | tee -a $O/summary.txt
}
for n in 1 2 4 8; do
  [ $n -le $N ] || continue
  run_bench $n fused --steps 50 --warmup 10
done
for n in 1 2 4 8; do
  [ $n -le $N ] || continue
  run_bench $n nccl --backend torch --trainer autograd --steps 20 --warmup 5 --no-e2e
done
