#!/bin/bash
# Round-2 single-GPU run: headline + verify, synthetic zoo (hand-scheduled vs autograd trainer),
# sort micro-benchmark, the single-GPU tests touched this round, and an ncu capture of the hot
# embedding / interaction kernels.   gpurun --timeout 1500 -- 'bash tools/runs/r2_n1.sh'
set -u
O=gpurun_out/r2_n1; mkdir -p $O
export DE_B200_FLAG_TIMEOUT_CYCLES=${DE_B200_FLAG_TIMEOUT_CYCLES:-30000000000}
timeout 300 python bench.py --steps 50 --warmup 10 > $O/bench_n1.log 2>&1
echo "bench rc=$?" | tee $O/summary.txt
grep -E '^\{' $O/bench_n1.log | tail -1 > $O/bench_n1.json; cut -c1-300 $O/bench_n1.json | tee -a $O/summary.txt
timeout 300 python bench.py --steps 30 --warmup 10 --alpha 1.05 --no-e2e > $O/bench_n1_alpha.log 2>&1
grep -E '^\{' $O/bench_n1_alpha.log | tail -1 > $O/bench_n1_alpha105.json; cut -c1-200 $O/bench_n1_alpha105.json | tee -a $O/summary.txt
SYN=examples/benchmarks/synthetic_models/main.py
for M in tiny small; do
  for T in fast autograd; do
    timeout 300 python $SYN --model $M --optimizer adagrad --batch_size 65536 --alpha 1.05 --num_steps 30 \
      --num_data_batches 2 --amp --trainer $T 2>&1 | grep -E '^\{' | tail -1 > $O/synth_${M}_${T}.json
    echo "synth $M $T: $(cut -c1-120 $O/synth_${M}_${T}.json)" | tee -a $O/summary.txt
  done
done
DE_B200_SORT=cub timeout 300 python $SYN --model small --optimizer adagrad --batch_size 65536 --alpha 1.05 \
  --num_steps 30 --num_data_batches 2 --amp --trainer fast 2>&1 | grep -E '^\{' | tail -1 > $O/synth_small_fast_cubsort.json
echo "synth small fast (CUB sort): $(cut -c1-120 $O/synth_small_fast_cubsort.json)" | tee -a $O/summary.txt
timeout 200 python tools/bench_sort.py --sizes 1703936,16777216 --bits 20,28,32 > $O/bench_sort.log 2>&1
tail -6 $O/bench_sort.log | tee -a $O/summary.txt
timeout 200 python examples/benchmarks/benchmark.py > $O/op_benchmark.log 2>&1; tail -8 $O/op_benchmark.log | tee -a $O/summary.txt
timeout 900 python -m pytest tests/test_radix_sort.py tests/test_fused_optimizers.py tests/test_embedding_ops.py \
  tests/test_integer_lookup.py tests/test_dense_kernels.py -m gpu -q -x -p no:cacheprovider > $O/pytest_n1.log 2>&1
echo "pytest rc=$?" | tee -a $O/summary.txt; tail -3 $O/pytest_n1.log | tee -a $O/summary.txt
timeout 600 python -m pytest tests/test_dist_gpu.py -m gpu -q -x -p no:cacheprovider -k "synthetic_fast_world1 or dlrm_fast_world1" > $O/pytest_fast.log 2>&1
echo "pytest fast rc=$?" | tee -a $O/summary.txt; tail -3 $O/pytest_fast.log | tee -a $O/summary.txt
# ncu: one launch of each hot kernel in steady state (eager launches, step 4 of 5)
timeout 600 ncu --set full --clock-control none --import-source on \
  -k regex:'lookup_fwd_kernel|scatter_add_bwd_kernel|interact_bwd_v2_kernel|interact_fwd_kernel|relu_bwd_bias_kernel' \
  -s 24 -c 9 -f -o $O/ncu_dlrm_n1 python bench.py --steps 2 --warmup 3 --no-e2e --no-verify --cuda-graph 0 \
  > $O/ncu_dlrm.log 2>&1
echo "ncu rc=$?" | tee -a $O/summary.txt
timeout 600 ncu --set full --clock-control none --import-source on \
  -k regex:'balanced_update_kernel|segment_update_kernel|build_keys_kernel|digit_scatter_kernel|lookup_fwd_kernel' \
  -s 12 -c 8 -f -o $O/ncu_synth_small python $SYN --model small --optimizer adagrad --batch_size 65536 \
  --alpha 1.05 --num_steps 2 --num_data_batches 1 --amp --trainer fast --cuda_graph 0 > $O/ncu_synth.log 2>&1
echo "ncu synth rc=$?" | tee -a $O/summary.txt
ls -la $O | tail -20
