#!/bin/bash
# One-GPU experiment batch queued at the end of round 1 (no GPU budget was left to run it):
#   gpurun --timeout 900 -- 'bash tools/runs/run_round2_experiments.sh'
# Everything lands in gpurun_out/round2/.
set -u
O=gpurun_out/round2; mkdir -p $O
# 1. CTA-pair tcgen05 GEMM: correctness first (a hang here must not eat the whole call)
DE_B200_TEST_EXPERIMENTAL=1 timeout 120 python -m pytest tests/test_gemm_tcgen05.py -x -q -k cta_pair > $O/pair_gemm_test.log 2>&1
echo "pair gemm test rc=$?" | tee -a $O/summary.txt
if grep -q " passed" $O/pair_gemm_test.log && ! grep -q failed $O/pair_gemm_test.log; then
  DE_B200_TEST_EXPERIMENTAL=1 timeout 120 python tools/bench_gemm.py > $O/bench_gemm_pair.log 2>&1
fi
# 2. own radix sort vs torch.sort (CUB) and the synthetic-small step with either sort
timeout 200 python tools/bench_sort.py > $O/bench_sort.log 2>&1
SYN=examples/benchmarks/synthetic_models/main.py
for S in cub own; do
  DE_B200_SORT=$S timeout 200 python $SYN --model small --optimizer adagrad --batch_size 65536 --alpha 1.05 \
    --num_steps 30 --num_data_batches 2 --amp 2>&1 | grep -E '^\{' | tail -1 > $O/synth_small_sort_$S.json
done
# 3. same-address atomic contention of the tiny MLPerf tables in the SGD scatter kernel
timeout 200 python tools/profile_step.py --model dlrm-mlperf --out $O/step_mlperf.txt > /dev/null 2>&1
timeout 200 python tools/profile_step.py --model dlrm-mlperf --min-table-rows 1000000 --out $O/step_mlperf_min1m.txt > /dev/null 2>&1
grep -h "scatter_add_bwd\|lookup_fwd" $O/step_mlperf.txt $O/step_mlperf_min1m.txt | tee -a $O/summary.txt
# 3a. tiny-table shared-memory scatter: numerics, then the scatter kernel time with it
DE_B200_TEST_EXPERIMENTAL=1 timeout 200 python -m pytest tests/test_fused_optimizers.py -x -q -k tiny > $O/tiny_test.log 2>&1
echo "tiny tables test rc=$?" | tee -a $O/summary.txt
DE_B200_TINY_TABLES=1 timeout 200 python tools/profile_step.py --model dlrm-mlperf --out $O/step_mlperf_tiny.txt > /dev/null 2>&1
grep -h "scatter_add" $O/step_mlperf_tiny.txt | tee -a $O/summary.txt
# 3b. double-buffered interaction backward: numerics, then the step with it
DE_B200_TEST_EXPERIMENTAL=1 timeout 200 python -m pytest tests/test_dense_kernels.py -x -q -k v2 > $O/interact_v2_test.log 2>&1
echo "interact v2 test rc=$?" | tee -a $O/summary.txt
DE_B200_INTERACT_V2=1 timeout 200 python tools/profile_step.py --model dlrm-mlperf --out $O/step_mlperf_interact_v2.txt > /dev/null 2>&1
grep -h "interact_bwd" $O/step_mlperf.txt $O/step_mlperf_interact_v2.txt | tee -a $O/summary.txt
# 3c. TMA bulk-copy forward: numerics (1 GPU), then the lookup kernel time with it
DE_B200_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_dist_gpu.py -x -q -k "bulk_lookup and 1-" > $O/bulk_lookup_test.log 2>&1
echo "bulk lookup test rc=$?" | tee -a $O/summary.txt
DE_B200_LOOKUP_BULK=1 timeout 200 python tools/profile_step.py --model dlrm-mlperf --out $O/step_mlperf_bulk.txt > /dev/null 2>&1
grep -h "lookup_fwd" $O/step_mlperf.txt $O/step_mlperf_bulk.txt | tee -a $O/summary.txt
# 4. headline
timeout 300 python bench.py --steps 50 --warmup 10 | tail -1 | tee -a $O/summary.txt

# 5. (multi-GPU, separate call) cross-rank timeline of graph replays, e.g. with gpurun --gpus 8:
#   python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29650 \
#     bench.py --gpus 8 --steps 20 --warmup 5 --no-e2e --profile gpurun_out/round2/prof_n8.txt \
#     --profile-all-ranks --profile-graph 1
#   python tools/critical_path.py gpurun_out/round2/prof_n8.txt --step 2 > gpurun_out/round2/critical_path_n8.txt
# 6. (2 GPUs) bucketed all-reduce overlap: numerics, then A/B
#   DE_B200_AR_OVERLAP=1 python -m pytest tests/test_dist_gpu.py -q -x -k "dlrm_fast_world2"
#   for V in 0 1; do DE_B200_AR_OVERLAP=$V python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 \
#     --master-addr 127.0.0.1 --master-port 2966$V bench.py --gpus 2 --steps 50 --warmup 10 --no-e2e | tail -1; done
# 7. (4 or 8 GPUs) does the persistent scatter grid starve the overlapped GEMMs?
#   for C in 4 2 1; do DE_B200_EMB_BLOCKS_PER_SM=$C python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 \
#     --master-addr 127.0.0.1 --master-port 2967$C bench.py --gpus 8 --steps 50 --warmup 10 --no-e2e | tail -1; done
#   then the same with DE_B200_VEC8_PULL=1
# 8. (2 GPUs, then 8) replicated tiny tables in the fast step: numerics, then the headline with them
#   DE_B200_TEST_EXPERIMENTAL=1 python -m pytest tests/test_dist_gpu.py -q -x -k replicated_tables
#   python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29680 \
#     bench.py --gpus 8 --steps 50 --warmup 10 --no-e2e --data-parallel-threshold 300000 | tail -1
