#!/bin/bash
# 2-GPU run: NVLink store probe, headline at N=2 (all-reduce overlap on / off), 2-rank tests of
# the hand-scheduled synthetic step.   gpurun --gpus 2 --timeout 900 -- 'bash tools/runs/r2_p2p.sh'
set -u
O=gpurun_out/r2_p2p; mkdir -p $O
export DE_B200_FLAG_TIMEOUT_CYCLES=${DE_B200_FLAG_TIMEOUT_CYCLES:-30000000000}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29611 tools/bench_p2p_store.py > $O/p2p_store.log 2>&1
echo "p2p rc=$?" | tee $O/summary.txt
grep -E '^\{' $O/p2p_store.log > $O/p2p_store.jsonl; wc -l $O/p2p_store.jsonl | tee -a $O/summary.txt
for V in 1 0; do
  DE_B200_AR_OVERLAP=$V timeout 300 $TR --master-port 2962$V bench.py --gpus 2 --steps 50 --warmup 10 --no-e2e \
    > $O/bench_n2_ar$V.log 2>&1
  grep -E '^\{' $O/bench_n2_ar$V.log | tail -1 > $O/bench_n2_ar$V.json
  echo "n2 ar_overlap=$V: $(cut -c1-200 $O/bench_n2_ar$V.json)" | tee -a $O/summary.txt
done
timeout 600 python -m pytest tests/test_dist_gpu.py -m gpu -q -x -p no:cacheprovider -k "synthetic_fast_world2" > $O/pytest.log 2>&1
echo "pytest rc=$?" | tee -a $O/summary.txt; tail -3 $O/pytest.log | tee -a $O/summary.txt
