#!/bin/bash
# Round-2 scaling run on N GPUs: headline at 1..N, per-rank graph-replay timeline at N, A/B of the
# all-reduce overlap, NCCL-collectives baseline, and the multi-GPU tests that were gated in round 1.
#   gpurun --gpus N --timeout 1500 -- 'bash tools/runs/r2_scale.sh N [tests]'
set -u
N=${1:-2}
TESTS=${2:-tests}
O=gpurun_out/r2s_n$N; mkdir -p $O
export DE_B200_FLAG_TIMEOUT_CYCLES=${DE_B200_FLAG_TIMEOUT_CYCLES:-30000000000}
run_bench() {  # n, tag, args...
  local n=$1 tag=$2; shift 2
  if [ $n -eq 1 ]; then
    CUDA_VISIBLE_DEVICES=0 timeout 600 python bench.py --gpus 1 "$@" > $O/bench_${tag}_n1.log 2>&1
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 \
      --master-port $((29500 + RANDOM % 400)) bench.py --gpus $n "$@" > $O/bench_${tag}_n$n.log 2>&1
  fi
  echo "bench $tag n=$n rc=$?" | tee -a $O/summary.txt
  grep -E '^\{' $O/bench_${tag}_n$n.log | tail -1 > $O/bench_${tag}_n$n.json
  python - "$O/bench_${tag}_n$n.json" <<'PY' | tee -a $O/summary.txt
import json, sys
try:
  d = json.loads(open(sys.argv[1]).read())
  v = d.get("verify") or {}
  print("   ms/step %.4f  value %.2f M  e2e %s  verify_ok %s rel_l2 %s  final_loss %s" % (
      d.get("ms_per_step", -1), d.get("value", 0) / 1e6,
      ("%.2f M" % (d["e2e"]["value"] / 1e6)) if d.get("e2e") else None, v.get("ok"),
      v.get("rel_l2_err_of_update"), d.get("final_loss")))
except Exception as e:
  print("   no JSON:", e)
PY
}
for n in 1 2 4 8; do
  [ $n -le $N ] || continue
  run_bench $n fused --steps 50 --warmup 10
done
# per-rank timeline of CUDA-graph replays at N GPUs (true device timeline, no launch skew)
if [ $N -gt 1 ]; then
  run_bench $N prof --steps 20 --warmup 5 --no-e2e --no-verify --profile $O/prof_n$N.txt \
    --profile-all-ranks --profile-graph 1
  python tools/critical_path.py $O/prof_n$N.txt --step 2 > $O/critical_path_n$N.txt 2>&1
  rm -f $O/prof_n$N.txt*.trace.json
  DE_B200_AR_OVERLAP=0 run_bench $N noaroverlap --steps 50 --warmup 10 --no-e2e --no-verify
fi
for n in 1 2 4 8; do
  [ $n -le $N ] || continue
  run_bench $n nccl --backend torch --trainer autograd --steps 20 --warmup 5 --no-e2e
done
if [ "$TESTS" = "tests" ]; then
  timeout 900 python -m pytest tests/test_dist_gpu.py tests/test_dense_kernels.py -m gpu -q -p no:cacheprovider \
    -k "fuzz or subgroups or replicated or interaction or dlrm_fast" > $O/pytest_subset.log 2>&1
  echo "pytest subset rc=$?" | tee -a $O/summary.txt
  tail -4 $O/pytest_subset.log | tee -a $O/summary.txt
fi
